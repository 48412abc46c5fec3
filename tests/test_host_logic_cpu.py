"""Host-side decision logic that needs no GPU: which search / window / shard form a problem gets, the deterministic row
spreads, the synthetic generators both bench arms share, and bench.py's clock sampler without NVML.  (The kernels these
rules choose between are covered by the -m gpu tests; here only the choice.)"""
import importlib.util
import os
import sys
from types import SimpleNamespace

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [p for p in (ROOT, os.path.join(ROOT, "multimodal-umap_b200")) if p not in sys.path]

from umap_b200 import dist as D, knn_pruned, knn_tc, layout, native, spectral  # noqa: E402


def _bench():
    spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    argv, sys.argv = sys.argv, ["bench.py"]
    try:
        spec.loader.exec_module(mod)
    finally:
        sys.argv = argv
    return mod


# --------------------------------------------------------------------------- cluster-pruned search: host rules
def test_spread_rows_is_a_deterministic_distinct_sample_that_does_not_alias_with_periodic_layouts():
    n, count = 1_000_000, 65536
    rows = knn_pruned._spread_rows(n, count, "cpu")
    assert rows.dtype == torch.int64 and rows.numel() <= count and rows.numel() >= 0.95 * count
    assert int(rows.min()) >= 0 and int(rows.max()) < n
    assert rows.unique().numel() == rows.numel()                       # distinct
    assert torch.equal(rows, knn_pruned._spread_rows(n, count, "cpu"))  # every rank picks the same rows
    # data laid out cluster by cluster modulo 1000: a stride-15 sample sees 200 of the 1000 clusters, this one all
    assert (rows % 1000).unique().numel() == 1000
    assert (torch.arange(0, n, n // count)[:count] % 1000).unique().numel() < 1000
    # asking for at least n rows gives every row once, in order
    assert torch.equal(knn_pruned._spread_rows(100, 100, "cpu"), torch.arange(100))
    assert torch.equal(knn_pruned._spread_rows(100, 5000, "cpu"), torch.arange(100))


def test_prune_rule_thresholds(monkeypatch):
    monkeypatch.delenv("MMUMAP_KNN_PRUNE", raising=False)
    assert knn_tc.prune_applicable(10_000_000, 128, 30)                # BASELINE configs[3]
    assert knn_tc.prune_applicable(knn_tc.PRUNE_MIN_ROWS, knn_tc.PRUNE_MAX_DIM, 63)
    assert not knn_tc.prune_applicable(knn_tc.PRUNE_MIN_ROWS - 1, 128, 30)
    assert not knn_tc.prune_applicable(158_915, 768, 15)               # configs[1] texts: too few rows, rows too long
    assert not knn_tc.prune_applicable(1_000_000, 768, 15)             # configs[2]: rows too long
    assert not knn_tc.prune_applicable(1_000_000, 128, 64)             # k + 1 must fit one 64-entry list
    monkeypatch.setenv("MMUMAP_KNN_PRUNE", "0")
    assert not knn_tc.prune_applicable(10_000_000, 128, 30)


# --------------------------------------------------------------------------- spectral initialisation: block widths, shard rule
def test_block_width_and_shard_rule(monkeypatch):
    assert spectral.block_width(2) == 8 and spectral.block_width(4) == 8
    assert spectral.block_width(5) == 16 and spectral.block_width(11) == 16
    assert spectral.block_width(16) == 32 and spectral.block_width(23) == 32
    for d in range(1, 24):
        assert spectral.block_width(d) in spectral.BLOCK_WIDTHS and spectral.block_width(d) >= d + 1
    wide = spectral.block_width(40)
    assert wide >= 41 + 8 and wide % 4 == 0 and wide not in spectral.BLOCK_WIDTHS   # falls back to the torch form
    c4 = SimpleNamespace(n_rows=10_000_000)
    c3 = SimpleNamespace(n_rows=1_000_000)
    assert not spectral.shardable(c4, 2)                               # one process: never a collective
    monkeypatch.setattr(D, "world", lambda: 8)
    assert spectral.shardable(c4, 2)                                   # 10M x 8 columns x 4 B = 320 MB: out of L2
    assert not spectral.shardable(c3, 16)                              # 1M x 32 x 4 B = 128 MB: L2 resident, measured slower sharded
    assert not spectral.shardable(c4, 2, method="lobpcg")
    monkeypatch.setenv("MMUMAP_SPECTRAL_SHARD", "0")
    assert not spectral.shardable(c4, 2)


# --------------------------------------------------------------------------- layout optimiser: L2 tail windows
def _window_rows(mode, dim, rows):
    return layout.LayoutOptimizer._window_rows(SimpleNamespace(mode=mode), SimpleNamespace(dim=dim, rep_count=rows))


def test_tail_window_rule():
    assert native.get_option("sgd_window_mb") == -1                   # automatic unless the option says otherwise
    # BASELINE configs[3]: 10M x 2-D, p + g = 160 MB -> two windows of <= 80 MB, whole multiples of 1024 rows
    rows = _window_rows("fit", 2, 10_000_000)
    assert rows % 1024 == 0 and 0 < rows < 10_000_000 and -(-10_000_000 // rows) == 2
    assert rows * 2 * 4 * 2 <= (layout.AUTO_WINDOW_MB << 20) + 1024 * 16
    assert _window_rows("fit", 16, 1_000_000) == 0                    # 64-byte rows: windows measured a loss (configs[2])
    assert _window_rows("fit", 16, 158_915) == 0                      # configs[1]: L2 resident anyway
    assert _window_rows("fit", 2, 1_000_000) == 0                     # 16 MB: fits the L2
    assert _window_rows("transform", 2, 10_000_000) == 0              # only p is touched: 80 MB <= 100 MB
    assert _window_rows("transform", 2, 30_000_000) > 0
    assert _window_rows("invert", 2, 100_000_000) == 0                # invert mode has its own kernel
    try:
        native.set_option("sgd_window_mb", 0)
        assert _window_rows("fit", 2, 10_000_000) == 0                # switched off
        native.set_option("sgd_window_mb", 32)
        rows32 = _window_rows("fit", 2, 10_000_000)
        assert rows32 % 1024 == 0 and -(-10_000_000 // rows32) == 5   # explicit size applies whatever the row width
        assert _window_rows("fit", 16, 1_000_000) > 0
    finally:
        native.set_option("sgd_window_mb", -1)


def test_push_tail_threshold_matches_design():
    # tables up to 32 MB use the push form of the epoch tail (configs[1]: 12.2 MB), larger ones the pull form
    c2_bytes = (158_915 + 31_783) * 16 * 4
    c3_bytes = 1_000_000 * 16 * 4
    assert c2_bytes <= layout.PUSH_TAIL_MAX_BYTES < c3_bytes


# --------------------------------------------------------------------------- row blocks across ranks
def test_row_blocks_tile_aligned_and_all_gather_padding():
    for n in (7, 128, 1000, 31_783, 158_915, 10_000_000):
        for w in (2, 3, 4, 8):
            per = D.block_size(n, w)
            assert per % 128 == 0 and per * w >= n
            for r in range(w):
                lo, hi = D.row_block(n, r, w)
                assert 0 <= lo <= hi <= n and hi - lo <= per
                assert lo == min(n, r * per)                           # rank r's block starts at r * per: the all-gather layout
    # item ranges (InfoNCE anchors, fallback rows): contiguous and covering
    for n in (0, 1, 5, 1000):
        for w in (1, 2, 8):
            rs = [D.item_range(n, r, w) for r in range(w)]
            assert rs[0][0] == 0 and rs[-1][1] == n and all(a[1] == b[0] for a, b in zip(rs, rs[1:]))


# --------------------------------------------------------------------------- bench.py: generators and sampler
@pytest.fixture(scope="module")
def bench():
    return _bench()


def test_generators_are_deterministic_paired_and_share_structure_across_seeds(bench):
    wl = bench.WORKLOADS["c2-tiny"]
    a, b = bench.make_data(wl, seed=0), bench.make_data(wl, seed=0)
    assert list(a) == ["texts", "images"]
    assert all(torch.equal(a[k], b[k]) for k in a)                     # both bench arms see the same inputs
    assert a["texts"].shape == (9932, 768) and a["images"].shape == (1986, 4096)
    assert a["texts"].dtype == torch.float32 and a["texts"].is_contiguous()
    assert float(a["texts"].abs().max()) <= 1.0                        # BERT pooler_output is tanh-bounded
    # caption c belongs to image c mod N_img and shares its cluster (c mod N_img mod 64): rows of one cluster are
    # closer to their own cluster mean than to any other, in both modalities, and the pairing follows
    n_img = 1986
    for name, x in a.items():
        cluster = (torch.arange(x.shape[0]) % n_img) % 64
        means = torch.stack([x[cluster == c].mean(dim=0) for c in range(64)])
        nearest = torch.cdist(x[:512], means).argmin(dim=1)
        assert (nearest == cluster[:512]).float().mean() > 0.95, name
    # held-out rows (another seed): different rows, same cluster centres
    q = bench.make_data(wl, seed=7, n_rows=1024)
    assert q["texts"].shape == (1024, 768) and q["images"].shape == (1024, 4096)
    assert not torch.equal(q["texts"][:8], a["texts"][:8])
    for name in a:
        c_fit = a[name][(torch.arange(a[name].shape[0]) % n_img) % 64 == 3].mean(dim=0)
        c_q = q[name][torch.arange(1024) % 64 == 3].mean(dim=0)        # n_rows: 1:1 pairs, cluster = row mod 64
        other = q[name][torch.arange(1024) % 64 == 4].mean(dim=0)
        assert torch.dist(c_fit, c_q) < 0.5 * torch.dist(c_fit, other), name


def test_blob_generators(bench):
    c1 = bench.make_data(bench.WORKLOADS["c1"], seed=0)["blobs"]
    assert c1.shape == (2000, 64)
    labels = torch.arange(2000) % 10                                   # configs[0]: 10 blobs, centres N(0, 5^2), unit noise
    means = torch.stack([c1[labels == c].mean(dim=0) for c in range(10)])
    assert (torch.cdist(c1, means).argmin(dim=1) == labels).all()
    within = (c1 - means[labels]).std()
    assert 0.9 < float(within) < 1.1


def test_clock_sampler_without_nvml_reports_instead_of_raising(bench, monkeypatch):
    # no GPU and no nvidia-smi here: the sampler must still hand bench.py a well-formed `clocks` object
    monkeypatch.setitem(sys.modules, "pynvml", None)                   # import pynvml -> ImportError
    monkeypatch.setenv("PATH", "/nonexistent")
    s = bench.ClockSampler(0)
    s.prepare()
    s.start()
    out = s.stop()
    assert set(out) >= {"sm_mhz", "sm_max_mhz", "reasons"} and out["sm_mhz"] is None and out["reasons"]


def test_clock_summary_reports_min_and_reasons(bench):
    s = bench.ClockSampler(0)
    nv = SimpleNamespace(nvmlClocksThrottleReasonHwSlowdown=0x8, nvmlClocksThrottleReasonHwThermalSlowdown=0x40,
                         nvmlClocksThrottleReasonSwThermalSlowdown=0x20, nvmlClocksThrottleReasonSwPowerCap=0x4)
    s._nv = nv
    s.thread = SimpleNamespace(join=lambda timeout=None: None)
    # idle samples (low power) are left out of the clock statistics; the power-capped sample sets the minimum
    s.samples = [(1965.0, 1965.0, 150.0, 0), (1965.0, 1965.0, 900.0, 0), (1500.0, 1965.0, 950.0, 0x4), (1965.0, 1965.0, 880.0, 0)]
    out = s.stop()
    assert out["sm_mhz"] == 1965.0 and out["sm_mhz_min"] == 1500.0 and out["sm_max_mhz"] == 1965.0
    assert out["reasons"] == ["sw_power_cap"] and out["power_w_max"] == 950.0 and out["samples"] == 4


def test_workload_table_matches_baseline_json(bench):
    import json
    base = json.load(open(os.path.join(ROOT, "BASELINE.json")))
    text = json.dumps(base).replace(",", "")                          # "31,783" / "158,915" in prose
    # the bench line is quoted on configs[1]; its shapes must be the ones BASELINE.json names
    for token in ("158915", "31783"):
        assert token in text
    c2 = bench.WORKLOADS["c2"]
    assert [(m[1], m[2]) for m in c2["mods"]] == [(158915, 768), (31783, 4096)]
    assert (c2["k"], c2["out_dim"], c2["epochs"]) == (15, 16, 600)
    assert bench.OPT == dict(min_dist=0.1, num_rep=8, lr=0.01, alpha=1.0, batch_size=256)   # reference main.py:15-21
    assert "fit" in base["metric"].lower()


@pytest.mark.timeout(600)
def test_reference_arm_prints_the_contract_line_on_cpu():
    """`bench.py --impl reference` (the arm the driver runs first on the GPU box) on the script's reduced self-test
    workload: one JSON line with the engine arm's metric / unit / direction, `impl`, `cpu_baseline` and a zero-copy `e2e`;
    no GPU involved (CUDA_VISIBLE_DEVICES is emptied by the script itself)."""
    import json
    import subprocess
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "c2-tiny",
                          "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=580, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.strip().split("\n") if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "umap_fit_seconds" and d["unit"] == "s"
    assert d["higher_is_better"] is False and d["n_gpus"] == 1 and d["steps"] == 1 and d["gpu_launches"] == 0
    assert d["value"] > 0 and abs(d["ms_per_step"] - 1e3 * d["value"]) < 1e-6 * d["ms_per_step"]
    cb = d["cpu_baseline"]
    assert cb["kind"] in ("reference", "port") and cb["cores"] >= 1 and cb["sample"] and cb["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"] == "c2-tiny" and d["data"] == "synthetic"
