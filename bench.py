#!/usr/bin/env python
"""bench.py -- UMAP fit benchmark of the B200 engine (and of the CPU reference arm).

    python bench.py --gpus N --steps K --warmup W            # this engine, one rank per GPU
    python bench.py --impl reference --gpus N --steps K ...  # CPU port of the reference path

One "step" is one complete fit of the workload (default: BASELINE.json configs[1], the
Flickr30k-shaped cross-modal fit: texts 158,915x768 + images 31,783x4096, k=15, 16-D, with the
reference CLI's optimiser defaults main.py:13-21 -- 600 epochs, num_rep 8, lr 0.01, alpha 1,
batch 256): exact kNN -> rho/sigma -> fuzzy union -> spectral init -> layout optimisation
(UMAPMixture.fit, /root/reference/impl/model.py:483-508).  Data is synthetic.

The JSON line printed by rank 0:
  metric/value   umap_fit_seconds with the inputs already resident in HBM (device timed, CUDA
                 events, max over ranks), lower is better;
  e2e            the same fit through the reference-facing API impl.util.train() with HOST
                 (pinned) inputs and the embeddings read back to the host, wall-clocked between
                 synchronisations;
  roofline       the dominant kernel (named by what actually launched: mmu_last_kernel) against the measured
                 peaks -- HBM copy rate, and for the L2-resident force kernel the random-row L2 roof measured live
                 by mmu_roof_random_rows; roofline_other = the second kernel (kNN contraction vs bf16 peak);
  quality        similarity_test / knn_test (k=1,5) of impl/validation.py (batched equivalents) on 100k held-out
                 queries (BASELINE.json configs[4]) and trustworthiness@15 on a 5k-row subsample per modality,
                 for the device and the host sample stream;
  stages         per-stage device times, kNN TFLOP/s, optimiser edge-updates/s and GB/s (per GPU);
  cpu_baseline   the CPU port (oracle/) on a bounded sample, extrapolated to the same metric.
"""
from __future__ import annotations

import argparse
import gc
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

if "--impl" in sys.argv[:-1] and sys.argv[sys.argv.index("--impl") + 1] == "reference" or "--impl=reference" in sys.argv:
    # the reference arm runs the reference's own CPU path: impl/model.py:10 picks cuda whenever it is visible
    os.environ["CUDA_VISIBLE_DEVICES"] = ""

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "multimodal-umap_b200")
for _p in (ROOT, PKG):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

WORKLOADS = {
    # BASELINE.json configs[1]
    "c2": dict(desc="Flickr30k-shaped cross-modal fit: texts 158915x768 + images 31783x4096, k=15, 16-D",
               mods=[("texts", 158915, 768, "bert"), ("images", 31783, 4096, "vae")],
               k=15, out_dim=16, epochs=600),
    # BASELINE.json configs[0] (the reference's own CPU-runnable case)
    "c1": dict(desc="2000x64 Gaussian blobs, k=15, 2-D, 200 epochs", mods=[("blobs", 2000, 64, "blobs")],
               k=15, out_dim=2, epochs=200),
    # BASELINE.json configs[2] and configs[3]: scale checks (parity-test / scaling cases, not the bench line)
    "c3": dict(desc="1M x 768 BERT-shaped single-modality fit, k=15, 16-D", mods=[("texts", 1000000, 768, "bert")],
               k=15, out_dim=16, epochs=600),
    "c4": dict(desc="10M x 128 points, k=30, 2-D, 500 epochs", mods=[("blobs", 10000000, 128, "blobs")],
               k=30, out_dim=2, epochs=500),
    # reduced C2 for functional checks of this script (NOT a benchmark configuration)
    "c2-tiny": dict(desc="C2 generators at 1/16 rows (script self-test only)",
                    mods=[("texts", 9932, 768, "bert"), ("images", 1986, 4096, "vae")], k=15, out_dim=16, epochs=50),
}
OPT = dict(min_dist=0.1, num_rep=8, lr=0.01, alpha=1.0, batch_size=256)      # reference main.py:15-21
SPINUP_S = 3.0                                                                # untimed load before the warm-up steps


# --------------------------------------------------------------------------- synthetic data
STRUCTURE_SEED = 20241018      # cluster centres and latent projections: the same for fit rows and held-out queries
LATENT_DIM = 8


def make_data(workload: dict, seed: int = 0, n_rows: int | None = None) -> dict:
    """SURVEY.md 8(d) generators on the CPU generator (identical for both arms).  The cluster centres and the
    projections of the shared latent come from a fixed structure seed, the rows from `seed`: a different `seed`
    gives held-out rows of the SAME distribution (BASELINE.json configs[4]).  Paired modalities ("bert" captions,
    "vae" image latents): caption c belongs to image c mod N_img, shares its cluster and an 8-D latent position that
    both modalities see through their own random projection -- without it a caption and its image would share
    nothing but the cluster and the reference's retrieval metric (validation.py:40-84) would measure chance.
    n_rows: that many rows in EVERY modality, paired 1:1 (the query sets)."""
    sgen = torch.Generator().manual_seed(STRUCTURE_SEED)
    gen = torch.Generator().manual_seed(seed)
    n_clusters = 64
    mods = [(nm, n_rows or n, d, kind) for (nm, n, d, kind) in workload["mods"]]
    n_img = min(n for (_, n, _, _) in mods)
    z = torch.randn((n_img, LATENT_DIM), generator=gen)
    out = {}
    for name, n, d, kind in mods:
        pair = torch.arange(n) % n_img                        # caption c <-> image c mod N_img
        cluster = pair % n_clusters
        if kind == "vae":       # SD-VAE latent_dist.mean scale: centre N(0,2^2) + latent + N(0,4^2) noise
            centres = torch.randn((n_clusters, d), generator=sgen) * 2.0
            proj = torch.randn((LATENT_DIM, d), generator=sgen) * 0.5
            x = centres[cluster] + z[pair] @ proj + torch.randn((n, d), generator=gen) * 4.0
        elif kind == "bert":    # BERT pooler_output: tanh-bounded
            centres = torch.randn((n_clusters, d), generator=sgen)
            proj = torch.randn((LATENT_DIM, d), generator=sgen) * 0.1
            x = torch.tanh(centres[cluster] + z[pair] @ proj + torch.randn((n, d), generator=gen) * 0.5)
        else:                   # C1/C4: centres N(0,5^2), points = centre + N(0,1)
            nc = 10 if n <= 100000 else 1000
            centres = torch.randn((nc, d), generator=sgen) * 5.0
            x = centres[torch.arange(n) % nc] + torch.randn((n, d), generator=gen)
        out[name] = x.contiguous()
    return out


# --------------------------------------------------------------------------- clocks sampler
class ClockSampler:
    """SM clock, power and throttle reasons DURING the timed region.  Uses NVML in a background thread
    (a polling nvidia-smi process takes driver locks for milliseconds per query and measurably slows the
    host-latency-bound stages it is supposed to observe); falls back to nvidia-smi without pynvml."""
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int = 0, period_s: float = 0.025):
        self.gpu_index, self.period = gpu_index, period_s
        self.samples = []          # (sm_mhz, max_mhz, power_w, reasons_bitmask)
        self.thread = self.proc = self.path = None
        self.stop_flag = False

    def _nvml_loop(self, nv, handle):
        while not self.stop_flag:
            try:
                sm = nv.nvmlDeviceGetClockInfo(handle, nv.NVML_CLOCK_SM)
                mx = self._max_mhz
                pw = nv.nvmlDeviceGetPowerUsage(handle) / 1000.0
                rs = nv.nvmlDeviceGetCurrentClocksEventReasons(handle) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else nv.nvmlDeviceGetCurrentClocksThrottleReasons(handle)
                self.samples.append((float(sm), float(mx), float(pw), int(rs)))
            except Exception:
                pass
            time.sleep(self.period)

    def prepare(self):
        """NVML initialisation, BEFORE the warm-up steps: nvmlInit takes driver-wide locks for tens to hundreds of
        milliseconds, and run inside the timed region (as start() used to) it stalled the launches of the first
        timed step's kNN stage -- the sampler must not perturb what it observes."""
        try:
            import pynvml as nv
            nv.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(visible.split(",")[self.gpu_index]) if visible and visible.split(",")[0].isdigit() else self.gpu_index
            self._handle = nv.nvmlDeviceGetHandleByIndex(phys)
            self._max_mhz = float(nv.nvmlDeviceGetMaxClockInfo(self._handle, nv.NVML_CLOCK_SM))
            self._nv = nv
        except Exception:
            self._nv = None

    def start(self):
        if getattr(self, "_nv", "unset") == "unset":
            self.prepare()
        if self._nv is not None:
            import threading
            self.thread = threading.Thread(target=self._nvml_loop, args=(self._nv, self._handle), daemon=True)
            self.thread.start()
            return
        try:
            self.path = tempfile.mktemp(prefix="clocks_", suffix=".csv")
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "500",
                 "-i", str(self.gpu_index)], stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self) -> dict:
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        sm, mx, power, reasons = [], [], [], set()
        if self.thread is not None:
            self.stop_flag = True
            self.thread.join(timeout=2)
            nv = self._nv
            bits = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                    "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                    "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                    "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
            for s_, m_, p_, r_ in self.samples:
                sm.append(s_); mx.append(m_); power.append(p_)
                for nm in names:
                    if r_ & bits[nm]:
                        reasons.add(nm)
        elif self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.proc.kill()
            try:
                for line in open(self.path):
                    f = [t.strip() for t in line.split(",")]
                    if len(f) < 9:
                        continue
                    try:
                        sm.append(float(f[1])); mx.append(float(f[2])); power.append(float(f[3]))
                    except ValueError:
                        continue
                    for nm, val in zip(names, f[5:9]):
                        if val.lower().startswith("active"):
                            reasons.add(nm)
                os.unlink(self.path)
            except OSError:
                pass
        else:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no sampler available"]}
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        busy = [s_ for s_, p_ in zip(sm, power) if p_ > 0.5 * max(power)] or sm
        # sm_mhz_min: the tensor-bound kNN stage is ~40 ms of a ~220 ms step and is the one the power cap slows (it
        # starts right after the previous fit's optimiser has held the board near its limit); the median hides it
        return {"sm_mhz": statistics.median(busy), "sm_mhz_min": min(busy), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                "power_w_max": max(power), "samples": len(sm)}


# --------------------------------------------------------------------------- CPU port (oracle)
def cpu_fit_sample(data: dict, workload: dict, threads: int, knn_rows: int | None = None) -> dict:
    """Times the CPU port of the reference path (oracle/) on a bounded sample of `workload` and
    extrapolates to fit seconds.  The reference's own kNN (NN-descent, model.py:81-195) does not
    complete at C2 scale, so the kNN leg is the exhaustive restatement the GPU
    path is held bit-exact to, timed on `knn_rows` query rows per modality against the FULL
    database with all host threads and scaled by N/knn_rows; sigma/union run at full size on the
    sampled rows' statistics (random k-regular stand-in graph for the rows not searched); the
    optimiser leg times ONE epoch of the restated _train (model.py:396-481) at full size and
    scales by the epoch count.  Spectral init: scipy lobpcg on the stand-in graph."""
    from oracle import umap_oracle as orc
    import warnings
    import scipy.sparse as sp
    import scipy.sparse.linalg as spla

    warnings.filterwarnings("ignore")
    torch.set_num_threads(threads)
    k, d_out, epochs = workload["k"], workload["out_dim"], workload["epochs"]
    rng = np.random.default_rng(0)
    t_knn = t_knn_port = t_sigma = t_union = t_spec = 0.0
    flops_port = 0.0
    graphs, inits = [], []
    knn_rows = knn_rows or max(8, 2 * threads)
    flops = 0.0
    for name, x in data.items():
        xn = x.numpy()
        n = xn.shape[0]
        q = min(knn_rows, n)
        t0 = time.perf_counter()
        idx_s, dist_s = orc.knn_exact(xn[:q], xn, k, True, nthreads=threads)
        dt = time.perf_counter() - t0
        t_knn_port += dt * n / q
        flops_port += 2.0 * q * n * xn.shape[1] / dt
        # the same search as a blocked dense contraction on all host cores (torch.cdist + topk):
        # the fastest CPU statement of the kNN leg, used for the extrapolated fit time
        qc = min(n, 2048)
        t0 = time.perf_counter()
        torch.cdist(x[:qc], x, compute_mode="use_mm_for_euclid_dist").topk(k + 1, dim=1, largest=False)
        dt = time.perf_counter() - t0
        t_knn += dt * n / qc
        flops += 2.0 * qc * n * xn.shape[1] / dt
        # stand-in graph for the rows not searched: random neighbours, distances resampled from the searched rows
        idx = rng.integers(0, n, (n, k)).astype(np.int32)
        dist = np.sort(rng.choice(dist_s.reshape(-1), (n, k)), axis=1).astype(np.float32)
        idx[:q], dist[:q] = idx_s, dist_s
        t0 = time.perf_counter()
        sig = orc.sigmas_bisect(dist)
        w = orc.membership_weights(dist, sig)
        ci, cw = orc.coalesce_rows(idx, w)
        t_sigma += time.perf_counter() - t0
        rows = np.repeat(np.arange(n, dtype=np.int64), k)
        # de-duplicate the stand-in rows so the COO is coalesced like the reference's
        key = rows * n + ci.reshape(-1)
        _, first = np.unique(key, return_index=True)
        rows, cols, vals = rows[first], ci.reshape(-1)[first].astype(np.int64), cw.reshape(-1)[first]
        t0 = time.perf_counter()
        ur, uc, uv = orc.fuzzy_union(rows, cols, vals, n)
        t_union += time.perf_counter() - t0
        t0 = time.perf_counter()
        s = sp.coo_matrix((uv.astype(np.float32), (ur, uc)), shape=(n, n)).tocsr()
        deg = np.maximum(np.asarray(s.sum(axis=1)).ravel(), 1e-6)
        dm = sp.diags((deg ** -0.5).astype(np.float32))
        lap = sp.identity(n, dtype=np.float32) * np.float32(1.0 + 1e-6) - dm @ s @ dm
        x0 = rng.standard_normal((n, d_out + 1)).astype(np.float32)
        try:
            _, vecs = spla.lobpcg(lap, x0, largest=False, tol=3.5e-4, maxiter=40)
        except Exception:
            vecs = x0
        t_spec += time.perf_counter() - t0
        graphs.append((ur, uc, uv))
        inits.append(np.ascontiguousarray(vecs[:, 1:d_out + 1] if vecs.shape[1] > d_out else vecs[:, :d_out],
                                          dtype=np.float32))
    torch.manual_seed(0)
    t0 = time.perf_counter()
    kept_rec = []
    orc.train_oracle(inits, graphs, 1, OPT["num_rep"], OPT["lr"], OPT["alpha"], OPT["batch_size"], 1.577, 0.8951,
                     mode="fit", infonce=orc.infonce_grad_vec, record=kept_rec)
    t_epoch = time.perf_counter() - t0
    kept = sum(float(g[2].sum()) for g in graphs)          # E[kept] = sum of weights (Bernoulli(w), model.py:432)
    total = t_knn + t_sigma + t_union + t_spec + t_epoch * epochs
    return {
        "fit_seconds": total,
        "parts_s": {"knn_extrapolated": t_knn, "knn_exact_port_extrapolated": t_knn_port, "sigma": t_sigma, "union": t_union, "spectral_40it": t_spec,
                    "one_epoch": t_epoch, "optimise_extrapolated": t_epoch * epochs},
        "knn_tflops": flops / len(data) / 1e12,
        "knn_exact_port_tflops": flops_port / len(data) / 1e12,
        "edge_updates_per_s": kept * (1 + OPT["num_rep"]) / t_epoch,
        "sample": (f"kNN as torch.cdist+topk on 2048 query rows/modality vs full db (x N/2048; the bit-exact C port "
                   f"on {knn_rows} rows is reported beside it); sigma+union+scipy-lobpcg(40 it) "
                   f"full size on a stand-in graph; 1 of {epochs} optimiser epochs (x{epochs})"),
    }


# --------------------------------------------------------------------------- this engine
def measure_roofs(dev, tables: list) -> dict:
    """Live roofs of the random-access force kernel (mmu_roof_random_rows, csrc/roofs.cu): random row gathers + random
    vector reds with the force kernel's own access shape on tables of the workload's own sizes, CUDA events, best of 3
    after a warm-up.  `tables` = [(label, n_rows, row_floats)]."""
    from umap_b200.native import check, lib, ptr, stream
    out = {}
    for label, n_rows, row_floats in tables:
        rf = row_floats if row_floats in (2, 4, 16, 64) else 16
        table = torch.randn((n_rows, rf), device=dev)
        accum = torch.zeros((n_rows, rf), device=dev)
        sink = torch.zeros(1, device=dev)
        touched = int(min(4e8, max(2e7, 64 * n_rows)))
        best = None
        for it in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            check(lib().mmu_roof_random_rows(ptr(table), ptr(accum), n_rows, rf, touched, 11 + it, 1, 1, ptr(sink), stream()),
                  "mmu_roof_random_rows")
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            if it and (best is None or ms < best):
                best = ms
        nbytes = touched * rf * 4 * 2
        out[label] = {"rows": n_rows, "row_bytes": rf * 4, "table_mb": round(2 * n_rows * rf * 4 / 2 ** 20, 1),
                      "gbs": nbytes / (best * 1e-3) / 1e9, "row_accesses_per_s": 2 * touched / (best * 1e-3)}
    return out


def quality_block(model, util_mod, cfg, workload, data, dev) -> dict:
    """The reference's end-of-run metrics (main.py:60-61 -> impl/validation.py:7-84) through their batched equivalents
    (umap_b200.metrics; tests/test_gpu_harness.py holds them equal to validation.py itself) on 100k held-out queries
    per modality (BASELINE.json configs[4]), plus sklearn trustworthiness@15 of the fitted embedding on a 5k-row
    subsample per modality."""
    from sklearn.manifold import trustworthiness
    from umap_b200 import metrics
    out = {}
    names = [m[0] for m in workload["mods"]]
    g = torch.Generator().manual_seed(99)
    for i, name in enumerate(names):
        n = data[name].shape[0]
        idx = torch.randperm(n, generator=g)[: min(5000, n)]
        x = data[name][idx].numpy()
        y = model.embeds[i].detach()[idx.to(model.embeds[i].device)].cpu().numpy()
        out[f"trustworthiness15_{name}"] = float(trustworthiness(x, y, n_neighbors=15))
    if len(names) > 1:
        nq = 100000
        qd = {k: v.to(dev) for k, v in make_data(workload, seed=7, n_rows=nq).items()}
        torch.manual_seed(4321)
        t0 = time.perf_counter()
        out["similarity_test"] = metrics.similarity_test(model, util_mod.embed, qd, cfg)
        accs = metrics.knn_test_multi(model, util_mod.embed, qd, cfg, ks=(1, 5))
        torch.cuda.synchronize()
        out["knn_test_k1"], out["knn_test_k5"] = accs[1], accs[5]
        out["queries_per_modality"] = nq
        out["chance_knn_k1"] = 1.0 / nq
        out["metrics_wall_s"] = time.perf_counter() - t0
    return out


def c3_stage(util_mod, world, dev) -> dict:
    """BASELINE.json configs[2] (1M x 768 BERT-shaped, k=15, 16-D, 600 epochs) as ONE fit inside every bench run, so that
    the driver's 1/2/4/8-GPU runs carry a number on a configuration meant for scaling.  Rows are generated on the
    device (same seed on every rank -> identical replicas), inputs resident, device timed, max over ranks."""
    import torch.distributed as dist
    from umap_b200 import profiler
    wl = WORKLOADS["c3"]
    n, d = wl["mods"][0][1], wl["mods"][0][2]
    g = torch.Generator(device=dev).manual_seed(5)
    centres = torch.randn((64, d), generator=g, device=dev)
    x = torch.empty((n, d), device=dev)
    for lo in range(0, n, 100000):                            # chunked: keeps the temporaries small
        hi = min(n, lo + 100000)
        x[lo:hi] = torch.tanh(centres[torch.arange(lo, hi, device=dev) % 64] + 0.5 * torch.randn((hi - lo, d), generator=g, device=dev))
    cfg = util_mod.Config(k_neighbors=wl["k"], out_dim=wl["out_dim"], min_dist=OPT["min_dist"], train_epochs=wl["epochs"],
                          num_rep=OPT["num_rep"], lr=OPT["lr"], alpha=OPT["alpha"], batch_size=OPT["batch_size"], test_epochs=120)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    profiler.enable(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.manual_seed(77)
    e0.record()
    model = util_mod.train({"texts": x}, cfg)
    e1.record()
    torch.cuda.synchronize()
    st = profiler.summarize(profiler.collect())
    profiler.enable(0)
    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    knn = st.get("knn", {"ms": 0.0, "flops": 0.0})
    out = {"desc": wl["desc"], "epochs": wl["epochs"], "fit_s": float(t[0]) / 1e3,
           "stage_ms_rank0": {k: round(v["ms"], 2) for k, v in st.items() if k != "edge_forces"},
           "knn_tflops_per_gpu": (knn["flops"] / (knn["ms"] * 1e-3) / 1e12) if knn["ms"] else None,
           "finite": bool(torch.isfinite(model.embeds[0]).all().item())}
    del model, x
    torch.cuda.empty_cache()
    return out


def load_profile_facts() -> dict:
    """ncu-derived per-launch figures (DRAM traffic, unit utilisations) live in a tracked file written from the
    .ncu-rep captures by scripts/ncu_facts.py -- never as constants in this script."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "r02_ncu_facts.json")))
    except (OSError, ValueError):
        return {}


def run_b200(args, workload, data):
    import torch.distributed as dist
    from umap_b200 import native, profiler
    import importlib
    model_mod = importlib.import_module("impl.model")
    util_mod = importlib.import_module("impl.util")

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 engine has no CPU fallback (use --impl reference)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if args.gpus != world:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run")
    dev = torch.device("cuda", local)
    cfg = util_mod.Config(k_neighbors=workload["k"], out_dim=workload["out_dim"], min_dist=OPT["min_dist"],
                          train_epochs=workload["epochs"], num_rep=OPT["num_rep"], lr=OPT["lr"], alpha=OPT["alpha"],
                          batch_size=OPT["batch_size"], test_epochs=120)
    host = {k: v.pin_memory() for k, v in data.items()}
    resident = {k: v.to(dev) for k, v in host.items()}
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def fit_resident():
        torch.manual_seed(1234)
        return util_mod.train(resident, cfg)

    h2d_rank = [0]

    def fit_e2e():
        torch.manual_seed(1234)
        model = util_mod.train(host, cfg)                   # H2D copies happen inside (model.py:496,634)
        h2d_rank[0] = int(model.last_h2d_bytes)             # what this rank pulled over its host link
        return [e.detach().cpu() for e in model.embeds]     # D2H read of the result

    # Spin-up, then the W warm-up steps: a fresh box needs a few seconds of load before clocks and power state
    # settle (the first process on a box measured its kNN stage up to 2x slower during its first ~2 s); untimed
    # fits until rank 0 has seen SPINUP_S seconds of them, the same number on every rank
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.prepare()
    # The warm-up holds the previous fit's model while the next one runs, exactly as the timed loop does: with the
    # result discarded instead, the caching allocator met a new lifetime pattern in timed steps 2-3 and its
    # cudaMalloc/cudaFree calls landed inside those steps' kNN stage (single searches of 108 ms instead of 31 ms).
    model = None
    spin_t0 = time.perf_counter()
    while not args.quick:
        model = fit_resident()
        torch.cuda.synchronize()
        go = torch.tensor([1 if time.perf_counter() - spin_t0 < SPINUP_S else 0], dtype=torch.int32, device=dev)
        if world > 1:
            dist.broadcast(go, src=0)
        if int(go.item()) == 0:
            break
    for _ in range(args.warmup):
        model = fit_resident()
    # Python's cyclic collector stays ON, but what the process holds after warm-up (torch, scipy, numpy: ~10^6 tracked
    # objects) is moved to the permanent generation: a full collection that walks all of it takes ~0.1 s and, landing
    # inside a 0.22 s step, was the other source of single-step outliers (scripts/e2e_probe.py; timeit switches the
    # collector off altogether for the same reason)
    gc.collect()
    gc.freeze()
    barrier()
    if rank == 0:
        sampler.start()
    launches0 = native.lib().mmu_launch_count()
    profiler.enable(1)            # coarse stages + the force kernel; the small kernels get their own pass below
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        model = fit_resident()
    ev1.record()
    barrier()
    total_ms = ev0.elapsed_time(ev1)
    launches = native.lib().mmu_launch_count() - launches0
    stage_rows = profiler.collect()
    stages = profiler.summarize(stage_rows)
    knn_each = [round(ms, 2) for (nm, ms, _) in stage_rows if nm == "knn"]      # per search call, in program order
    force_kernel = native.last_kernel("edge_forces")
    knn_kernel = native.last_kernel("knn_candidates")
    tail_kernel = native.last_kernel("epoch_tail")
    if os.environ.get("MMUMAP_BENCH_DEBUG") == "1":
        print(f"[rank {rank}] stages ms/step: " + ", ".join(f"{k}={v['ms'] / args.steps:.1f}" for k, v in stages.items()),
              file=sys.stderr, flush=True)
    # one extra, untimed, fully instrumented fit for the per-kernel breakdown of an epoch
    fine = {}
    if not args.quick:
        profiler.enable(2)
        fit_resident()
        fine = profiler.summarize(profiler.collect())
    profiler.enable(0)
    opt = model.last_optimizer
    kept = opt.kept_last_epoch()
    nnz = [int(m.graph.nnz) for m in opt.mods]
    rows = [int(m.count) for m in opt.mods]
    exchange = "none (1 GPU)" if world == 1 else ("peer" if opt.peer is not None else "nccl")
    knn_dist = "single GPU" if world == 1 else os.environ.get("MMUMAP_KNN_DIST", model_mod.DEFAULT_KNN_DIST)

    # end to end through the reference-facing API, host buffers in, host result out
    e2e_s = float("nan")
    e2e_each = []
    if not args.quick:
        for _ in range(max(1, args.warmup)):                # the same W warm-up steps as the device-timed loop
            fit_e2e()
        gc.collect()
        gc.freeze()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            t1 = time.perf_counter()
            fit_e2e()
            e2e_each.append(round(time.perf_counter() - t1, 4))
        barrier()
        e2e_s = (time.perf_counter() - t0) / args.steps
    clocks = sampler.stop() if rank == 0 else None

    # BASELINE.json configs[4]: cross-modal transform of 100k held-out queries per modality against the
    # fitted model (util.embed -> UMAPMixture.transform, model.py:527-555), 120 test epochs; reported
    # beside the headline, not part of it
    transform = None
    if args.workload == "c2" and not args.no_transform:
        nq = 100000
        qd = make_data(workload, seed=7, n_rows=nq)
        qdev = [v.to(dev) for v in qd.values()]
        util_mod.embed(model, qdev, list(range(len(qdev))), cfg)       # untimed first call (kernel loads, allocations)
        barrier()
        tr0, tr1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        tr0.record()
        out = util_mod.embed(model, qdev, list(range(len(qdev))), cfg)
        tr1.record()
        torch.cuda.synchronize()
        tt = torch.tensor([tr0.elapsed_time(tr1)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        transform = {"queries_per_modality": nq, "test_epochs": cfg.test_epochs, "seconds": float(tt[0]) / 1e3,
                     "finite": bool(all(torch.isfinite(o).all().item() for o in out))}
        del qdev, out

    # quality of what was just timed (device stream) and of the same fit under the reference's own host sample
    # stream (one GPU only: 600 epochs of CPU-generator draws take about two minutes)
    quality = None
    if args.quality != "off" and not args.quick:
        quality = {"device_stream": quality_block(model, util_mod, cfg, workload, data, dev) if (world == 1 or args.workload == "c2") else None}
        if args.quality == "both" and world == 1:
            os.environ["MMUMAP_SAMPLE_STREAM"] = "host"
            try:
                t0 = time.perf_counter()
                torch.manual_seed(1234)
                mh = util_mod.train(resident, cfg)
                torch.cuda.synchronize()
                host_fit_s = time.perf_counter() - t0
                quality["host_stream"] = quality_block(mh, util_mod, cfg, workload, data, dev)
                quality["host_stream"]["fit_wall_s"] = host_fit_s
                del mh
            finally:
                os.environ.pop("MMUMAP_SAMPLE_STREAM", None)

    c3 = None
    if args.workload == "c2" and not args.quick and not args.no_c3:
        del resident
        torch.cuda.empty_cache()
        c3 = c3_stage(util_mod, world, dev)

    roofs = None
    if rank == 0 and not args.quick:
        d_out = workload["out_dim"]
        roofs = measure_roofs(dev, [(nm, n, d_out) for (nm, n, _, _) in workload["mods"]])

    t = torch.tensor([total_ms, e2e_s], dtype=torch.float64, device=dev)
    hb = torch.tensor([h2d_rank[0]], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(hb, op=dist.ReduceOp.SUM)
    total_ms, e2e_s = float(t[0]), float(t[1])
    h2d_total = int(hb[0])
    if rank != 0:
        return None
    ms_per_step = total_ms / args.steps
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except OSError:
        pass
    facts = load_profile_facts()
    steps = args.steps
    st = {k: dict(v, ms=v["ms"] / steps) for k, v in stages.items()}
    knn = stages.get("knn", {"ms": 0.0, "calls": 1, "flops": 0.0})
    knn_tflops = knn.get("flops", 0.0) / (knn["ms"] * 1e-3) / 1e12 if knn["ms"] > 0 else 0.0
    # kernels timed inside a seconds-long step: sustained tensor peak, measured HBM copy rate
    tensor_peak = peaks.get("bf16_tflops_sustained", 1400.0)
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    peak_src = "measured (MEASURED_PEAKS.json)" if peaks else "fallback (B200_PROFILING.md)"
    d = workload["out_dim"]
    epochs = workload["epochs"]
    opt_ms = stages.get("optimise", {"ms": 0.0})["ms"] / steps
    # SURVEY.md 8(d) canonical per-epoch bytes, fit mode, device RNG, int32 COO -- whole job; every rank holds 1/world of
    # the edge work and all of the Adam work divided by world on the peer path
    bytes_epoch = sum(12 * z for z in nnz) + kept * (2 + OPT["num_rep"]) * d * 4 * 2 + sum(28 * r * d for r in rows)
    edge_updates = kept * (1 + OPT["num_rep"])
    # N > 1 replays whole epochs from CUDA graphs (no per-kernel events inside): the force kernel's launch time then
    # comes from the instrumented pass that follows the timed region (same kernels, launched one by one)
    forces_from = "timed region"
    forces = stages.get("edge_forces")
    forces_scale = 1.0
    if not forces or not forces.get("calls"):
        forces, forces_from = fine.get("edge_forces"), "instrumented pass after the timed region (the timed region replays CUDA graphs)"
        forces_scale = float(steps)            # one instrumented fit against `steps` timed fits
    if not forces or not forces.get("calls"):
        forces = {"ms": 0.0, "calls": 0, "bytes": 0.0, "edge_updates": 0.0}
    forces_gbs = forces["bytes"] / (forces["ms"] * 1e-3) / 1e9 if forces["ms"] > 0 else 0.0
    kf = facts.get("knn_candidates", {})
    ff = facts.get("edge_forces", {})
    from umap_b200 import knn_tc as _kt
    pruning = _kt.last_stats.get("pruning") if _kt.last_stats.get("pruned") else None
    roof_knn = {"kernel": f"knn_tc: tc_prep + {knn_kernel or 'knn_tc_candidates_kernel'} [tcgen05] + knn_tc_rescore_kernel",
                "bound": "tensor", "achieved": knn_tflops, "peak": tensor_peak, "unit": "TFLOP/s",
                "frac": knn_tflops / tensor_peak, "traffic": kf.get("dram_bytes_per_launch"), "peak_source": peak_src,
                # cluster-pruned search (knn_pruned.py): `achieved` is the ALGORITHMIC 2QND of the exhaustive search per second;
                # the tensor cores evaluate only visited_tile_fraction of it (x3 for the split-fp16 operands), so frac > 1
                # measures the pruning, not the tensor pipe
                "pruned_search": pruning,
                "traffic_note": "DRAM read + write bytes per CANDIDATES-KERNEL launch from the ncu capture (a texts search walks "
                                "the database in ~48 MB windows: ten launches); `achieved` is per kNN stage call",
                "share_of_step": knn["ms"] / total_ms if total_ms else None,
                "ms_per_launch": knn["ms"] / max(knn["calls"], 1), "ncu": kf or None}
    # the force kernel: algorithmic bytes per launch / measured launch time, against the HBM copy peak as
    # the contract asks -- and, because the tables of this workload are L2 resident, against the random-row roof
    # measured live on tables of the same size (the bound that actually applies)
    l2_roof = None
    if roofs:
        w_sum = sum(rows)
        peak_l2 = sum(roofs[nm]["gbs"] * r for (nm, _, _, _), r in zip(workload["mods"], rows)) / max(w_sum, 1)
        l2_roof = {"achieved_gbs": forces_gbs, "peak_gbs": peak_l2, "frac": forces_gbs / peak_l2 if peak_l2 else None,
                   "how": "mmu_roof_random_rows: random row gathers + vector reds, same access shape and table sizes, "
                          "no arithmetic; weighted by rows per modality", "per_table": roofs}
    table_mb = sum(2 * r * d * 4 for r in rows) / 2 ** 20
    roof_sgd = {"kernel": force_kernel or "edge_forces", "bound": "hbm", "achieved": forces_gbs, "peak": hbm_peak,
                "unit": "GB/s", "frac": forces_gbs / hbm_peak if hbm_peak else None,
                "traffic": ff.get("dram_bytes_per_launch"),
                "effective_bound": "l2 (p and g tables %.0f MB, resident in the 126 MB L2)" % table_mb if table_mb < 100 else "hbm",
                "l2_roof": l2_roof, "peak_source": peak_src,
                "share_of_step": forces["ms"] * forces_scale / total_ms if total_ms else None, "timed_in": forces_from,
                "ms_per_launch": forces["ms"] / max(forces["calls"], 1),
                "edge_updates_per_s": forces["edge_updates"] / (forces["ms"] * 1e-3) if forces["ms"] > 0 else None,
                "ncu": ff or None}
    dominant, other = (roof_sgd, roof_knn) if forces["ms"] * forces_scale >= knn["ms"] else (roof_knn, roof_sgd)
    sgd_gbs_job = bytes_epoch * epochs / (opt_ms * 1e-3) / 1e9 if opt_ms else None
    line = {
        "metric": "umap_fit_seconds", "value": ms_per_step / 1e3, "unit": "s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": False,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "desc": workload["desc"], "epochs": epochs, **OPT,
                   "knn_method": os.environ.get("MMUMAP_KNN", "default"),
                   "sample_stream": os.environ.get("MMUMAP_SAMPLE_STREAM", "device"),
                   "exchange": exchange, "knn_dist": knn_dist, "epoch_tail_kernel": tail_kernel or None,
                   "l2": "inputs (1.0 GB) exceed the 126 MB L2; every step re-reads them from HBM",
                   "spinup_s": 0.0 if args.quick else SPINUP_S,
                   "python_gc": "enabled; gc.collect() + gc.freeze() after the warm-up steps",
                   "parallelism": f"kNN query-row blocks x{world}, optimiser edge shards x{world}" if world > 1 else "1 GPU"},
        "e2e": {"value": None if e2e_s != e2e_s else e2e_s, "unit": "s",
                "h2d_bytes_per_step": h2d_total,
                "h2d_bytes_per_rank": h2d_rank[0],
                "d2h_bytes_per_step": int(sum(r * d * 4 for r in rows)),
                "seconds_each_step_rank0": e2e_each},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "roofline": dominant,
        "roofline_other": other,
        "quality": quality,
        "stages": {
            "ms": {k: round(v["ms"], 3) for k, v in st.items()},
            "epoch_kernels_us_per_launch": {k: round(v["ms"] / max(v["calls"], 1) * 1e3, 1) for k, v in fine.items()
                                            if k in ("edge_sample", "edge_forces", "infonce", "adam", "epoch_tail")},
            "knn_tflops": knn_tflops,
            "knn_ms_each_call": knn_each,
            "sgd_edge_updates_per_s": edge_updates * epochs / (opt_ms * 1e-3) if opt_ms else None,
            "sgd_gbs_whole_job": sgd_gbs_job,
            "sgd_gbs_per_gpu": sgd_gbs_job / world if sgd_gbs_job else None,
            "sgd_hbm_frac_per_gpu": sgd_gbs_job / world / hbm_peak if sgd_gbs_job else None,
            "kept_edges_last_epoch": kept, "union_nnz": nnz,
            "transform_100k": transform,
            "c3": c3,
        },
    }
    if world == 1 and not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        cb = cpu_fit_sample(data, workload, threads)
        line["cpu_baseline"] = {"value": cb["fit_seconds"], "unit": "s", "cores": threads, "kind": "port",
                                "sample": cb["sample"], "parts_s": cb["parts_s"], "knn_tflops": cb["knn_tflops"],
                                "edge_updates_per_s": cb["edge_updates_per_s"]}
    return line


# --------------------------------------------------------------------------- reference arm
REF_STAGE = os.path.join(ROOT, "baseline", "_ref")


def _reference_model_module():
    """The reference's own impl/model.py (unmodified, staged under baseline/_ref by __graft_entry__.build()),
    imported as refimpl.model.  CUDA_VISIBLE_DEVICES="" was set before torch was imported, so its
    `device` (model.py:10) is the CPU."""
    import importlib
    import types
    if not os.path.isfile(os.path.join(REF_STAGE, "impl", "model.py")):
        return None
    if "refimpl" not in sys.modules:
        pkg = types.ModuleType("refimpl")
        pkg.__path__ = [os.path.join(REF_STAGE, "impl")]
        sys.modules["refimpl"] = pkg
    return importlib.import_module("refimpl.model")


class ReferenceSampler:
    """Times the REFERENCE's own code on the host cores, on bounded samples of the workload:
      graph legs (once per run; 2 % of the total): UMAPEncoder.fuzzy_knn_graph (model.py:63-209) + the union
        (model.py:271) + embed_all (model.py:211-234) on a `sub`-row subsample of each modality at full width, scaled by
        N / sub (the reference's graph time is linear in N at fixed D: SURVEY.md section 6; the full 158,915-row text
        graph cannot be built by the reference at all);
      optimiser (every step; 98 % of the total): ONE real epoch of UMAPMixture._train (model.py:396-481) at FULL
        size -- both modalities, all rows, InfoNCE included -- on stand-in graphs of the right shape (k-regular
        random pattern, weights resampled from the subsample's graph, symmetrised by the reference's own union
        expression), x epochs."""

    def __init__(self, ref, data, workload, threads, sub=2048):
        self.ref, self.data, self.workload, self.threads, self.sub = ref, data, workload, threads, sub
        torch.set_num_threads(threads)
        k, d_out = workload["k"], workload["out_dim"]
        self.model = ref.UMAPMixture(k_neighbors=k, out_dim=d_out, min_dist=OPT["min_dist"], num_encoders=len(data))
        self.graph_legs = None
        self.graphs = None
        self.embeds = None

    def _graph_legs(self):
        import warnings
        warnings.filterwarnings("ignore")
        k = self.workload["k"]
        legs = {"fuzzy_knn_graph": 0.0, "union": 0.0, "embed_all": 0.0}
        graphs, embeds = [], []
        g = torch.Generator().manual_seed(3)
        for i, (name, x) in enumerate(self.data.items()):
            n = x.shape[0]
            sub = min(self.sub, n)
            xs = x[torch.randperm(n, generator=g)[:sub]].contiguous()
            enc = self.model.encoders[i]
            t0 = time.perf_counter()
            gr = enc.fuzzy_knn_graph(xs, mode="fit")                        # model.py:63-209
            t1 = time.perf_counter()
            sym = (gr + gr.T - gr * gr.T).coalesce()                        # model.py:271
            t2 = time.perf_counter()
            enc.embed_all(sym)                                              # model.py:211-234
            t3 = time.perf_counter()
            scale = n / sub
            legs["fuzzy_knn_graph"] += (t1 - t0) * scale
            legs["union"] += (t2 - t1) * scale
            legs["embed_all"] += (t3 - t2) * scale
            # full-size stand-in graph for the optimiser leg: k random neighbours per row, weights resampled
            cols = torch.randint(0, n, (n, k), generator=g)
            rows = torch.arange(n).repeat_interleave(k)
            wsrc = gr.values()
            w = wsrc[torch.randint(0, wsrc.numel(), (n * k,), generator=g)]
            full = torch.sparse_coo_tensor(torch.stack([rows, cols.reshape(-1)]), w, (n, n)).coalesce()
            graphs.append((full + full.T - full * full.T).coalesce())
            embeds.append(torch.randn((n, self.workload["out_dim"]), generator=g) * (1.0 / n ** 0.5))
            enc.sigmas = torch.ones(n)
            enc.rhos = torch.zeros(n)
        self.graph_legs, self.graphs, self.embeds = legs, graphs, embeds

    def sample(self) -> dict:
        if self.graph_legs is None:
            self._graph_legs()
        epochs = self.workload["epochs"]
        torch.manual_seed(0)
        t0 = time.perf_counter()
        self.model._train(self.embeds, self.graphs, 1, OPT["num_rep"], OPT["lr"], OPT["alpha"], OPT["batch_size"],
                          mode="fit")                                       # model.py:396-481, one real epoch
        t_epoch = time.perf_counter() - t0
        kept = sum(float(gph.values().sum()) for gph in self.graphs)
        legs = self.graph_legs
        total = sum(legs.values()) + t_epoch * epochs
        return {"fit_seconds": total,
                "parts_s": {**{k: round(v, 2) for k, v in legs.items()}, "one_epoch": t_epoch,
                            "optimise_extrapolated": t_epoch * epochs},
                "edge_updates_per_s": kept * (1 + OPT["num_rep"]) / t_epoch,
                "sample": (f"the reference's own impl/model.py on CPU: 1 real _train epoch at full size per step (x{epochs}); "
                           f"fuzzy_knn_graph + union + embed_all on a {self.sub}-row subsample per modality at full width, "
                           f"timed once per run and scaled by N/{self.sub}")}


def reference_c1_full(ref, threads) -> dict:
    """BASELINE.json configs[0] in full through the reference's own fit (its one fully CPU-runnable case)."""
    torch.set_num_threads(threads)
    data = make_data(WORKLOADS["c1"])
    wl = WORKLOADS["c1"]
    torch.manual_seed(0)
    t0 = time.perf_counter()
    m = ref.UMAPMixture(k_neighbors=wl["k"], out_dim=wl["out_dim"], min_dist=OPT["min_dist"], num_encoders=1)
    m.fit([data["blobs"]], epochs=wl["epochs"], num_rep=OPT["num_rep"], lr=OPT["lr"], alpha=OPT["alpha"],
          batch_size=OPT["batch_size"])
    return {"workload": "c1", "desc": wl["desc"], "fit_s": time.perf_counter() - t0, "cores": threads}


def run_reference(args, workload, data):
    """Times the reference's own CPU implementation of the path on the box's host cores (all threads): the staged,
    unmodified impl/model.py when baseline/_ref is present (kind "reference"), otherwise the CPU port in oracle/
    (kind "port").  The reference's graph stage does not complete at C2 scale, so every step is a
    bounded sample -- see ReferenceSampler."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return None
    threads = os.cpu_count() or 1
    ref = None if os.environ.get("MMUMAP_REFERENCE") == "port" else _reference_model_module()
    vals = []
    extra = {}
    t0 = time.perf_counter()
    if ref is not None:
        import contextlib
        import io
        cb = None
        with contextlib.redirect_stderr(io.StringIO()):               # tqdm bars
            rs = ReferenceSampler(ref, data, workload, threads)
            for _ in range(min(args.warmup, 1)):
                rs.sample()                                           # absorbs the graph legs and the first-epoch overhead
            for _ in range(args.steps):
                cb = rs.sample()
                vals.append(cb["fit_seconds"])
            if not args.quick and args.workload == "c2":
                extra["c1_full_fit"] = reference_c1_full(ref, threads)
        kind = "reference"
    else:
        for _ in range(min(args.warmup, 1)):
            cpu_fit_sample(data, workload, threads, knn_rows=max(2, threads // 4))
        cb = None
        for _ in range(args.steps):
            cb = cpu_fit_sample(data, workload, threads)
            vals.append(cb["fit_seconds"])
        kind = "port"
    wall = time.perf_counter() - t0
    v = statistics.mean(vals)
    return {
        "impl": "reference", "metric": "umap_fit_seconds", "value": v, "unit": "s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": min(args.warmup, 1), "ms_per_step": v * 1e3, "higher_is_better": False,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "desc": workload["desc"], "epochs": workload["epochs"], **OPT},
        "cpu_baseline": {"value": v, "unit": "s", "cores": threads, "kind": kind, "sample": cb["sample"],
                         "parts_s": cb["parts_s"], "edge_updates_per_s": cb["edge_updates_per_s"],
                         "spread_s": [min(vals), max(vals)], "run_wall_s": wall, **extra},
        "e2e": {"value": v, "unit": "s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--workload", choices=sorted(WORKLOADS), default="c2")
    ap.add_argument("--epochs", type=int, default=None, help="override the workload's epoch count (not a benchmark configuration)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-transform", action="store_true")
    ap.add_argument("--quick", action="store_true", help="scale checks only: skip the end-to-end and instrumented passes")
    ap.add_argument("--quality", choices=["off", "device", "both"], default="both",
                    help="quality metrics after the timed region: device sample stream, or device and host stream (1 GPU)")
    ap.add_argument("--no-c3", action="store_true", help="skip the embedded C3 (1M x 768) fit")
    args = ap.parse_args()
    workload = dict(WORKLOADS[args.workload])
    if args.epochs is not None:
        workload["epochs"] = args.epochs
    if args.impl == "reference" and int(os.environ.get("RANK", "0")) != 0:
        return
    data = make_data(workload)
    line = run_reference(args, workload, data) if args.impl == "reference" else run_b200(args, workload, data)
    if line is not None:
        print(json.dumps(line), flush=True)
    if args.impl == "b200" and int(os.environ.get("WORLD_SIZE", "1")) > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
