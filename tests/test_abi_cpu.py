"""CPU-side checks of the drop-in boundary: the shared library loads, exports exactly the
symbols include/mmumap.h declares, validates arguments, and fails loudly without a GPU."""
import ctypes
import os
import re
import subprocess

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "mmumap.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mmu_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from umap_b200 import native
    lib = native.lib()
    declared = _declared()
    assert len(declared) >= 15
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/mmumap.h but not exported"
    assert sorted(native.EXPORTED_SYMBOLS) == declared, "ctypes table out of sync with include/mmumap.h"
    out = subprocess.check_output(["nm", "-D", "--defined-only", native.LIB_PATH], text=True)
    exported = sorted(set(re.findall(r" T (mmu_[a-z0-9_]+)", out)))
    assert exported == declared, "library exports symbols the header does not declare (or vice versa)"
    assert lib.mmu_abi_version() == 1


def test_library_is_sm100a_only():
    from umap_b200 import native
    out = subprocess.check_output(["cuobjdump", "--list-elf", native.LIB_PATH], text=True)
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_argument_validation_without_gpu():
    from umap_b200 import native
    lib = native.lib()
    buf = ctypes.create_string_buffer(64)
    p = ctypes.addressof(buf)
    rc = lib.mmu_smooth_knn(p, p, 4, 999, 0, 64, None, None, p + 8, p + 16, None)      # k out of range
    assert rc == 1 and b"k=999" in lib.mmu_last_error()
    rc = lib.mmu_knn_exact_f32(None, 1, None, None, 1, 1, 1, 0, 0, 0, 0, None, None, None)
    assert rc == 1 and b"null" in lib.mmu_last_error()
    assert lib.mmu_union_workspace_bytes(1000, 15) > 1000 * 15 * 24
    assert lib.mmu_union_workspace_bytes(0, 15) == 0


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback():
    """The product path must fail loudly when there is no CUDA device."""
    from umap_b200 import native
    from umap_b200 import graph as G
    with pytest.raises(native.NativeError):
        G.knn_graph(torch.zeros(40, 4), torch.zeros(40, 4), 5, True)
    import importlib
    model = importlib.import_module("impl.model")
    m = model.UMAPMixture(k_neighbors=5, out_dim=2, min_dist=0.1, num_encoders=1)
    with pytest.raises(native.NativeError):
        m.fit([torch.randn(100, 8)], epochs=1)
    with pytest.raises(ValueError):
        m.init([torch.randn(100, 8)], mode="bogus")          # ref: model.py:631-632


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "multimodal-umap_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f


def _reference_modules():
    """The reference's own impl/model.py and impl/util.py, imported under the package name `refimpl` (from
    /root/reference in the build container, from the staged copy baseline/_ref on the GPU box)."""
    import importlib
    import sys
    import types
    for base in ("/root/reference", os.path.join(ROOT, "baseline", "_ref")):
        if os.path.isfile(os.path.join(base, "impl", "model.py")):
            if "refimpl" not in sys.modules:
                pkg = types.ModuleType("refimpl")
                pkg.__path__ = [os.path.join(base, "impl")]
                sys.modules["refimpl"] = pkg
            return importlib.import_module("refimpl.model"), importlib.import_module("refimpl.util")
    return None, None


def test_api_surface_matches_reference_signatures():
    """Names, argument order and DEFAULTS of the mirrored API, compared with the live
    reference: every public callable of impl/model.py and impl/util.py must have the reference's signature."""
    import dataclasses
    import importlib
    import inspect
    ref_model, ref_util = _reference_modules()
    if ref_model is None:
        pytest.skip("reference not available (neither /root/reference nor baseline/_ref)")
    model = importlib.import_module("impl.model")
    util = importlib.import_module("impl.util")

    def params(fn):
        return [(n, p.default if p.default is not inspect.Parameter.empty else "<required>", p.kind)
                for n, p in inspect.signature(fn).parameters.items()]

    checked = 0
    for cls_name in ("UMAPEncoder", "UMAPMixture"):
        ref_cls, our_cls = getattr(ref_model, cls_name), getattr(model, cls_name)
        for name, member in vars(ref_cls).items():
            raw = member.__func__ if isinstance(member, (classmethod, staticmethod)) else member
            if not inspect.isfunction(raw) or (name.startswith("_") and name not in ("__init__", "_train")):
                continue                                     # the private loss helpers are replaced by kernels
            assert hasattr(our_cls, name), f"{cls_name}.{name} missing"
            ours = inspect.getattr_static(our_cls, name)
            assert type(ours) is type(member), f"{cls_name}.{name}: {type(member).__name__} expected"
            ours_raw = ours.__func__ if isinstance(ours, (classmethod, staticmethod)) else ours
            assert params(ours_raw) == params(raw), f"{cls_name}.{name}: {params(ours_raw)} != {params(raw)}"
            checked += 1
    for name in ("train", "embed", "recon", "embed_and_recon"):
        assert params(getattr(util, name)) == params(getattr(ref_util, name)), name
        checked += 1
    assert checked >= 18, checked
    assert [(f.name, f.default) for f in dataclasses.fields(util.Config)] == \
           [(f.name, f.default) for f in dataclasses.fields(ref_util.Config)]
    assert isinstance(model.device, torch.device) and isinstance(ref_model.device, torch.device)
    # public attributes the harness and checkpoints rely on (model.py:298-310, :26-31)
    m = model.UMAPMixture(k_neighbors=5, out_dim=2, min_dist=0.1, num_encoders=2)
    r = ref_model.UMAPMixture(k_neighbors=5, out_dim=2, min_dist=0.1, num_encoders=2)
    for attr in vars(r):
        assert hasattr(m, attr), f"UMAPMixture.{attr} missing"
    for attr in vars(r.encoders[0]):
        assert hasattr(m.encoders[0], attr), f"UMAPEncoder.{attr} missing"


def test_options_are_read_once_and_settable():
    """A/B switches: environment read at load, changed with mmu_set_option (never getenv per launch)."""
    from umap_b200 import native
    assert native.get_option("force_staged") in (0, 1)
    old = native.get_option("knn_window_mb")
    native.set_option("knn_window_mb", 7)
    assert native.get_option("knn_window_mb") == 7
    native.set_option("knn_window_mb", old)
    with pytest.raises(native.NativeError):
        native.set_option("no_such_option", 1)
    assert native.last_kernel("edge_forces") == "" and native.last_kernel("nonsense") == ""
    src = open(os.path.join(ROOT, "multimodal-umap_b200", "csrc", "layout_sgd.cu")).read() + \
        open(os.path.join(ROOT, "multimodal-umap_b200", "csrc", "knn_tc.cu")).read()
    assert "getenv" not in src


def test_ab_coefficients_match_reference(golden_dir):
    """get_ab_coeffs (model.py:587-618) with a closed-form Jacobian vs the reference's autograd one."""
    import importlib
    import numpy as np
    model = importlib.import_module("impl.model")
    g = np.load(os.path.join(golden_dir, "ab.npz"))
    m = model.UMAPMixture(k_neighbors=15, out_dim=2, min_dist=0.1, num_encoders=1)
    assert abs(m.a - float(g["a"])) < 2e-4 and abs(m.b - float(g["b"])) < 2e-4


def test_peer_entry_points_validate_arguments():
    """mmu_peer_barrier / mmu_adam_step_peer (include/mmumap.h: multi-GPU optimiser step over peer memory)
    reject bad geometry before any launch."""
    from umap_b200 import native
    lib = native.lib()
    buf = ctypes.create_string_buffer(256)
    p = ctypes.addressof(buf)
    ptrs = (ctypes.c_uint64 * 2)(p & ~15, (p & ~15) + 64)
    assert lib.mmu_peer_barrier(ptrs, 2, 2, 0, 1, None) == 1 and b"world/rank" in lib.mmu_last_error()
    assert lib.mmu_peer_barrier(ptrs, 2, 0, 2, 1, None) == 1 and b"slot" in lib.mmu_last_error()
    assert lib.mmu_peer_barrier(ptrs, native.PEER_MAX + 1, 0, 0, 1, None) == 1
    bad = (ctypes.c_uint64 * 2)(p & ~15, 0)
    assert lib.mmu_peer_barrier(bad, 2, 0, 0, 1, None) == 1 and b"peer pointer" in lib.mmu_last_error()
    assert lib.mmu_adam_step_peer(ptrs, ptrs, p & ~15, p & ~15, 6, 2, 0, 0.9, 0.999, 1e-8, p, None) == 1
    assert b"multiple of 4" in lib.mmu_last_error()
    assert lib.mmu_eigh_small(None, 4, None, None, None) == 1
    assert lib.mmu_eigh_small(p, 65, p, p, None) == 1 and b"n=65" in lib.mmu_last_error()


def test_round2_entry_points_validate_arguments():
    """Argument checks of the round-2 entry points run before any launch (no GPU needed): the staged / pruned kNN call,
    the record-form optimiser calls, the block eigensolver operations, the epoch tails, the roofs."""
    from umap_b200 import native
    lib = native.lib()
    buf = ctypes.create_string_buffer(4096)
    p = (ctypes.addressof(buf) + 255) & ~255
    err = lambda: lib.mmu_last_error()
    # kNN: stage mask, paired tile arguments, pinned split count for per-block tile sets
    args = [p, 256, p, 256, 8, 5, 1, 0, None, 1, 0, 0, p, 1 << 30, p, p, p, p]
    assert lib.mmu_knn_tc_ex(*args, 0, None, None, None, None, 0, None, None) == 1 and b"stages" in err()
    assert lib.mmu_knn_tc_ex(*args, 7, p, None, None, None, 0, None, None) == 1 and b"pairs" in err()
    assert lib.mmu_knn_tc_ex(*args, 2, p, p, None, None, 0, None, None) == 1 and b"pinned split count" in err()
    words = (ctypes.c_int64 * 12)()
    consts = (ctypes.c_float * 4)()
    assert lib.mmu_knn_tc_layout(1000, 5000, 128, 0, -2, 1, words, consts) == 0
    assert words[7] == 2 and words[9] == 64 and words[10] == 128 and words[11] == 256 and consts[0] > 0
    assert lib.mmu_knn_tc_workspace_bytes(1000, 5000, 128, 0, -2, 1) > int(words[4])
    # optimiser: record alignment, dimensions
    assert lib.mmu_edge_forces(p + 4, p, None, p, 1, 8, 10, p, p, p, p, 16, 1.5, 0.9, 0, p, None, 1, 0, None) == 1
    assert b"16-byte aligned" in err()
    assert lib.mmu_edge_forces(p, p, None, p, 1, 8, 10, p, p, p, p, 200, 1.5, 0.9, 0, p, None, 1, 0, None) == 1 and b"dim=200" in err()
    assert lib.mmu_edge_records(p, p, p, -1, 256, p, p, None) == 1
    assert lib.mmu_edge_sample_range(p, p, p, 5, 3, 256, 1, 0, p, p, p, p, None) == 1 and b"edge range" in err()
    # block eigensolver operations
    assert lib.mmu_block_ctl_words() >= 64
    assert lib.mmu_block_spmm(p, p, p, 10, p, 12, p, 0, None, p + 256, None) == 1 and b"block width 12" in err()
    assert lib.mmu_block_spmm(p, p, p, 10, p, 8, p, 2, None, p + 256, None) == 1 and b"needs z" in err()
    assert lib.mmu_block_gram(p, p, 10, 8, 2, p, p, None, None, None) == 1 and b"dinv" in err()
    assert lib.mmu_block_rotate(p, p, p, None, 10, 8, p, 0, None, None, None) == 1
    assert lib.mmu_block_ritz(p, 8, 9, 1e-3, 10, p, None) == 1
    assert lib.mmu_block_gram_workspace_bytes(32) == 4 * 592 * 32 * 32
    # epoch tails
    ptrs = (ctypes.c_uint64 * 2)(p, p + 1024)
    assert lib.mmu_epoch_tail_push(ptrs, ptrs, ptrs, p, p, p, 64, 8, 2, 0, 0.01, 0.9, 0.999, 1e-8, p, p, None) == 1
    assert b"inbox slot" in err()
    assert lib.mmu_epoch_tail_peer(ptrs, ptrs, ptrs, 16, 0, p, p, 64, 2, 0, 0, 0.01, 0.9, 0.999, 1e-8, p, p, None) == 1
    assert b"both multicast" in err()
    # roofs, farthest-point sampling
    assert lib.mmu_roof_random_rows(p, p, 100, 3, 1000, 0, 1, 1, p, None) == 1 and b"row_floats" in err()
    assert lib.mmu_fps_centroids(p, 100, 8, 4, p, 16, p, p, None) == 1 and b"workspace" in err()
    assert lib.mmu_fps_workspace_bytes(1000) == 4064
