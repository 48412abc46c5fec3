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


def test_api_surface_matches_reference_signatures():
    """Names, argument order and defaults of the mirrored API."""
    import importlib
    import inspect
    model = importlib.import_module("impl.model")
    util = importlib.import_module("impl.util")
    sig = inspect.signature
    assert list(sig(model.UMAPMixture.__init__).parameters) == ["self", "k_neighbors", "out_dim", "min_dist", "num_encoders"]
    p = sig(model.UMAPMixture.fit).parameters
    assert list(p) == ["self", "inputs", "epochs", "num_rep", "lr", "alpha", "batch_size"]
    assert (p["num_rep"].default, p["lr"].default, p["alpha"].default, p["batch_size"].default) == (8, 0.2, 0.5, 512)
    p = sig(model.UMAPMixture.transform).parameters
    assert list(p) == ["self", "inputs", "epochs", "data_indices", "num_rep", "lr", "alpha", "batch_size"]
    assert list(sig(model.UMAPMixture.inverse_transform).parameters) == list(p)
    p = sig(model.UMAPMixture._train).parameters
    assert list(p) == ["self", "embeds", "graphs", "epochs", "num_rep", "lr", "alpha", "batch_size", "mode",
                       "data_indices", "desc"]
    assert list(sig(model.UMAPEncoder.__init__).parameters) == ["self", "k_neighbors", "out_dim", "id"]
    p = sig(model.UMAPEncoder.fuzzy_knn_graph).parameters
    assert list(p) == ["self", "inputs", "mode", "query", "ref_data", "num_iters", "a", "b"]
    assert list(sig(model.UMAPEncoder.init).parameters) == ["self", "input", "mode", "query", "ref_data", "ref_embeds", "a", "b"]
    assert isinstance(inspect.getattr_static(model.UMAPMixture, "load_state_dict"), classmethod)
    import dataclasses
    assert [f.name for f in dataclasses.fields(util.Config)] == [
        "k_neighbors", "out_dim", "min_dist", "train_epochs", "num_rep", "lr", "alpha", "batch_size", "test_epochs"]
    for fn, names in (("train", ["data", "cfg"]), ("embed", ["model", "data", "src", "cfg"]),
                      ("recon", ["model", "embeds", "dst", "cfg"]),
                      ("embed_and_recon", ["model", "data", "src", "dst", "cfg"])):
        assert list(sig(getattr(util, fn)).parameters) == names
    assert isinstance(model.device, torch.device)


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
