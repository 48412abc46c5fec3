"""ctypes binding of libmmumap_b200.so (the C ABI declared in include/mmumap.h).

There is no CPU fallback: if the shared library is missing this module raises at import of
the first symbol, and every compute call raises NativeError when no CUDA device is present.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_double, c_float, c_int, c_int64, c_size_t, c_uint32, c_uint64, c_void_p

import torch

_PKG_DIR = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB_PATH = os.path.join(_PKG_DIR, "libmmumap_b200.so")

MAX_K = 64
KNN_TC_MAX_K = 32
OPT_STATE_WORDS = 8
KEPT_HDR_WORDS = 4          # [count, capacity, overflow flag, reserved]
PEER_MAX = 16
SIGMA_BISECT = 0
SIGMA_NEWTON = 1


class NativeError(RuntimeError):
    pass


# name -> (restype, argtypes); mirrors include/mmumap.h one to one
_SIGNATURES = {
    "mmu_abi_version": (c_int, []),
    "mmu_last_error": (ctypes.c_char_p, []),
    "mmu_launch_count": (c_uint64, []),
    "mmu_launch_count_add": (None, [c_uint64]),
    "mmu_device_info": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p]),
    "mmu_set_option": (c_int, [ctypes.c_char_p, c_int64]),
    "mmu_get_option": (c_int, [ctypes.c_char_p, c_void_p]),
    "mmu_last_kernel": (ctypes.c_char_p, [ctypes.c_char_p]),
    "mmu_knn_exact_f32": (c_int, [c_void_p, c_int64, c_void_p, c_void_p, c_int64, c_int, c_int, c_int, c_int64,
                                  c_int64, c_int, c_void_p, c_void_p, c_void_p]),
    "mmu_knn_tc_workspace_bytes": (c_size_t, [c_int64, c_int64, c_int, c_int, c_int, c_int]),
    "mmu_knn_tc": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int, c_int, c_int, c_int64, c_void_p, c_int, c_int, c_int,
                           c_void_p, c_size_t, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "mmu_knn_tc_ex": (c_int, [c_void_p, c_int64, c_void_p, c_int64, c_int, c_int, c_int, c_int64, c_void_p, c_int, c_int, c_int,
                              c_void_p, c_size_t, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p,
                              c_void_p, c_int, c_void_p, c_void_p]),
    "mmu_knn_tc_layout": (c_int, [c_int64, c_int64, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
    "mmu_fps_workspace_bytes": (c_size_t, [c_int64]),
    "mmu_fps_centroids": (c_int, [c_void_p, c_int64, c_int, c_int, c_void_p, c_size_t, c_void_p, c_void_p, c_void_p]),
    "mmu_knn_merge": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p, c_void_p]),
    "mmu_smooth_knn": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p,
                               c_void_p, c_void_p]),
    "mmu_invert_weights": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_float, c_float, c_void_p, c_void_p, c_void_p]),
    "mmu_union_workspace_bytes": (c_size_t, [c_int64, c_int]),
    "mmu_fuzzy_union": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_void_p, c_size_t, c_void_p, c_void_p,
                                c_void_p, c_void_p, c_void_p]),
    "mmu_union_rows_workspace_bytes": (c_size_t, [c_int64, c_int64]),
    "mmu_fuzzy_union_rows": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_void_p,
                                     c_size_t, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "mmu_embed_query": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_void_p, c_int, c_void_p, c_void_p]),
    "mmu_spmm_csr": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_int, c_void_p, c_void_p]),
    "mmu_spmm_csr_axpby": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_int, c_float, c_float, c_void_p,
                                   c_float, c_void_p, c_void_p]),
    "mmu_eigh_small": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "mmu_eigh_small_flag": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "mmu_block_ctl_words": (c_int, []),
    "mmu_block_ctl_init": (c_int, [c_void_p, c_void_p]),
    "mmu_block_spmm": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_void_p, c_int, c_void_p, c_int, c_void_p, c_void_p,
                               c_void_p]),
    "mmu_block_spmm_rows": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_void_p, c_int, c_void_p, c_int, c_void_p,
                                    c_void_p, c_void_p]),
    "mmu_block_gram_workspace_bytes": (c_size_t, [c_int]),
    "mmu_block_gram": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "mmu_block_rotate": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_void_p, c_int, c_void_p, c_void_p,
                                 c_void_p]),
    "mmu_block_ritz": (c_int, [c_void_p, c_int, c_int, c_float, c_int, c_void_p, c_void_p]),
    "mmu_block_svqb": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "mmu_block_cholqr": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p]),
    "mmu_peer_barrier": (c_int, [c_void_p, c_int, c_int, c_int, c_uint32, c_void_p]),
    "mmu_adam_step_peer": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int, c_int, c_double, c_double,
                                   c_double, c_void_p, c_void_p]),
    "mmu_epoch_tail_peer": (c_int, [c_void_p, c_void_p, c_void_p, c_uint64, c_uint64, c_void_p, c_void_p, c_int64, c_int, c_int,
                                    c_uint32, c_double, c_double, c_double, c_double, c_void_p, c_void_p, c_void_p]),
    "mmu_epoch_tail_push": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int, c_int,
                                    c_double, c_double, c_double, c_double, c_void_p, c_void_p, c_void_p]),
    "mmu_opt_state_init": (c_int, [c_void_p, c_void_p]),
    "mmu_opt_state_advance": (c_int, [c_void_p, c_double, c_double, c_double, c_void_p]),
    "mmu_edge_sample_range": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int, c_int, c_uint64, c_void_p,
                                      c_void_p, c_void_p, c_void_p, c_void_p]),
    "mmu_edge_sample_at": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int64, c_int, c_int, c_uint64, c_int64,
                                   c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "mmu_edge_records": (c_int, [c_void_p, c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p, c_void_p]),
    "mmu_edge_forces": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int64, c_void_p, c_void_p,
                                c_void_p, c_void_p, c_int, c_float, c_float, c_uint64, c_void_p, c_void_p, c_int,
                                c_int64, c_void_p]),
    "mmu_invert_forces": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int64, c_void_p, c_void_p,
                                  c_void_p, c_void_p, c_void_p, c_int, c_float, c_float, c_uint64, c_void_p, c_void_p,
                                  c_void_p]),
    "mmu_infonce": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_void_p, c_void_p, c_int, c_int, c_float, c_float,
                            c_void_p, c_void_p, c_uint64, c_uint32, c_void_p, c_void_p, c_void_p]),
    "mmu_infonce_range": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int, c_void_p, c_void_p, c_int, c_int,
                                  c_float, c_float, c_void_p, c_void_p, c_uint64, c_uint32, c_void_p, c_void_p, c_void_p]),
    "mmu_infonce_bidir": (c_int, [c_void_p, c_void_p, c_int64, c_int64, c_int64, c_int, c_void_p, c_void_p, c_void_p, c_void_p,
                                  c_int, c_int, c_float, c_float, c_void_p, c_void_p, c_uint64, c_uint32, c_void_p, c_void_p,
                                  c_void_p]),
    "mmu_roof_random_rows": (c_int, [c_void_p, c_void_p, c_int64, c_int, c_int64, c_uint64, c_int, c_int, c_void_p, c_void_p]),
    "mmu_adam_step": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_double, c_double, c_double, c_void_p,
                              c_int, c_void_p]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib = None


def lib() -> ctypes.CDLL:
    """Load the shared library (once).  Fails loudly when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NativeError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C multimodal-umap_b200/csrc`. There is no CPU fallback.")
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(handle, name)      # AttributeError if the .so lacks a declared symbol
            fn.restype = res
            fn.argtypes = args
        if handle.mmu_abi_version() != 1:
            raise NativeError("libmmumap_b200.so ABI version mismatch")
        _lib = handle
    return _lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = lib().mmu_last_error()
        raise NativeError(f"{what} failed (rc={rc}): {msg.decode() if msg else '?'}")


def require_cuda() -> None:
    if not torch.cuda.is_available():
        raise NativeError("no CUDA device: the B200 engine has no CPU fallback")


def ptr(t: torch.Tensor | None) -> int | None:
    """Device pointer of a contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise NativeError("expected a CUDA tensor at the C-ABI boundary")
    if not t.is_contiguous():
        raise NativeError("expected a contiguous tensor at the C-ABI boundary")
    return t.data_ptr()


def stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def set_option(name: str, value: int) -> None:
    """A/B switch of the kernels (include/mmumap.h: mmu_set_option); the environment is read once at load."""
    check(lib().mmu_set_option(name.encode(), int(value)), f"mmu_set_option({name})")


def get_option(name: str) -> int:
    v = c_int64()
    check(lib().mmu_get_option(name.encode(), ctypes.byref(v)), f"mmu_get_option({name})")
    return int(v.value)


def last_kernel(site: str) -> str:
    """Kernel variant the launch site chose last ("edge_forces", "knn_candidates", ...)."""
    return (lib().mmu_last_kernel(site.encode()) or b"").decode()


def device_info():
    require_cuda()
    sm, mj, mn, l2 = c_int(), c_int(), c_int(), c_size_t()
    check(lib().mmu_device_info(ctypes.byref(sm), ctypes.byref(mj), ctypes.byref(mn), ctypes.byref(l2)),
          "mmu_device_info")
    return {"sm_count": sm.value, "cc": (mj.value, mn.value), "l2_bytes": l2.value}
