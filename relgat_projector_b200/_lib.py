"""ctypes binding of librelgat_b200.so (the C ABI declared in include/relgat_b200.h).

There is no CPU fallback: every entry point launches sm_100a kernels.  If the shared library
cannot be loaded (or built with nvcc) the import of any op raises.
"""
from __future__ import annotations

import ctypes
import os
import re
from ctypes import c_int, c_longlong, c_ulonglong, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "librelgat_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "relgat_b200.h")

ERRORS = {
    -1: "RG_ERR_ARG (null pointer or negative size)",
    -2: "RG_ERR_SHAPE (shape not supported by the kernel mapping)",
    -3: "RG_ERR_ALIGN (pointer or stride not 16-byte aligned)",
    -4: "RG_ERR_WORKSPACE (workspace too small)",
    -5: "RG_ERR_DTYPE (dtype not supported)",
    -6: "RG_ERR_DRIVER (cuTensorMapEncodeTiled unavailable or failed)",
}

_P = c_void_p
_I = c_int
_L = c_longlong
_F = ctypes.c_float

# name -> (restype, argtypes); mirrors include/relgat_b200.h one to one
SIGNATURES = {
    "relgat_abi_version": (_I, []),
    "relgat_graph_index_workspace_bytes": (_L, [_L]),
    "relgat_graph_index_build": (_I, [_P, _P, _P, _L, _L, _L, _L] + [_P] * 11 + [_P, _L, _P]),
    "relgat_split_bf16": (_I, [_P, _P, _P, _L, _P]),
    "relgat_gemm_workspace_bytes": (_L, [_I, _I, _I, _I, _I, _I]),
    "relgat_gemm_bf16": (_I, [_P, _P, _L, _I, _P, _P, _L, _I, _P, _I, _L, _I, _I, _I, _I, _P, _L, _I, _P]),
    "relgat_gemm_tile_n": (_I, [_I]),
    "relgat_gemm_plan": (_L, [_I, _I, _I, _I, _P, _P, _P]),
    "relgat_gemm_dx_prep": (_I, [_P, _P, _L, _P, _P, _L, _P, _I, _I, _I, _P, _P, _P, _I, _F, _I, _I, _I, _P, _P, _P, _P, _I, _P]),
    "relgat_layer_fwd": (_I, [_P, _I, _L, _P, _P, _P, _P, _P, _P, _I, _P, _I, _P, _P, _I, _P, _P, _P, _P, _P, _P, _I, _P, _P, _P, _P,
                              _P, _I, _F, _P, _F, _P, _I, _I, _I, _I, _P, _P]),
    "relgat_layer_bwd_prep": (_I, [_P, _P, _P, _P, _I, _P, _P, _I, _I, _I, _I, _P, _I, _P, _I, _F, _P, _I, _P]),
    "relgat_layer_bwd_src": (_I, [_P, _L, _P, _I, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _P, _I, _P, _P, _I, _P, _P, _P, _P, _P,
                                  _P, _F, _P, _P, _I, _I, _L, _I, _I, _I, _I, _P, _P]),
    "relgat_bitmap_ranks_workspace_bytes": (_L, [_L]),
    "relgat_bitmap_ranks": (_I, [_P, _L, _P, _P, _P, _P, _L, _P]),
    "relgat_stream_chunks_workspace_bytes": (_L, [_I]),
    "relgat_stream_chunks_build": (_I, [_P, _I, _I, _I, _I, _I, _P, _I, _P, _I, _P, _P, _I, _P, _P, _L, _P]),
    "relgat_stream_chunks_for_rows": (_I, [_P, _P, _I, _P, _I, _I, _P, _I, _P, _I, _P, _P, _I, _P, _P, _L, _P]),
    "relgat_mark_rows": (_I, [_P, _L, _L, _P, _P]),
    "relgat_mark_sources": (_I, [_P, _P, _P, _I, _P, _P]),
    "relgat_layer_bwd_src2": (_I, [_P, _L, _P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _P, _I, _P, _P, _I, _P, _P, _P, _P, _L,
                                   _P, _F, _L, _I, _I, _I, _I, _P, _P]),
    "relgat_layer_bwd_src3": (_I, [_P, _L, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _P, _I, _P, _P, _I, _P, _P, _P, _P, _P, _F,
                                   _P, _P, _I, _L, _I, _I, _I, _I, _P, _P]),
    "relgat_layer_bwd_beta": (_I, [_P, _P, _P, _P, _P, _P, _I, _P, _P, _I, _I, _P]),
    "relgat_layer_bwd_rel": (_I, [_P, _I, _L, _P, _P, _P, _P, _P, _P, _P, _P, _I, _P, _P, _P, _P, _I, _I, _I, _P]),
    "relgat_score_fwd": (_I, [_I, _I, _P, _P, _P, _P, _P, _P, _I, _I, _P, _P, _I, _P, _P, _P]),
    "relgat_score_bwd": (_I, [_I, _I, _P, _P, _P, _P, _P, _P, _I, _I, _P, _P, _I, _P, _P, _P, _P]),
    "relgat_index_add_sorted": (_I, [_P, _P, _P, _P, _I, _I, _I, _P]),
    "relgat_margin_loss": (_I, [_P, _I, _I, ctypes.c_float, _I, _P, _P, _P]),
    "relgat_rank_loss": (_I, [_P, _P, _I, _I, _L, _L, _I, _F, _F, _I, _P, _P, _P, _P]),
    "relgat_recon_loss": (_I, [_P, _P, _P, _I, _I, _I, _L, _L, _F, _F, _F, _P, _P, _P, _P, _P, _P]),
    "relgat_gelu_layernorm_fwd": (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _F, _P]),
    "relgat_gelu_layernorm_groups": (_I, [_I]),
    "relgat_gelu_layernorm_bwd": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _P]),
    "relgat_bernoulli_bits": (_I, [_P, _L, _F, c_ulonglong, _P]),
    "relgat_zero_rows": (_I, [_P, _L, _P, _L, _I, _P]),
    "relgat_host_sample_batch": (_I, [_P, _P, _L, _P, _I, _I, _L, _P, _P, _P]),
    "relgat_host_shuffle": (_I, [_P, _P, _L]),
    "relgat_peer_table_granularity": (_I, [_I, _P]),
    "relgat_peer_table_create": (_I, [_I, c_ulonglong, _P, _P]),
    "relgat_peer_table_map": (_I, [_I, _I, _I, c_ulonglong, _P, c_ulonglong, _P]),
    "relgat_peer_table_unmap": (_I, [_P, _I, c_ulonglong, c_ulonglong]),
    "relgat_peer_table_last_driver_error": (_I, []),
    "relgat_pull_rows": (_I, [_P, _L, _P, _P, _L, _I, _P, _L, _I, _P]),
    "relgat_pull_rows_bf16": (_I, [_P, _L, _P, _P, _L, _I, _P, _L, _I, _P]),
}

ABI_VERSION = 13  # bumped whenever a signature in include/relgat_b200.h changes
_lib = None


def header_symbols() -> list[str]:
    """Function names declared in include/relgat_b200.h."""
    with open(HEADER_PATH, "r", encoding="utf-8") as f:
        text = f.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(relgat_[a-z0-9_]+)\s*\(", text)))


def load() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        from . import build as _build  # builds in-tree with nvcc; raises if nvcc is missing

        _build.build()
    lib = ctypes.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError = ABI mismatch: fail loudly
        fn.restype = res
        fn.argtypes = args
    if lib.relgat_abi_version() != ABI_VERSION:
        raise RuntimeError("librelgat_b200.so: ABI version mismatch, rebuild with relgat_projector_b200/build.py")
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc == 0:
        return
    if rc < 0:
        raise RuntimeError(f"{what}: {ERRORS.get(rc, rc)}")
    raise RuntimeError(f"{what}: CUDA error {rc} (cudaError_t)")


def ptr(t) -> int | None:
    """Device pointer of a torch tensor (None -> NULL)."""
    if t is None:
        return None
    return t.data_ptr()


def require_cuda(*tensors) -> None:
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError(
                "relgat_projector_b200 kernels run on a CUDA device only (sm_100a); got a CPU tensor. "
                "There is no CPU fallback."
            )
