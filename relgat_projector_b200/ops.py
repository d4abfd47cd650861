"""Thin tensor-level wrappers over the C ABI (one function per entry point).

PyTorch is plumbing here: it owns device memory and streams.  Every function launches on the
current stream of the tensors' device and returns torch tensors; no arithmetic is done in
torch.  CPU tensors are rejected (no fallback).
"""
from __future__ import annotations

import ctypes
import os
from typing import Optional, Tuple

import torch

from . import _lib
from .graph import GraphIndex

Planes = Tuple[torch.Tensor, Optional[torch.Tensor]]  # (hi, lo) bf16 planes of an fp32 matrix

SCORER_KIND = {"distmult": 0, "transe": 1}

LAUNCHES = 0  # kernels of librelgat_b200.so launched so far (bench.py reports the per-run delta)


def _count(n: int) -> None:
    global LAUNCHES
    LAUNCHES += n


def _stream(t: torch.Tensor) -> int:
    """Raw handle of the current stream of ``t``'s device (the fast internal query: this runs once per launch, and
    ``torch.cuda.current_stream`` costs ~8 us of Python per call — 0.4 ms per step on the small-step paths)."""
    return torch._C._cuda_getCurrentRawStream(t.device.index)


class _NoSwitch:
    def __enter__(self):
        return None

    def __exit__(self, *exc):
        return False


_NO_SWITCH = _NoSwitch()


def _on(device):
    """Context that makes ``device`` current for a launch; free when it already is (the usual case)."""
    dev = torch.device(device)
    return _NO_SWITCH if torch._C._cuda_getDevice() == dev.index else torch.cuda.device(dev)


def _f32c(t: torch.Tensor, name: str) -> torch.Tensor:
    _lib.require_cuda(t)
    if t.dtype != torch.float32:
        raise TypeError(f"{name} must be float32, got {t.dtype}")
    return t.contiguous()


def _feat(t: torch.Tensor, name: str) -> torch.Tensor:
    """Feature matrices are fp32 or bf16 (bf16 feature-storage mode)."""
    _lib.require_cuda(t)
    if t.dtype not in (torch.float32, torch.bfloat16):
        raise TypeError(f"{name} must be float32 or bfloat16, got {t.dtype}")
    return t.contiguous()


def _work_counter(device):
    """32 ints of scratch for the dynamic chunk claim (zeroed by the C entry point on the launch stream)."""
    if os.environ.get("RELGAT_STATIC_CHUNKS"):
        return None
    return torch.empty(32, dtype=torch.int32, device=device)


def _out_buf(buf: Optional[torch.Tensor], shape, device, name: str) -> torch.Tensor:
    """A fresh fp32 output, or the caller's buffer after checking it can be written in place."""
    if buf is None:
        return torch.empty(shape, dtype=torch.float32, device=device)
    _lib.require_cuda(buf)
    if buf.dtype != torch.float32 or tuple(buf.shape) != tuple(shape) or not buf.is_contiguous():
        raise ValueError(f"{name} must be a contiguous float32 tensor of shape {tuple(shape)}, got {tuple(buf.shape)}")
    return buf


_SM_COUNT = {}


_SM_LIMIT = [None]


def sm_count(device) -> int:
    """SMs the persistent kernels launched from here may fill: the device's, or the current ``sm_limit``."""
    idx = torch.device(device).index
    if idx not in _SM_COUNT:
        _SM_COUNT[idx] = torch.cuda.get_device_properties(device).multi_processor_count
    lim = _SM_LIMIT[0]
    return _SM_COUNT[idx] if lim is None else max(2, min(_SM_COUNT[idx], lim))


class sm_limit:
    """``with ops.sm_limit(n):`` the persistent kernels (GEMM, edge passes) launched inside size their grids for n SMs.
    Two such kernels on different streams then share the chip side by side instead of queueing behind each other (every
    one of them takes a whole SM's shared memory, so their CTAs never share an SM)."""

    def __init__(self, n: Optional[int]):
        self.n = None if n is None else int(n)

    def __enter__(self):
        self.prev = _SM_LIMIT[0]
        _SM_LIMIT[0] = self.n
        return self

    def __exit__(self, *exc):
        _SM_LIMIT[0] = self.prev
        return False


# ------------------------------------------------------------------------------------------
# dense transforms
# ------------------------------------------------------------------------------------------
def split_bf16(x: torch.Tensor, with_lo: bool = True, out_hi: Optional[torch.Tensor] = None) -> Planes:
    """fp32 matrix -> bf16 (hi, lo) planes, hi = rn(x), lo = rn(x - hi).  ``out_hi``: caller-owned contiguous bf16
    buffer of x's shape for the hi plane (rows of a peer table)."""
    x = _f32c(x, "x")
    if out_hi is not None:
        _lib.require_cuda(out_hi)
        if out_hi.dtype != torch.bfloat16 or tuple(out_hi.shape) != tuple(x.shape) or not out_hi.is_contiguous():
            raise ValueError("split_bf16: out_hi must be a contiguous bfloat16 tensor of x's shape")
    hi = out_hi if out_hi is not None else torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    lo = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device) if with_lo else None
    with _on(x.device):
        rc = _lib.load().relgat_split_bf16(_lib.ptr(x), _lib.ptr(hi), _lib.ptr(lo), x.numel(), _stream(x))
    _lib.check(rc, "relgat_split_bf16")
    _count(1)
    return hi, lo


def _tma_ready(t: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
    """TMA needs 16-byte aligned bases and row strides: pad the row stride to a multiple of 8 bf16
    (view of a zero-padded copy) in the rare case the feature width is not a multiple of 8."""
    if t is None or (t.stride(0) % 8 == 0 and t.data_ptr() % 16 == 0):
        return t
    cols = t.size(1)
    padded = torch.zeros((t.size(0), (cols + 7) // 8 * 8), dtype=t.dtype, device=t.device)
    padded[:, :cols].copy_(t)
    return padded[:, :cols]


def gemm(a: Planes, a_mn: bool, b: Planes, b_mn: bool, M: int, N: int, K: int,
         splits_k: int = 1, out: Optional[torch.Tensor] = None, out_dtype=torch.float32) -> torch.Tensor:
    """D[M, N] = A · Bᵀ on tcgen05 (fp32 accumulation; D fp32 or bf16).  Operands are 2-D bf16 planes:
    a_mn False: A stored [M, K]; True: stored [K, M].  Same for B with N."""
    a_hi, a_lo = a
    b_hi, b_lo = b
    _lib.require_cuda(a_hi, b_hi)
    for t in (a_hi, a_lo, b_hi, b_lo):
        if t is not None and (t.dtype != torch.bfloat16 or t.dim() != 2 or t.stride(1) != 1):
            raise TypeError("gemm operands must be 2-D bf16 tensors with unit inner stride")
    exp_a = (K, M) if a_mn else (M, K)
    exp_b = (K, N) if b_mn else (N, K)
    if tuple(a_hi.shape) != exp_a or tuple(b_hi.shape) != exp_b:
        raise ValueError(f"gemm shapes: A {tuple(a_hi.shape)} != {exp_a} or B {tuple(b_hi.shape)} != {exp_b}")
    if (a_lo is None) != (b_lo is None):
        raise ValueError("either both operands carry a lo plane (fp32-parity mode) or neither")
    a_hi, a_lo, b_hi, b_lo = _tma_ready(a_hi), _tma_ready(a_lo), _tma_ready(b_hi), _tma_ready(b_lo)
    dev = a_hi.device
    if out is None:
        out = torch.empty((M, N), dtype=out_dtype, device=dev)
    if out.dtype not in (torch.float32, torch.bfloat16) or out.stride(1) != 1:
        raise TypeError("gemm output must be fp32 or bf16 with unit inner stride")
    lib = _lib.load()
    ws = None
    ws_bytes = 0
    if splits_k > 1:
        ws_bytes = int(lib.relgat_gemm_workspace_bytes(M, N, K, int(a_mn), int(b_mn), splits_k))
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    with _on(dev):
        rc = lib.relgat_gemm_bf16(
            _lib.ptr(a_hi), _lib.ptr(a_lo), a_hi.stride(0), int(a_mn),
            _lib.ptr(b_hi), _lib.ptr(b_lo), b_hi.stride(0), int(b_mn),
            _lib.ptr(out), int(out.dtype == torch.bfloat16), out.stride(0), M, N, K, splits_k, _lib.ptr(ws), ws_bytes,
            sm_count(dev), _stream(out))
    _lib.check(rc, "relgat_gemm_bf16")
    _count(2 if splits_k > 1 else 1)
    return out


def gemm_dx_prep_supported(N: int, F: int) -> bool:
    """The fused prep epilogue needs an N tile inside at most two heads (tile width <= F) and 16-column pieces."""
    return N % 16 == 0 and N % F == 0 and int(_lib.load().relgat_gemm_tile_n(int(N))) <= F


def gemm_dx_prep(dP: Planes, WT: Planes, M: int, N: int, K: int, y: torch.Tensor, bias: torch.Tensor, H: int, F: int,
                 apply_elu: bool, feat_drop: Optional["DropMask"] = None):
    """dX = dP · W with the backward prep of the layer below fused into the GEMM epilogue.
    Returns (G [M, N] = dX * act'(y) * mask, t [M, H], hsum [M, H]) — what ``edge_bwd_prep(dX, y, bias, ...)`` returns,
    without materialising dX (saves one [M, N] write and two [M, N] reads per hidden layer)."""
    a_hi, a_lo = dP
    b_hi, b_lo = WT
    _lib.require_cuda(a_hi, b_hi, y, bias)
    if tuple(a_hi.shape) != (M, K) or tuple(b_hi.shape) != (N, K) or (a_lo is None) != (b_lo is None):
        raise ValueError("gemm_dx_prep: dP planes [M, K], W^T planes [N, K] expected")
    y, bias = _f32c(y, "y"), _f32c(bias, "bias")
    if tuple(y.shape) != (M, N) or N != H * F or bias.numel() != M:
        raise ValueError("gemm_dx_prep: y [M, H*F] and bias [M] expected")
    a_hi, a_lo, b_hi, b_lo = _tma_ready(a_hi), _tma_ready(a_lo), _tma_ready(b_hi), _tma_ready(b_lo)
    dev = y.device
    lib = _lib.load()
    n_tiles = -(-N // int(lib.relgat_gemm_tile_n(N)))
    G = torch.empty((M, N), dtype=torch.float32, device=dev)
    tpart = torch.empty((M, n_tiles, 2), dtype=torch.float32, device=dev)
    hpart = torch.empty((M, n_tiles, 2), dtype=torch.float32, device=dev)
    t = torch.empty((M, H), dtype=torch.float32, device=dev)
    hsum = torch.empty((M, H), dtype=torch.float32, device=dev)
    with _on(dev):
        rc = lib.relgat_gemm_dx_prep(_lib.ptr(a_hi), _lib.ptr(a_lo), a_hi.stride(0), _lib.ptr(b_hi), _lib.ptr(b_lo),
                                     b_hi.stride(0), _lib.ptr(G), M, N, K, _lib.ptr(y), _lib.ptr(bias),
                                     *_feat_mask_args(feat_drop, M, N), H, F, int(apply_elu), _lib.ptr(tpart),
                                     _lib.ptr(hpart), _lib.ptr(t), _lib.ptr(hsum), sm_count(dev), _stream(y))
    _lib.check(rc, "relgat_gemm_dx_prep")
    _count(2)
    return G, t, hsum


def gemm_plan(M: int, N: int, b_mn: bool = True, sms: int = 148):
    """(cost, tile_m, tile_n, slots) of the tcgen05 GEMM for an [M, N] output: rows and columns of one work unit (256
    rows when it runs as CTA pairs; two N tiles when clusters of two pairs multicast their shared A rows), the units
    in flight at once, and the modelled cost of one k-block over all tiles (L2 -> shared-memory bytes or MMA clocks,
    whichever is slower) — comparable between the two orientations of a weight-gradient GEMM."""
    tm, tn, sl = ctypes.c_int(0), ctypes.c_int(0), ctypes.c_int(0)
    cost = int(_lib.load().relgat_gemm_plan(int(M), int(N), int(b_mn), int(sms), ctypes.byref(tm), ctypes.byref(tn),
                                            ctypes.byref(sl)))
    if cost < 0:
        _lib.check(cost, "relgat_gemm_plan")
    return cost, tm.value, tn.value, sl.value


def gemm_cost_model(M: int, N: int) -> float:
    """Modelled cost of a split-K weight-gradient GEMM (both operands MN-major) with an [M, N] output — used to choose
    which operand becomes M."""
    return float(gemm_plan(M, N, True)[0])


def pick_splits_k(M: int, N: int, K: int, device) -> int:
    """Split-K factor for GEMMs with few output tiles and a long reduction (the dW GEMMs): the one (<= 16) that fills
    the waves of the persistent grid (one CTA, CTA pair or cluster per work unit) best, smaller factors winning ties."""
    _, tm, tn, slots = gemm_plan(M, N, True, sm_count(device))
    tiles = (-(-M // tm)) * (-(-N // tn))
    kb = (K + 63) // 64
    if tiles >= 4 * slots or kb < 16:
        return 1
    best, best_eff = 1, 0.0
    for sk in range(1, 17):
        if kb // sk < 8:
            break
        units = tiles * sk
        eff = units / (-(-units // slots) * slots)
        if eff > best_eff + 0.02:
            best, best_eff = sk, eff
    return best


# ------------------------------------------------------------------------------------------
# dropout masks (keep bits: element i -> word i >> 5, bit i & 31)
# ------------------------------------------------------------------------------------------
class DropMask:
    """Keep-bit mask of one dropout site plus the 1/(1-p) scale of the kept elements.

    Feature masks: ``bits`` int32 [rows, words] with words = ceil(C / 32) (bit = column inside the row);
    attention masks: ``bits`` int32 [ceil(E*H / 32)] (bit index = csr slot * H + head)."""

    def __init__(self, bits: torch.Tensor, scale: float):
        _lib.require_cuda(bits)
        if bits.dtype != torch.int32 or not bits.is_contiguous():
            raise TypeError("DropMask bits must be a contiguous int32 tensor")
        self.bits, self.scale = bits, float(scale)

    @staticmethod
    def draw(shape, p: float, device, seed: Optional[int] = None) -> "DropMask":
        """Bernoulli(1-p) keep bits from Philox4x32-10; the seed is drawn from torch's default (CPU) generator, so
        ``torch.manual_seed`` governs the masks and no device sync is needed."""
        if not 0.0 <= p < 1.0:
            raise ValueError(f"dropout probability has to be in [0, 1), got {p}")
        if seed is None:
            seed = int(torch.randint(0, 2 ** 62, (1,)).item())
        bits = torch.empty(shape, dtype=torch.int32, device=device)
        with _on(bits.device):
            rc = _lib.load().relgat_bernoulli_bits(_lib.ptr(bits), bits.numel(), float(p), seed, _stream(bits))
        _lib.check(rc, "relgat_bernoulli_bits")
        _count(1)
        return DropMask(bits, 1.0 / (1.0 - p))

    @staticmethod
    def feature_mask(keep: torch.Tensor, p: float) -> "DropMask":
        """Packs a boolean keep tensor [rows, C] (tests inject masks this way and replay them in the oracle)."""
        return DropMask(_pack_bits(keep.to(torch.bool), per_row=True), 1.0 / (1.0 - p))

    @staticmethod
    def edge_mask(keep: torch.Tensor, p: float) -> "DropMask":
        """Packs a boolean keep tensor [E, H] in CSR slot order."""
        return DropMask(_pack_bits(keep.to(torch.bool).reshape(1, -1), per_row=True).reshape(-1), 1.0 / (1.0 - p))


def _pack_bits(keep: torch.Tensor, per_row: bool) -> torch.Tensor:
    rows, cols = keep.shape
    words = (cols + 31) // 32
    pad = torch.zeros((rows, words * 32), dtype=torch.int64, device=keep.device)
    pad[:, :cols] = keep.to(torch.int64)
    w = (pad.view(rows, words, 32) << torch.arange(32, device=keep.device, dtype=torch.int64)).sum(-1)
    w = torch.where(w >= 2 ** 31, w - 2 ** 32, w)  # two's complement into int32
    return w.to(torch.int32).contiguous()


def _feat_mask_args(m: Optional[DropMask], rows: int, C: int):
    if m is None:
        return None, 0, 1.0
    if m.bits.dim() != 2 or m.bits.size(0) != rows or m.bits.size(1) * 32 < C:
        raise ValueError(f"feature dropout mask must be int32 [{rows}, >= {(C + 31) // 32}], got {tuple(m.bits.shape)}")
    return _lib.ptr(m.bits), int(m.bits.size(1)), m.scale


def _edge_mask_args(m: Optional[DropMask], E: int, H: int):
    if m is None:
        return None, 1.0
    if m.bits.numel() * 32 < E * H:
        raise ValueError(f"attention dropout mask needs >= {(E * H + 31) // 32} words, got {m.bits.numel()}")
    return _lib.ptr(m.bits), m.scale


def zero_rows(table: torch.Tensor, ids: torch.Tensor) -> None:
    """table[ids, :] = 0 (fp32 rows; ids int64, may repeat)."""
    _lib.require_cuda(table, ids)
    if table.dtype != torch.float32 or table.dim() != 2 or table.stride(1) != 1 or ids.dtype != torch.int64:
        raise TypeError("zero_rows: table must be a 2-D float32 tensor with unit inner stride, ids int64")
    if ids.numel() == 0:
        return
    with _on(table.device):
        rc = _lib.load().relgat_zero_rows(_lib.ptr(table), table.stride(0), _lib.ptr(ids.contiguous()), ids.numel(),
                                          table.size(1), _stream(table))
    _lib.check(rc, "relgat_zero_rows")
    _count(1)

# ------------------------------------------------------------------------------------------
# edge kernels
# ------------------------------------------------------------------------------------------
def edge_fwd(P: torch.Tensor, A: torch.Tensor, beta: Optional[torch.Tensor], g: GraphIndex, H: int, F: int,
             want_act: bool = False, apply_elu: bool = False, act_lo: bool = True, want_out: bool = True,
             want_alpha: bool = False, z_out: Optional[torch.Tensor] = None,
             minv_out: Optional[torch.Tensor] = None, out_buf: Optional[torch.Tensor] = None,
             feat_drop: Optional["DropMask"] = None, edge_drop: Optional["DropMask"] = None,
             chunks=None, src_row: Optional[torch.Tensor] = None):
    """Returns (out [N, H*F] fp32 or None, act planes or None, alpha [E,H] or None, z [E,H],
    minv [N,H,2], bias [N]).  ``z_out`` / ``minv_out`` / ``out_buf``: caller-owned buffers for the saved
    statistics and the output rows (rows of a peer table on the partitioned path).
    ``feat_drop`` / ``edge_drop``: keep-bit masks of the feature dropout (reference layer.py:321-322; ``out`` then
    holds the POST-dropout rows) and of the attention dropout (layer.py:296-297).
    ``chunks`` (a StreamChunks over a list of destinations) + ``src_row`` (int32 [N_src] compact numbering of the rows
    P holds): the receptive-field forward — only the listed destinations are computed, every other row of the outputs
    is left unwritten."""
    P = _feat(P, "P")
    A = _f32c(A, "A")
    if beta is not None:
        beta = _f32c(beta, "beta")
    dev = P.device
    N, E, R, C = g.N, g.E, g.R, H * F
    if src_row is not None:
        if src_row.dtype != torch.int32 or src_row.device != dev or src_row.numel() != g.N_src or not src_row.is_contiguous():
            raise ValueError("src_row must be a contiguous int32 [N_src] tensor on the feature device")
        if want_alpha or P.dim() != 2 or P.size(1) != C:
            raise ValueError(f"compacted sources: P must be [rows, {C}] and alpha cannot be requested")
    elif P.dim() != 2 or P.size(1) != C or P.size(0) != g.N_src:
        raise ValueError(f"P must be [{g.N_src}, {C}], got {tuple(P.shape)}")
    if tuple(A.shape) != (H, R, F):
        raise ValueError(f"A must be [{H}, {R}, {F}], got {tuple(A.shape)}")
    out = _out_buf(out_buf, (N, C), dev, "out_buf") if want_out else None
    hi = torch.empty((N, C), dtype=torch.bfloat16, device=dev) if want_act else None
    lo = torch.empty((N, C), dtype=torch.bfloat16, device=dev) if (want_act and act_lo) else None
    alpha = torch.empty((E, H), dtype=torch.float32, device=dev) if want_alpha else None
    z = _out_buf(z_out, (E, H), dev, "z_out")
    minv = _out_buf(minv_out, (N, H, 2), dev, "minv_out")
    bias = torch.empty((N,), dtype=torch.float32, device=dev)
    ck = chunks if chunks is not None else g.fwd_chunks
    part_ml = torch.empty((ck.n_parts, H, 2), dtype=torch.float32, device=dev) if ck.n_parts else None
    part_b = torch.empty((ck.n_parts,), dtype=torch.float32, device=dev) if ck.n_parts else None
    part_acc = torch.empty((ck.n_parts, C), dtype=torch.float32, device=dev) if ck.n_parts else None
    with _on(dev):
        rc = _lib.load().relgat_layer_fwd(
            _lib.ptr(P), int(P.dtype == torch.bfloat16), P.stride(0), _lib.ptr(A), _lib.ptr(beta),
            _lib.ptr(g.rowptr), _lib.ptr(g.csr_src), _lib.ptr(g.csr_rel),
            _lib.ptr(ck.chunks), ck.n_chunks, _lib.ptr(ck.parts), ck.n_parts,
            _lib.ptr(ck.long_node), _lib.ptr(ck.long_part_ptr), ck.n_long,
            _lib.ptr(part_ml), _lib.ptr(part_b), _lib.ptr(part_acc),
            _lib.ptr(out), _lib.ptr(hi), _lib.ptr(lo), int(apply_elu),
            _lib.ptr(alpha), _lib.ptr(z), _lib.ptr(minv), _lib.ptr(bias),
            *_feat_mask_args(feat_drop, N, C), *_edge_mask_args(edge_drop, E, H), _lib.ptr(src_row), H, F, R,
            sm_count(dev), _lib.ptr(_work_counter(dev)), _stream(P))
    _lib.check(rc, "relgat_layer_fwd")
    _count(2 if ck.n_long else 1)
    return out, ((hi, lo) if want_act else None), alpha, z, minv, bias


def _export_buf(buf: Optional[torch.Tensor], like: torch.Tensor) -> Optional[torch.Tensor]:
    if buf is None:
        return None
    _lib.require_cuda(buf)
    if buf.dtype != torch.bfloat16 or tuple(buf.shape) != tuple(like.shape) or not buf.is_contiguous():
        raise ValueError("g_export must be a contiguous bfloat16 tensor of dY's shape")
    return buf


def edge_bwd_prep(dY: torch.Tensor, out: torch.Tensor, bias: torch.Tensor, H: int, F: int,
                  apply_elu: bool, inplace: bool = False, g_bf16: bool = False,
                  G_out: Optional[torch.Tensor] = None, t_out: Optional[torch.Tensor] = None,
                  hsum_out: Optional[torch.Tensor] = None, nonzero_rows: Optional[torch.Tensor] = None,
                  feat_drop: Optional["DropMask"] = None, g_export: Optional[torch.Tensor] = None,
                  compact_rows: Optional[torch.Tensor] = None):
    """Returns (G [N,C] fp32 or bf16, t [N,H], hsum [N,H]).  ``G_out`` / ``t_out`` / ``hsum_out``:
    caller-owned fp32 buffers (rows of a peer table on the partitioned path).  ``nonzero_rows`` (int64, may
    repeat): all other rows of dY are known to be zero; used only when G can alias dY (fp32, no activation).
    ``compact_rows`` (int64 [n], distinct): dY is [n, C] and holds the gradient of node compact_rows[k] in row k; G_out
    is an [N, C] table whose other rows are zero (the caller clears the written rows afterwards); t / hsum come back
    dense with zeros elsewhere."""
    dY = _f32c(dY, "dY")
    out = _f32c(out, "out")
    N = out.size(0)
    if compact_rows is not None:
        rows = _ids(compact_rows, "compact_rows")
        if G_out is None or tuple(G_out.shape) != tuple(out.shape) or dY.size(0) != rows.numel() or dY.size(1) != out.size(1):
            raise ValueError("compact_rows: dY must be [len(rows), C] and G_out an [N, C] table")
        G = _out_buf(G_out, tuple(out.shape), dY.device, "G_out")
        t = _out_buf(t_out, (N, H), dY.device, "t_out")
        hsum = _out_buf(hsum_out, (N, H), dY.device, "hsum_out")
        with _on(dY.device):
            rc = _lib.load().relgat_layer_bwd_prep(_lib.ptr(dY), _lib.ptr(out), _lib.ptr(bias), _lib.ptr(G), 0,
                                                   _lib.ptr(t), _lib.ptr(hsum), N, H, F, int(apply_elu),
                                                   _lib.ptr(rows), int(rows.numel()),
                                                   *_feat_mask_args(feat_drop, N, H * F), None, 1, _stream(dY))
        _lib.check(rc, "relgat_layer_bwd_prep")
        _count(1 if rows.numel() else 0)
        return G, t, hsum
    if G_out is not None:
        G = _out_buf(G_out, tuple(dY.shape), dY.device, "G_out")
        g_bf16 = False
    elif g_bf16:
        G = torch.empty(dY.shape, dtype=torch.bfloat16, device=dY.device)
    else:
        # without an activation G == dY: nothing to write, alias it (saves a full [N, C] copy)
        G = dY if (inplace or not (apply_elu or feat_drop is not None)) else torch.empty_like(dY)
    t = _out_buf(t_out, (N, H), dY.device, "t_out")
    hsum = _out_buf(hsum_out, (N, H), dY.device, "hsum_out")
    rows = None
    if nonzero_rows is not None and not apply_elu and not g_bf16 and (G.data_ptr() == dY.data_ptr() or G_out is not None):
        rows = _ids(nonzero_rows, "nonzero_rows")  # with G_out the caller keeps the other rows of G at zero
    with _on(dY.device):
        rc = _lib.load().relgat_layer_bwd_prep(_lib.ptr(dY), _lib.ptr(out), _lib.ptr(bias), _lib.ptr(G), int(g_bf16),
                                               _lib.ptr(t), _lib.ptr(hsum), N, H, F, int(apply_elu),
                                               _lib.ptr(rows), 0 if rows is None else int(rows.numel()),
                                               *_feat_mask_args(feat_drop, N, H * F), _lib.ptr(_export_buf(g_export, dY)),
                                               0, _stream(dY))
    _lib.check(rc, "relgat_layer_bwd_prep")
    _count(1)
    return G, t, hsum


# second-generation by-source kernel (csrc/edge_bwd_src2.cu): 26 % fewer instructions, measured at the SAME time as
# the first generation (1.36 vs 1.35 ms at config 2: the pass is bound by rows in flight, not by issue) — opt-in
SRC_V2 = os.environ.get("RELGAT_SRC_V2", "0") != "0"


def ds_row_width(H: int, F: int, R: int) -> int:
    """Row width of the widened dP rows [dP | dS] (multiple of 8 elements: TMA strides, 128-bit stores)."""
    return (H * F + H * R + 7) // 8 * 8


def src3_supported(P: torch.Tensor, F: int) -> bool:
    """Layouts the third-generation by-source kernel covers: fp32 rows, 128-bit vectors, 16-byte aligned row starts."""
    return (P.dtype == torch.float32 and F % 4 == 0 and P.dim() == 2 and P.stride(1) == 1 and P.stride(0) % 4 == 0
            and P.data_ptr() % 16 == 0)


def edge_bwd_src(P, G, A, z, minv, t, g: GraphIndex, H: int, F: int, want_fp32: bool = True,
                 want_planes: bool = False, planes_lo: bool = True, edge_drop: Optional["DropMask"] = None,
                 want_ds: bool = False, dst_nz: Optional[torch.Tensor] = None,
                 src_rows: Optional[Tuple[torch.Tensor, int]] = None, p_compact: bool = False,
                 a_term: bool = True):
    """Returns (dP fp32 or None, dP planes or None, dz [E,H] or None).  ``want_ds``: the rows are
    ``ds_row_width`` wide, columns H*F + h*R + r hold dS and dz is not written.
    ``dst_nz`` (want_ds only): row bitmap from ``mark_rows`` / ``mark_sources`` — rows of G outside it are exact zeros
    (and so are their t), the edges into them are skipped.  ``src_rows`` = (rank int32 [N_src], n) from
    ``bitmap_ranks`` of the sources of the marked rows: the output has n rows, source i is written to row rank[i] and
    sources with rank -1 are skipped altogether; ``p_compact``: P itself is compact (row rank[i] holds source i)."""
    P = _feat(P, "P")
    G = _feat(G, "G")
    if P.dtype != G.dtype:
        raise TypeError("P and G must share one storage type (both fp32 or both bf16)")
    A = _f32c(A, "A")
    dev = P.device
    n_src, C = g.N_src, H * F
    if p_compact and src_rows is None:
        raise ValueError("p_compact needs src_rows")
    if (P.size(0) != (int(src_rows[1]) if p_compact else n_src)) or G.size(0) != g.N:
        raise ValueError("P / G row counts do not match the graph index")
    W = ds_row_width(H, F, g.R) if want_ds else C
    mk = torch.zeros if W > C + H * g.R else torch.empty  # padding columns feed the GEMM: keep them finite
    n_out, rank = n_src, None
    if src_rows is not None:
        rank, n_out = src_rows[0], int(src_rows[1])
        if not want_ds or dst_nz is None:
            raise ValueError("src_rows needs want_ds and dst_nz (the sources of the marked rows)")
        if rank.dtype != torch.int32 or rank.device != dev or rank.numel() != n_src or not rank.is_contiguous():
            raise ValueError("src_rows: rank must be a contiguous int32 [N_src] tensor on the feature device")
    dP = mk((n_out, W), dtype=torch.float32, device=dev) if want_fp32 else None
    hi = mk((n_out, W), dtype=torch.bfloat16, device=dev) if want_planes else None
    lo = mk((n_out, W), dtype=torch.bfloat16, device=dev) if (want_planes and planes_lo) else None
    dz = None if want_ds else torch.empty((g.E, H), dtype=torch.float32, device=dev)
    ck = g.src_chunks
    part_acc = torch.empty((ck.n_parts, W), dtype=torch.float32, device=dev) if ck.n_parts else None
    if dst_nz is not None:
        if not want_ds:
            raise ValueError("dst_nz needs want_ds (the dz output of the plain variant is not written for skipped edges)")
        if dst_nz.dtype != torch.int32 or dst_nz.device != dev or dst_nz.numel() < (g.N + 31) // 32:
            raise ValueError("dst_nz must be an int32 bitmap of ceil(N / 32) words on the feature device")
    if not a_term:
        # third generation (csrc/edge_bwd_src3.cu): rows [dPa | dS] WITHOUT the dz * A[rel] term — the caller folds
        # dS·A into the GEMMs that consume the rows; gathered rows travel as bulk async copies, 4 deep per warp
        if not want_ds or not src3_supported(P, F) or not src3_supported(G, F):
            raise ValueError("a_term=False needs want_ds and fp32 rows with F % 4 == 0 (see src3_supported)")
        with _on(dev):
            rc = _lib.load().relgat_layer_bwd_src3(
                _lib.ptr(P), P.stride(0), _lib.ptr(G), _lib.ptr(z), _lib.ptr(minv), _lib.ptr(t),
                _lib.ptr(g.colptr), _lib.ptr(g.csc_slot), _lib.ptr(g.csc_dst), _lib.ptr(g.csc_rel),
                _lib.ptr(ck.chunks), ck.n_chunks, _lib.ptr(ck.parts), ck.n_parts,
                _lib.ptr(ck.long_node), _lib.ptr(ck.long_part_ptr), ck.n_long, _lib.ptr(part_acc),
                _lib.ptr(dP), _lib.ptr(hi), _lib.ptr(lo), *_edge_mask_args(edge_drop, g.E, H),
                _lib.ptr(dst_nz), _lib.ptr(rank), int(p_compact), W, H, F, g.R, sm_count(dev),
                _lib.ptr(_work_counter(dev)), _stream(P))
        _lib.check(rc, "relgat_layer_bwd_src3")
        _count(2 if ck.n_long else 1)
        return dP, ((hi, lo) if want_planes else None), dz
    if (SRC_V2 and want_ds and want_planes and not want_fp32 and P.dtype == torch.float32 and F % 4 == 0
            and P.stride(0) % 4 == 0 and dst_nz is None):
        # second-generation kernel of the training path (coefficient pre-pass + lean edge loop)
        coef = torch.empty((max(g.E, 1), H, 4), dtype=torch.float32, device=dev)
        mask_ptr, mask_scale = _edge_mask_args(edge_drop, g.E, H)
        with _on(dev):
            rc = _lib.load().relgat_layer_bwd_src2(
                _lib.ptr(P), P.stride(0), _lib.ptr(G), _lib.ptr(A), _lib.ptr(z), _lib.ptr(minv), _lib.ptr(t),
                _lib.ptr(g.colptr), _lib.ptr(g.csc_slot), _lib.ptr(g.csc_dst), _lib.ptr(g.csc_rel),
                _lib.ptr(ck.chunks), ck.n_chunks, _lib.ptr(ck.parts), ck.n_parts,
                _lib.ptr(ck.long_node), _lib.ptr(ck.long_part_ptr), ck.n_long, _lib.ptr(part_acc),
                _lib.ptr(hi), _lib.ptr(lo), _lib.ptr(coef), g.E, mask_ptr, mask_scale, W, H, F, g.R, sm_count(dev),
                _lib.ptr(_work_counter(dev)), _stream(P))
        if rc != -2:  # RG_ERR_SHAPE: layout not covered -> first-generation kernel below
            _lib.check(rc, "relgat_layer_bwd_src2")
            _count(3 if ck.n_long else 2)
            return dP, (hi, lo), dz
    with _on(dev):
        rc = _lib.load().relgat_layer_bwd_src(
            _lib.ptr(P), P.stride(0), _lib.ptr(G), int(P.dtype == torch.bfloat16), _lib.ptr(A), _lib.ptr(z),
            _lib.ptr(minv), _lib.ptr(t), _lib.ptr(g.colptr), _lib.ptr(g.csc_slot), _lib.ptr(g.csc_dst), _lib.ptr(g.csc_rel),
            _lib.ptr(ck.chunks), ck.n_chunks, _lib.ptr(ck.parts), ck.n_parts,
            _lib.ptr(ck.long_node), _lib.ptr(ck.long_part_ptr), ck.n_long, _lib.ptr(part_acc),
            _lib.ptr(dP), _lib.ptr(hi), _lib.ptr(lo), _lib.ptr(dz), *_edge_mask_args(edge_drop, g.E, H),
            _lib.ptr(dst_nz), _lib.ptr(rank), int(p_compact), int(want_ds), W, H, F, g.R, sm_count(dev), _lib.ptr(_work_counter(dev)), _stream(P))
    _lib.check(rc, "relgat_layer_bwd_src")
    _count(2 if ck.n_long else 1)
    return dP, ((hi, lo) if want_planes else None), dz


def row_bitmap(n_rows: int, device) -> torch.Tensor:
    """All-clear row bitmap (int32 words, bit j of word j >> 5 = row j) for mark_rows / mark_sources."""
    return torch.zeros(((n_rows + 31) // 32,), dtype=torch.int32, device=device)


def mark_rows(ids: torch.Tensor, n_rows: int, bits: Optional[torch.Tensor] = None) -> torch.Tensor:
    """bits |= {ids}: the rows of a gradient table that may be non-zero (ids outside [0, n_rows) are ignored)."""
    _lib.require_cuda(ids)
    if ids.dtype != torch.int64 or ids.dim() != 1:
        raise TypeError("mark_rows: ids must be a 1-D int64 tensor")
    ids = ids.contiguous()
    if bits is None:
        bits = row_bitmap(n_rows, ids.device)
    with _on(ids.device):
        _lib.check(_lib.load().relgat_mark_rows(_lib.ptr(ids), ids.numel(), n_rows, _lib.ptr(bits), _stream(ids)),
                   "relgat_mark_rows")
    _count(1 if ids.numel() else 0)
    return bits


def mark_sources(dst_bits: torch.Tensor, g: GraphIndex) -> torch.Tensor:
    """Bitmap of the sources of the edges into the rows marked in ``dst_bits``: the rows of dP — and of the gradient
    handed to the layer below — that can be non-zero."""
    src_bits = row_bitmap(g.N_src, dst_bits.device)
    with _on(dst_bits.device):
        _lib.check(_lib.load().relgat_mark_sources(_lib.ptr(dst_bits), _lib.ptr(g.rowptr), _lib.ptr(g.csr_src), g.N,
                                                   _lib.ptr(src_bits), _stream(dst_bits)), "relgat_mark_sources")
    _count(1 if g.N else 0)
    return src_bits


def bitmap_ranks(bits: torch.Tensor, n_rows: int, want_list: bool = True):
    """(rank int32 [n_rows] — position of each marked row among the marked rows or -1, list int64 [n_rows] whose first
    ``count`` entries are the marked rows in ascending order (or None), count int32 [1]) — all on the device."""
    _lib.require_cuda(bits)
    if bits.dtype != torch.int32 or bits.numel() < (n_rows + 31) // 32 or not bits.is_contiguous():
        raise ValueError("bits must be a contiguous int32 bitmap of ceil(n_rows / 32) words")
    dev = bits.device
    rank = torch.empty((n_rows,), dtype=torch.int32, device=dev)
    lst = torch.empty((n_rows,), dtype=torch.int64, device=dev) if want_list else None
    count = torch.empty((1,), dtype=torch.int32, device=dev)
    lib = _lib.load()
    ws_bytes = int(lib.relgat_bitmap_ranks_workspace_bytes(n_rows))
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    with _on(dev):
        _lib.check(lib.relgat_bitmap_ranks(_lib.ptr(bits), n_rows, _lib.ptr(rank), _lib.ptr(lst), _lib.ptr(count),
                                           _lib.ptr(ws), ws_bytes, _stream(bits)), "relgat_bitmap_ranks")
    _count(3 if n_rows else 0)
    return rank, lst, count


def edge_bwd_beta(hsum: torch.Tensor, g: GraphIndex, H: int) -> torch.Tensor:
    """dbeta [R] = sum over the edges of a relation of sum_h hsum[dst] (reference layer.py:313-318 backward)."""
    hsum = _f32c(hsum, "hsum")
    dev = hsum.device
    partB = torch.empty((max(g.n_chunks, 1),), dtype=torch.float32, device=dev)
    dbeta = torch.empty((g.R,), dtype=torch.float32, device=dev)
    with _on(dev):
        rc = _lib.load().relgat_layer_bwd_beta(_lib.ptr(hsum), _lib.ptr(g.rel_slot), _lib.ptr(g.csr_dst),
                                               _lib.ptr(g.chunk_lo), _lib.ptr(g.chunk_hi), _lib.ptr(g.rel_chunk_ptr),
                                               g.n_chunks, _lib.ptr(partB), _lib.ptr(dbeta), H, g.R, _stream(hsum))
    _lib.check(rc, "relgat_layer_bwd_beta")
    _count(2)
    return dbeta


def edge_bwd_rel(P, dz, hsum, g: GraphIndex, H: int, F: int, want_dbeta: bool = True):
    """Returns (dA [H,R,F], dbeta [R] or None)."""
    P = _feat(P, "P")
    dev = P.device
    C = H * F
    partA = torch.empty((max(g.n_chunks, 1), C), dtype=torch.float32, device=dev)
    partB = torch.empty((max(g.n_chunks, 1),), dtype=torch.float32, device=dev)
    dA = torch.empty((H, g.R, F), dtype=torch.float32, device=dev)
    dbeta = torch.empty((g.R,), dtype=torch.float32, device=dev) if want_dbeta else None
    with _on(dev):
        rc = _lib.load().relgat_layer_bwd_rel(
            _lib.ptr(P), int(P.dtype == torch.bfloat16), P.stride(0), _lib.ptr(dz), _lib.ptr(hsum), _lib.ptr(g.rel_slot),
            _lib.ptr(g.csr_src),
            _lib.ptr(g.csr_dst), _lib.ptr(g.chunk_lo), _lib.ptr(g.chunk_hi), _lib.ptr(g.rel_chunk_ptr),
            g.n_chunks, _lib.ptr(partA), _lib.ptr(partB), _lib.ptr(dA), _lib.ptr(dbeta), H, F, g.R, _stream(P))
    _lib.check(rc, "relgat_layer_bwd_rel")
    _count(2)
    return dA, dbeta


# ------------------------------------------------------------------------------------------
# scorers
# ------------------------------------------------------------------------------------------
def _ids(t: Optional[torch.Tensor], name: str) -> Optional[torch.Tensor]:
    if t is None:
        return None
    _lib.require_cuda(t)
    if t.dtype != torch.int64:
        raise TypeError(f"{name} must be int64")
    return t.contiguous()


CHECK_ID_RANGES = bool(int(os.environ.get("RELGAT_CHECK_IDS", "0")))  # debug: range-check ids (one device sync)


def _check_score_args(xs, src_idx, xd, dst_idx, rel_emb, rel_ids, n_transform=0):
    """Shape errors the reference would raise from torch (scorer.py:80-83, 176-186): every operand is [*, D] with
    D = rel_emb.size(1), un-indexed operands carry one row per triple, index vectors one entry per triple."""
    if rel_emb.dim() != 2 or xs.dim() != 2 or xd.dim() != 2:
        raise ValueError("scorer operands must be 2-D: xs/xd [*, D], rel_emb [R, D]")
    B, D = int(rel_ids.numel()), int(rel_emb.size(1))
    if xs.size(1) != D or xd.size(1) != D:
        raise ValueError(f"scorer width mismatch: xs [*, {xs.size(1)}], xd [*, {xd.size(1)}] vs rel_emb [*, {D}]")
    for nm, idx, x in (("src", src_idx, xs), ("dst", dst_idx, xd)):
        if idx is None:
            if x.size(0) < B:
                raise ValueError(f"{nm} rows: {x.size(0)} < {B} triples")
        elif idx.dim() != 1 or idx.numel() != B:
            raise ValueError(f"{nm}_idx must have one entry per triple ({B}), got {tuple(idx.shape)}")
    if not 0 <= n_transform <= B:
        raise ValueError(f"n_transform must be in [0, {B}]")
    if CHECK_ID_RANGES and B:
        for nm, idx, hi in (("src_idx", src_idx, xs.size(0)), ("dst_idx", dst_idx, xd.size(0)), ("rel_ids", rel_ids, rel_emb.size(0))):
            if idx is not None and (int(idx.min()) < 0 or int(idx.max()) >= hi):
                raise IndexError(f"{nm} out of range [0, {hi})")
    return B, D


def score_fwd(kind: str, normalize: bool, xs, src_idx, xd, dst_idx, rel_emb, rel_ids, *,
              n_transform: int = 0, want_src_vec: bool = False, want_dst_vec: bool = False):
    xs, xd, rel_emb = _f32c(xs, "xs"), _f32c(xd, "xd"), _f32c(rel_emb, "rel_emb")
    src_idx, dst_idx, rel_ids = _ids(src_idx, "src_idx"), _ids(dst_idx, "dst_idx"), _ids(rel_ids, "rel_ids")
    B, D = _check_score_args(xs, src_idx, xd, dst_idx, rel_emb, rel_ids, n_transform)
    dev = xs.device
    score = torch.empty((B,), dtype=torch.float32, device=dev)
    tr = torch.empty((n_transform, D), dtype=torch.float32, device=dev) if n_transform > 0 else None
    sv = torch.empty((B, D), dtype=torch.float32, device=dev) if want_src_vec else None
    dv = torch.empty((B, D), dtype=torch.float32, device=dev) if want_dst_vec else None
    with _on(dev):
        rc = _lib.load().relgat_score_fwd(
            SCORER_KIND[kind], int(normalize), _lib.ptr(xs), _lib.ptr(src_idx), _lib.ptr(xd), _lib.ptr(dst_idx),
            _lib.ptr(rel_emb), _lib.ptr(rel_ids), B, D, _lib.ptr(score), _lib.ptr(tr), n_transform,
            _lib.ptr(sv), _lib.ptr(dv), _stream(xs))
    _lib.check(rc, "relgat_score_fwd")
    _count(1)
    return score, tr, sv, dv


def score_bwd(kind: str, normalize: bool, xs, src_idx, xd, dst_idx, rel_emb, rel_ids, dscore, dtransform, *,
              want_src: bool = True, want_dst: bool = True, want_rel: bool = True):
    xs, xd, rel_emb = _f32c(xs, "xs"), _f32c(xd, "xd"), _f32c(rel_emb, "rel_emb")
    src_idx, dst_idx, rel_ids = _ids(src_idx, "src_idx"), _ids(dst_idx, "dst_idx"), _ids(rel_ids, "rel_ids")
    B, D = _check_score_args(xs, src_idx, xd, dst_idx, rel_emb, rel_ids)
    dev = xs.device
    if dscore is not None:
        dscore = _f32c(dscore, "dscore")
        if dscore.numel() != B:
            raise ValueError(f"dscore must have {B} entries, got {tuple(dscore.shape)}")
    n_tr = 0
    if dtransform is not None:
        dtransform = _f32c(dtransform, "dtransform")
        n_tr = int(dtransform.size(0))
        if dtransform.dim() != 2 or dtransform.size(1) != D or n_tr > B:
            raise ValueError(f"dtransform must be [<= {B}, {D}], got {tuple(dtransform.shape)}")
    mk = lambda w: torch.empty((B, D), dtype=torch.float32, device=dev) if w else None  # noqa: E731
    d_src, d_dst, d_rel = mk(want_src), mk(want_dst), mk(want_rel)
    with _on(dev):
        rc = _lib.load().relgat_score_bwd(
            SCORER_KIND[kind], int(normalize), _lib.ptr(xs), _lib.ptr(src_idx), _lib.ptr(xd), _lib.ptr(dst_idx),
            _lib.ptr(rel_emb), _lib.ptr(rel_ids), B, D, _lib.ptr(dscore), _lib.ptr(dtransform), n_tr,
            _lib.ptr(d_src), _lib.ptr(d_dst), _lib.ptr(d_rel), _stream(xs))
    _lib.check(rc, "relgat_score_bwd")
    _count(1)
    return d_src, d_dst, d_rel


def index_add_sorted(rows: torch.Tensor, keys: torch.Tensor, n_out: int, out: Optional[torch.Tensor] = None,
                     presorted=None, return_keys: bool = False, accumulate: Optional[bool] = None):
    """out[k] (+)= ordered sum of rows whose key == k.  ``keys`` int64 [M]; rows [M, D].
    ``presorted`` = (sorted_keys, perm) of a stable sort of ``keys`` when the caller cached it."""
    rows = _f32c(rows, "rows")
    M, D = rows.shape
    if accumulate is None:  # with accumulate=False only the rows named by ``keys`` are written (to the run sums)
        accumulate = out is not None
    if out is None:
        out = torch.zeros((n_out, D), dtype=torch.float32, device=rows.device)
    if M == 0:
        return (out, keys) if return_keys else out
    # plumbing: stable order = deterministic sum order
    sorted_keys, perm = presorted if presorted is not None else torch.sort(keys, stable=True)
    with _on(rows.device):
        rc = _lib.load().relgat_index_add_sorted(_lib.ptr(rows), _lib.ptr(perm), _lib.ptr(sorted_keys), _lib.ptr(out),
                                                 M, D, int(accumulate), _stream(rows))
    _lib.check(rc, "relgat_index_add_sorted")
    _count(1)
    return (out, sorted_keys) if return_keys else out


def margin_loss(score: torch.Tensor, B: int, K: int, margin: float, projection_layout: bool = False):
    """Returns (loss [1], dscore [B*(1+K)])."""
    score = _f32c(score, "score")
    if score.numel() != B * (1 + K):
        raise ValueError("score must have B*(1+K) entries")
    loss = torch.empty((1,), dtype=torch.float32, device=score.device)
    dscore = torch.empty_like(score)
    with _on(score.device):
        rc = _lib.load().relgat_margin_loss(_lib.ptr(score), B, K, float(margin), int(projection_layout),
                                            _lib.ptr(loss), _lib.ptr(dscore), _stream(score))
    _lib.check(rc, "relgat_margin_loss")
    _count(1)
    return loss, dscore


def gelu_layernorm_fwd(h: torch.Tensor, gamma: Optional[torch.Tensor], beta: Optional[torch.Tensor], eps: float):
    """y = LayerNorm(GELU(h)) over the last dim of h [M, D]; returns (y, mean [M], rstd [M])."""
    h = _f32c(h, "h")
    M, D = h.shape
    y = torch.empty_like(h)
    mean = torch.empty((M,), dtype=torch.float32, device=h.device)
    rstd = torch.empty((M,), dtype=torch.float32, device=h.device)
    with _on(h.device):
        rc = _lib.load().relgat_gelu_layernorm_fwd(_lib.ptr(h), _lib.ptr(None if gamma is None else _f32c(gamma, "gamma")),
                                                   _lib.ptr(None if beta is None else _f32c(beta, "beta")), _lib.ptr(y),
                                                   _lib.ptr(mean), _lib.ptr(rstd), M, D, float(eps), _stream(h))
    _lib.check(rc, "relgat_gelu_layernorm_fwd")
    _count(1)
    return y, mean, rstd


def gelu_layernorm_bwd(dy: torch.Tensor, h: torch.Tensor, gamma: Optional[torch.Tensor], mean: torch.Tensor,
                       rstd: torch.Tensor, want_params: bool = True):
    """Returns (dh [M, D], dgamma [D] or None, dbeta [D] or None)."""
    dy, h = _f32c(dy, "dy"), _f32c(h, "h")
    M, D = h.shape
    lib = _lib.load()
    groups = max(int(lib.relgat_gelu_layernorm_groups(M)), 1)
    dev = h.device
    dh = torch.empty_like(h)
    part_g = torch.empty((groups, D), dtype=torch.float32, device=dev)
    part_b = torch.empty((groups, D), dtype=torch.float32, device=dev)
    dgamma = torch.empty((D,), dtype=torch.float32, device=dev) if want_params else None
    dbeta = torch.empty((D,), dtype=torch.float32, device=dev) if want_params else None
    with _on(dev):
        rc = lib.relgat_gelu_layernorm_bwd(_lib.ptr(dy), _lib.ptr(h), _lib.ptr(None if gamma is None else _f32c(gamma, "gamma")),
                                           _lib.ptr(mean), _lib.ptr(rstd), _lib.ptr(dh), _lib.ptr(part_g), _lib.ptr(part_b),
                                           _lib.ptr(dgamma), _lib.ptr(dbeta), M, D, _stream(h))
    _lib.check(rc, "relgat_gelu_layernorm_bwd")
    _count(2)
    return dh, dgamma, dbeta


RANK_LOSS_KIND = {"margin": 0, "self_adversarial_loss": 1}


def rank_loss(pos: torch.Tensor, neg: torch.Tensor, kind: str, margin: float = 1.0, alpha: float = 1.0,
              sanitize: bool = False):
    """Ranking loss of reference core/loss/relgat_loss.py:32-71 on pos [B] and neg [B, K] (any strides: the two
    negative layouts of the reference trainer are views of the flat score vector).  Returns (loss [1], dpos [B],
    dneg with neg's strides).  ``sanitize``: nan_to_num(nan=0, +-inf=+-1e9) as in trainer:584, 647-648."""
    _lib.require_cuda(pos, neg)
    if pos.dtype != torch.float32 or neg.dtype != torch.float32:
        raise TypeError("rank_loss: scores must be float32")
    if pos.dim() != 1 or neg.dim() != 2 or neg.size(0) != pos.size(0):
        raise ValueError(f"rank_loss: pos [B] and neg [B, K] expected, got {tuple(pos.shape)} / {tuple(neg.shape)}")
    if kind not in RANK_LOSS_KIND:
        raise ValueError(f"unknown ranking loss {kind!r}")
    B, K = int(neg.size(0)), int(neg.size(1))
    pos = pos.contiguous()
    dneg = torch.empty_like(neg)  # keeps the strides of a dense view
    if dneg.stride() != neg.stride():
        neg = neg.contiguous()
        dneg = torch.empty_like(neg)
    loss = torch.empty((1,), dtype=torch.float32, device=pos.device)
    dpos = torch.empty_like(pos)
    with _on(pos.device):
        rc = _lib.load().relgat_rank_loss(_lib.ptr(pos), _lib.ptr(neg), B, K, neg.stride(0) if K else 0,
                                          neg.stride(1) if K else 0, RANK_LOSS_KIND[kind], float(margin), float(alpha),
                                          int(sanitize), _lib.ptr(loss), _lib.ptr(dpos), _lib.ptr(dneg), _stream(pos))
    _lib.check(rc, "relgat_rank_loss")
    _count(1)
    return loss, dpos, dneg


def recon_loss(tr: torch.Tensor, dst: torch.Tensor, negdst: Optional[torch.Tensor], w_pos: float, w_neg: float,
               w_mse: float):
    """Reconstruction terms of reference core/loss/multi_objective_loss.py:47-83 (cosine.py:4-13, mse.py:4-10):
    tr = f_r(A) [B, D], dst [B, D], negdst [K, B, D] (any outer strides, unit inner stride; None = no negatives).
    Returns (values [3] = (cos_pos, cos_neg, mse) losses, d_tr, d_dst, d_negdst) where the gradients are those of
    w_pos*cos_pos + w_neg*(1 - cos_neg) + w_mse*mse."""
    _lib.require_cuda(tr, dst)
    tr, dst = _f32c(tr, "transformed_src"), _f32c(dst, "dst_vec")
    if tr.dim() != 2 or tr.shape != dst.shape:
        raise ValueError(f"recon_loss: transformed_src and dst_vec must both be [B, D], got {tuple(tr.shape)} / {tuple(dst.shape)}")
    B, D = tr.shape
    K, sb, sk = 0, 0, 0
    d_neg = None
    if negdst is not None and negdst.numel() > 0:
        _lib.require_cuda(negdst)
        if negdst.dim() != 3 or negdst.size(1) != B or negdst.size(2) != D or negdst.dtype != torch.float32:
            raise ValueError(f"recon_loss: neg_dst_vec must be float32 [K, {B}, {D}], got {tuple(negdst.shape)}")
        if negdst.stride(2) != 1:
            negdst = negdst.contiguous()
        d_neg = torch.empty_like(negdst)
        if d_neg.stride() != negdst.stride():
            negdst = negdst.contiguous()
            d_neg = torch.empty_like(negdst)
        K, sk, sb = int(negdst.size(0)), negdst.stride(0), negdst.stride(1)
    else:
        negdst = None
    dev = tr.device
    values = torch.empty((3,), dtype=torch.float32, device=dev)
    partial = torch.empty((max(B, 1), 3), dtype=torch.float32, device=dev)
    d_tr, d_dst = torch.empty_like(tr), torch.empty_like(dst)
    with _on(dev):
        rc = _lib.load().relgat_recon_loss(_lib.ptr(tr), _lib.ptr(dst), _lib.ptr(negdst), B, K, D, sb, sk, float(w_pos),
                                           float(w_neg), float(w_mse), _lib.ptr(values), _lib.ptr(partial), _lib.ptr(d_tr),
                                           _lib.ptr(d_dst), _lib.ptr(d_neg), _stream(tr))
    _lib.check(rc, "relgat_recon_loss")
    _count(2)
    return values, d_tr, d_dst, d_neg


def gather_plane_rows(plane: torch.Tensor, ids: torch.Tensor) -> torch.Tensor:
    """rows ``ids`` of a bf16 plane [N, D] as a new bf16 [n, D] (byte copy of whole rows; the input rows of a batch's
    first receptive-field block)."""
    _lib.require_cuda(plane, ids)
    if plane.dtype != torch.bfloat16 or plane.dim() != 2 or not plane.is_contiguous() or ids.dtype != torch.int64:
        raise TypeError("gather_plane_rows: contiguous 2-D bfloat16 plane and int64 ids expected")
    n, D = int(ids.numel()), int(plane.size(1))
    out = torch.empty((n, D), dtype=torch.bfloat16, device=plane.device)
    if n == 0:
        return out
    if D % 2:  # odd width: rows are not whole 32-bit words
        return torch.index_select(plane, 0, ids)
    pull_rows(plane.view(torch.float32), ids, out.view(torch.float32))
    return out


def pull_rows(table: torch.Tensor, ids: torch.Tensor, out: torch.Tensor, out_ids: Optional[torch.Tensor] = None) -> torch.Tensor:
    """out[i] = table[ids[i]] (rows of a mapped peer table -> local rows); with ``out_ids`` the rows are scattered:
    out[out_ids[i]] = table[ids[i]].  ``table`` / ``out``: fp32 with unit inner stride, any number of trailing dims
    (flattened); ``ids`` / ``out_ids`` int64."""
    _lib.require_cuda(table, ids, out)
    if table.dtype not in (torch.float32, torch.bfloat16) or out.dtype != torch.float32 or ids.dtype != torch.int64:
        raise TypeError("pull_rows: table must be float32 or bfloat16, out float32 and ids int64")
    n = int(ids.numel())
    D = int(table[0].numel()) if table.size(0) else int(out[0].numel())
    if out_ids is None and out.size(0) != n:
        raise ValueError("pull_rows: out must have len(ids) rows")
    if out_ids is not None and (out_ids.dtype != torch.int64 or out_ids.numel() != n):
        raise ValueError("pull_rows: out_ids must be int64 with len(ids) entries")
    if (out.size(0) and int(out[0].numel()) != D) or not table.is_contiguous() or not out.is_contiguous():
        raise ValueError("pull_rows: out must be a contiguous tensor with the table's row shape")
    if n == 0:
        return out
    fn = _lib.load().relgat_pull_rows_bf16 if table.dtype == torch.bfloat16 else _lib.load().relgat_pull_rows
    with _on(out.device):
        rc = fn(_lib.ptr(table), D, _lib.ptr(ids.contiguous()),
                _lib.ptr(None if out_ids is None else out_ids.contiguous()), n, D,
                _lib.ptr(out), D, sm_count(out.device), _stream(out))
    _lib.check(rc, "relgat_pull_rows")
    _count(1)
    return out
