"""ctypes binding of libwf_b200.so (include/wf_b200.h) over torch tensors.

PyTorch is used here only for device memory and streams: every function takes CUDA tensors,
passes ``data_ptr()`` and the current stream to the C ABI, and raises ``WfError`` when the
library reports a failure.  There is no CPU or PyTorch fallback: if the shared library is
missing or the tensors are not on a CUDA device the call fails loudly.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libwf_b200.so")

EPI_BF16, EPI_GELU_BF16, EPI_RESID_F32, EPI_F32_OF_BF16, EPI_RESID_BF16 = 0, 1, 2, 3, 4


class WfError(RuntimeError):
    pass


_lib = None
launches = 0     # kernels launched through this binding (bench.py's gpu_launches)
trace = None             # bench.py (WF_TRACE=1): a dict name -> [(start event, end event)] filled by ``phase``


class phase:
    """``with lib.phase("vae.decode"):`` - CUDA-event bracket on the current stream when ``lib.trace`` is a dict (a development
    breakdown of a step; costs nothing otherwise)."""

    def __init__(self, name):
        self.name = name

    def __enter__(self):
        if trace is not None:
            self.ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            self.ev[0].record()
        return self

    def __exit__(self, *a):
        if trace is not None:
            self.ev[1].record()
            trace.setdefault(self.name, []).append(self.ev)
        return False


def trace_summary():
    """name -> (calls, total ms) of the recorded phases (synchronises)."""
    torch.cuda.synchronize()
    return {k: (len(v), sum(a.elapsed_time(b) for a, b in v)) for k, v in (trace or {}).items()}


timed_attention = None   # bench.py: a list here collects (start event, end event, Lq, Lk) of every self-attention launch

_vp, _i, _f, _ll, _u = C.c_void_p, C.c_int, C.c_float, C.c_longlong, C.c_uint
_SIGNATURES = {
    "wf_gemm_bf16": [_vp, _i, _vp, _i, _vp, _vp, _i, _vp, _i, _i, _i, _i, _i, _vp],
    "wf_attention_bf16": [_vp, _i, _vp, _i, _vp, _i, _vp, _i, _vp, _i, _i, _i, _i, _f, _vp],
    "wf_bsa_mean_pool": [_vp, _i, _vp, _i, _i, _i, _i, _i, _i, _i, _vp],
    "wf_bsa_select_topk": [_vp, _vp, _vp, _i, _i, _i, _i, _vp],
    "wf_bsa_select_cdf": [_vp, _vp, _vp, _vp, _i, _i, _i, _f, _i, _vp],
    "wf_attention_bsa_bf16": [_vp, _i, _vp, _i, _vp, _i, _vp, _i, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i, _f, _vp],
    "wf_layer_norm": [_vp, _i, _i, _vp, _i, _i, _vp, _vp, _vp, _vp, _i, _i, _f, _i, _i, _vp],
    "wf_rms_norm_rope": [_vp, _i, _vp, _vp, _i, _i, _f, _vp],
    "wf_qkv_norm_rope_scatter": [_vp, _i, _vp, _vp, _vp, _i, _i, _f, _vp, _i, _i, _i, _vp],
    "wf_attention_bf16_peers": [_vp, _i, _vp, _i, _vp, _i, _vp, _i, _i, _i, _i, _i, _i, _f, _vp],
    "wf_farneback_u8": [_vp, _i, _i, _i, _i, _i, _i, _vp, _vp, _vp],
    "wf_flow_metrics": [_vp, _vp, _vp, _i, _ll, _i, _vp],
    "wf_peer_alloc": [_ll, _vp, _vp],
    "wf_peer_open": [_vp, _vp],
    "wf_peer_close": [_vp],
    "wf_peer_free": [_vp],
    "wf_patchify": [_vp, _vp, _i, _i, _i, _i, _vp],
    "wf_dit_head": [_vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _i, _vp, _i, _i, _i, _f, _i, _i, _i, _vp],
    "wf_rms_norm_head_rope": [_vp, _i, _vp, _vp, _ll, _i, _f, _vp],
    "wf_swiglu_bf16": [_vp, _vp, _ll, _i, _vp],
    "wf_timestep_embedding_f32": [_vp, _vp, _i, _i, _vp],
    "wf_small_gemm_f32": [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp],
    "wf_gemv_f32": [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp],
    "wf_gelu_erf_bf16": [_vp, _ll, _vp],
    "wf_time_sinusoid": [_vp, _vp, _i, _vp],
    "wf_add_bcast_f32": [_vp, _vp, _vp, _ll, _i, _vp],
    "wf_cfg_combine": [_vp, _vp, _vp, _i, _f, _ll, _vp],
    "wf_x0_convert": [_vp, _i, _vp, _i, _vp, _f, _ll, _vp],
    "wf_unip_update": [_vp, _i, _vp, _i, _vp, _i, _vp, _i, _f, _f, _f, _i, _f, _ll, _vp],
    "wf_renoise": [_vp, _i, _vp, _vp, _f, _f, _ll, _vp],
    "wf_dsg": [_vp, _vp, _vp, _i, _f, _ll, _vp, _vp, _vp],
    "wf_flf_blend": [_vp, _vp, _vp, _vp, _i, _ll, _vp],
    "wf_latent_denorm": [_vp, _i, _vp, _vp, _vp, _i, _ll, _vp],
    "wf_latent_norm_replace": [_vp, _vp, _i, _vp, _vp, _vp, _u, _i, _ll, _vp],
    "wf_quantise_u8": [_vp, _i, _vp, _ll, _i, _vp, _vp],
    "wf_refine_upsample": [_vp, _i, _i, _i, _vp, _i, _i, _i, _i, _i, _vp],
    "wf_cfg_zero": [_vp, _vp, _vp, _f, _ll, _vp, _vp, _vp],
    "wf_conv_tf32": [_vp, _i, _i, _i, _i, _vp, _vp, _i, _i, _vp, _i, _i, _i, _i, _i, _vp, _i, _i, _i, _i, _i, _i, _i, _i, _i,
                     _vp, _i, _ll, _i, _i, _vp, _vp, _i, _vp],
    "wf_debug_conv_profile": [_vp],
    "wf_attention_small_bf16": [_vp, _i, _vp, _i, _vp, _i, _vp, _i, _vp, _vp, _i, _i, _i, _i, _i, _f, _i, _vp],
    "wf_geglu_bf16": [_vp, _vp, _ll, _i, _vp],
    "wf_soften_mask": [_vp, _vp, _i, _i, _i, _i, _i, _vp, _vp],
    "wf_clip_from_u8": [_vp, _vp, _ll, _vp],
    "wf_rms_norm_cl": [_vp, _i, _vp, _i, _vp, _ll, _i, _i, _i, _vp],
    "wf_planar_to_cl": [_vp, _vp, _ll, _i, _i, _i, _vp],
    "wf_round_tf32": [_vp, _vp, _ll, _vp],
    "wf_cl_to_planar": [_vp, _vp, _ll, _i, _i, _vp],
    "wf_space_to_depth": [_vp, _vp, _i, _i, _i, _i, _vp],
    "wf_softmax_rows": [_vp, _i, _i, _i, _f, _vp],
    "wf_transpose_f32": [_vp, _vp, _i, _i, _i, _i, _vp],
    "wf_split_tf32": [_vp, _vp, _vp, _ll, _vp],
}
_PLAIN = {"wf_last_error": (C.c_char_p, []), "wf_abi_version": (_i, []), "wf_sm_count": (_i, []),
          "wf_dsg_workspace_bytes": (_ll, []), "wf_quantise_workspace_bytes": (_ll, []),
          "wf_farneback_workspace_bytes": (_ll, [_i, _i, _i, _i])}


def exported_symbols():
    """Every symbol include/wf_b200.h declares (checked by the CPU test-suite)."""
    return sorted(list(_SIGNATURES) + list(_PLAIN))


def load(path: str = LIB_PATH):
    """Load the shared library (no GPU needed) and declare the prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(path):
        raise WfError(f"{path} is missing - run `python -c 'import __graft_entry__ as g; g.build()'` "
                      "(worldforge_b200 has no CPU fallback)")
    lib = C.CDLL(path)
    for name, args in _SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = _i
    for name, (res, args) in _PLAIN.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = res
    if lib.wf_abi_version() != 1:
        raise WfError("libwf_b200.so ABI version mismatch")
    _lib = lib
    return lib


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _p(t: Optional[torch.Tensor]):
    if t is None:
        return None
    if not t.is_cuda:
        raise WfError("worldforge_b200 kernels take CUDA tensors only (no CPU fallback)")
    return t.data_ptr()


def _call(name: str, *args):
    global launches
    lib = load()
    rc = getattr(lib, name)(*args)
    if rc != 0:
        raise WfError(f"{name} failed ({rc}): {lib.wf_last_error().decode()}")
    launches += 1


def _is_bf16(t: torch.Tensor) -> int:
    if t.dtype == torch.bfloat16:
        return 1
    if t.dtype == torch.float32:
        return 0
    raise WfError(f"unsupported dtype {t.dtype}")


# ---------------------------------------------------------------------------- DiT kernels

def gemm_bf16(a: torch.Tensor, w: torch.Tensor, bias: Optional[torch.Tensor], out: torch.Tensor,
              epilogue: int = EPI_BF16, gate: Optional[torch.Tensor] = None, gate_rows: int = 0):
    """out = epilogue(a[M,K] @ w[N,K]^T + bias); a/w bf16 with contiguous K, out [M, N] view (row stride ldo)."""
    M, K = a.shape
    N = w.shape[0]
    assert w.shape[1] == K and a.stride(1) == 1 and w.stride(1) == 1 and out.stride(1) == 1
    assert a.dtype == torch.bfloat16 and w.dtype == torch.bfloat16
    assert out.shape == (M, N)
    want = torch.bfloat16 if epilogue in (EPI_BF16, EPI_GELU_BF16, EPI_RESID_BF16) else torch.float32
    assert out.dtype == want, f"epilogue {epilogue} writes {want}"
    assert bias is None or (bias.dtype == torch.bfloat16 and bias.numel() == N)
    assert gate is None or (gate.dtype == torch.float32 and gate.is_contiguous() and
                            gate.numel() == (N if gate_rows == 0 else N * ((M + gate_rows - 1) // gate_rows)))
    _call("wf_gemm_bf16", _p(a), a.stride(0), _p(w), w.stride(0), _p(bias), _p(out), out.stride(0), _p(gate), gate_rows,
          M, N, K, epilogue, _stream())
    return out


def attention_bf16(q, k, v, out, heads: int, add_in=None, softmax_scale: Optional[float] = None):
    """q [Lq, >=heads*128], k/v [Lk, ...], out [Lq, ...]: bf16 2-D views with contiguous columns."""
    for t in (q, k, v, out):
        assert t.dtype == torch.bfloat16 and t.dim() == 2 and t.stride(1) == 1
    scale = softmax_scale if softmax_scale is not None else 128 ** -0.5
    ev = None
    if timed_attention is not None and k.shape[0] >= 2048 and 2 * q.shape[0] >= k.shape[0]:     # self-attention, not the 769 / 512-key cross-attention
        ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        ev[0].record()          # on the current stream = the stream the kernel is launched on
    _call("wf_attention_bf16", _p(q), q.stride(0), _p(k), k.stride(0), _p(v), v.stride(0), _p(out), out.stride(0),
          _p(add_in), add_in.stride(0) if add_in is not None else 0, q.shape[0], k.shape[0], heads, scale, _stream())
    if ev is not None:
        ev[1].record()
        timed_attention.append((ev[0], ev[1], q.shape[0], k.shape[0]))
    return out


def bsa_mean_pool(x, grid, chunk, heads: int):
    """x [T*H*W, >=heads*128] bf16 in (t,h,w) order -> [heads, chunks, 128] bf16 chunk means."""
    (T, H, W), (ct, ch, cw) = grid, chunk
    assert x.dtype == torch.bfloat16 and x.dim() == 2 and x.stride(1) == 1 and x.shape[0] == T * H * W
    out = torch.empty(heads, (T // ct) * (H // ch) * (W // cw), 128, dtype=torch.bfloat16, device=x.device)
    _call("wf_bsa_mean_pool", _p(x), x.stride(0), _p(out), T, H, W, ct, ch, cw, heads, _stream())
    return out


def bsa_select_topk(q_cmp, k_cmp, n_sel: int):
    """[heads, Nq, 128], [heads, Nk, 128] bf16 -> int32 [heads, Nq, n_sel], ascending."""
    assert q_cmp.is_contiguous() and k_cmp.is_contiguous() and q_cmp.dtype == k_cmp.dtype == torch.bfloat16
    heads, Nq, _ = q_cmp.shape
    idx = torch.empty(heads, Nq, n_sel, dtype=torch.int32, device=q_cmp.device)
    _call("wf_bsa_select_topk", _p(q_cmp), _p(k_cmp), _p(idx), Nq, k_cmp.shape[1], heads, n_sel, _stream())
    return idx


def bsa_select_cdf(q_cmp, k_cmp, cdf_threshold: float, n_floor: int = 0):
    """[heads, Nq, 128], [heads, Nk, 128] bf16 -> (int32 [heads, Nq, Nk] sorted by weight, int32 [heads, Nq] selected counts)."""
    assert q_cmp.is_contiguous() and k_cmp.is_contiguous() and q_cmp.dtype == k_cmp.dtype == torch.bfloat16
    heads, Nq, _ = q_cmp.shape
    Nk = k_cmp.shape[1]
    idx = torch.empty(heads, Nq, Nk, dtype=torch.int32, device=q_cmp.device)
    lens = torch.empty(heads, Nq, dtype=torch.int32, device=q_cmp.device)
    _call("wf_bsa_select_cdf", _p(q_cmp), _p(k_cmp), _p(idx), _p(lens), Nq, Nk, heads, float(cdf_threshold), int(n_floor), _stream())
    return idx, lens


def attention_bsa_bf16(q, k, v, out, heads: int, block_idx, block_lens, grid_q, grid_k, chunk,
                       softmax_scale: Optional[float] = None):
    """Block-sparse attention in (t,h,w) token order; block_idx int32 [heads, q_chunks, max_sel], block_lens int32 or None."""
    for t in (q, k, v, out):
        assert t.dtype == torch.bfloat16 and t.dim() == 2 and t.stride(1) == 1
    (Tq, H, W), (Tk, Hk, Wk), (ct, ch, cw) = grid_q, grid_k, chunk
    assert (H, W) == (Hk, Wk) and q.shape[0] == Tq * H * W and k.shape[0] == v.shape[0] == Tk * H * W
    assert block_idx.dtype == torch.int32 and block_idx.is_contiguous() and block_idx.shape[0] == heads
    assert block_idx.shape[1] == (Tq // ct) * (H // ch) * (W // cw)
    if block_lens is not None:
        assert block_lens.dtype == torch.int32 and block_lens.is_contiguous() and block_lens.shape == block_idx.shape[:2]
    scale = softmax_scale if softmax_scale is not None else 128 ** -0.5
    _call("wf_attention_bsa_bf16", _p(q), q.stride(0), _p(k), k.stride(0), _p(v), v.stride(0), _p(out), out.stride(0),
          _p(block_idx), _p(block_lens), block_idx.shape[2], Tq, Tk, H, W, ct, ch, cw, heads, scale, _stream())
    return out


def layer_norm(x, out, eps: float, scale=None, shift=None, weight=None, bias=None, round_norm_bf16: bool = False,
               rows_per_group: int = 0):
    rows, D = x.shape
    assert x.stride(1) == 1 and out.stride(1) == 1 and out.shape == x.shape
    groups = 1 if rows_per_group == 0 else (rows + rows_per_group - 1) // rows_per_group
    for t in (scale, shift):
        assert t is None or (t.dtype == torch.float32 and t.numel() == D * groups and t.is_contiguous())
    for t in (weight, bias):
        assert t is None or (t.dtype == torch.float32 and t.numel() == D and t.is_contiguous())
    _call("wf_layer_norm", _p(x), x.stride(0), _is_bf16(x), _p(out), out.stride(0), _is_bf16(out), _p(scale), _p(shift),
          _p(weight), _p(bias), rows, D, eps, int(round_norm_bf16), rows_per_group, _stream())
    return out


def attention_bf16_peers(q, k, v, peer_ptrs, rows_per_peer: int, ldo: int, heads: int, softmax_scale: Optional[float] = None):
    """Attention whose output row r is stored into peer_ptrs[r // rows_per_peer] (raw device pointers of the ranks'
    buffers, column offset of this rank's head block included) at row r % rows_per_peer, leading dimension ldo."""
    for t in (q, k, v):
        assert t.dtype == torch.bfloat16 and t.dim() == 2 and t.stride(1) == 1
    scale = softmax_scale if softmax_scale is not None else 128 ** -0.5
    ev = None
    if timed_attention is not None and q.shape[0] == k.shape[0]:
        ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
        ev[0].record()
    tab = (C.c_void_p * len(peer_ptrs))(*peer_ptrs)
    _call("wf_attention_bf16_peers", _p(q), q.stride(0), _p(k), k.stride(0), _p(v), v.stride(0), C.cast(tab, _vp), len(peer_ptrs),
          rows_per_peer, ldo, q.shape[0], k.shape[0], heads, scale, _stream())
    if ev is not None:
        ev[1].record()
        timed_attention.append((ev[0], ev[1], q.shape[0], k.shape[0]))


def qkv_norm_rope_scatter(qkv, weight_q, weight_k, rope, eps: float, peer_ptrs, ld_dst: int, row0: int):
    """qkv bf16 [rows, 3*D] -> RMSNorm + RoPE of q and k, v copied, each head written to the peer that owns it."""
    rows, W3 = qkv.shape
    D = W3 // 3
    assert qkv.dtype == torch.bfloat16 and qkv.stride(1) == 1
    assert rope.dtype == torch.float64 and rope.shape == (rows, 64, 2) and rope.is_contiguous()
    tab = (C.c_void_p * len(peer_ptrs))(*peer_ptrs)
    _call("wf_qkv_norm_rope_scatter", _p(qkv), qkv.stride(0), _p(weight_q), _p(weight_k), _p(rope), rows, D, eps, C.cast(tab, _vp),
          len(peer_ptrs), ld_dst, row0, _stream())


class PeerBuffer:
    """Device memory other ranks of the box can map (cudaMalloc + CUDA IPC): ``ptr`` / ``handle`` here, ``open`` there."""

    def __init__(self, nbytes: int):
        lib = load()
        ptr, handle = C.c_void_p(), C.create_string_buffer(64)
        rc = lib.wf_peer_alloc(nbytes, C.byref(ptr), handle)
        if rc != 0:
            raise WfError(f"wf_peer_alloc failed ({rc}): {lib.wf_last_error().decode()}")
        self.ptr, self.handle, self.nbytes = ptr.value, handle.raw, nbytes

    def free(self):
        """Release the allocation (idempotent).  The caller makes sure no kernel or peer still uses it (PeerSequenceParallel.close
        synchronises the ranks first)."""
        if self.ptr:
            load().wf_peer_free(C.c_void_p(self.ptr))
            self.ptr = 0

    def __del__(self):
        try:
            self.free()
        except Exception:        # noqa: BLE001 - interpreter shutdown
            pass

    @staticmethod
    def close_mapping(ptr: int):
        """Unmap a peer's allocation opened with ``open``."""
        if ptr:
            load().wf_peer_close(C.c_void_p(ptr))

    def tensor(self, shape, dtype, device):
        """A torch view of the allocation (no ownership)."""
        n = 1
        for d in shape:
            n *= d
        iface = {"shape": (self.nbytes,), "typestr": "|u1", "data": (self.ptr, False), "version": 2}
        holder = type("_Raw", (), {"__cuda_array_interface__": iface})()
        flat = torch.as_tensor(holder, device=device)
        return flat[:n * torch.empty((), dtype=dtype).element_size()].view(dtype).view(*shape)

    @staticmethod
    def open(handle: bytes) -> int:
        lib = load()
        ptr = C.c_void_p()
        rc = lib.wf_peer_open(C.create_string_buffer(handle, 64), C.byref(ptr))
        if rc != 0:
            raise WfError(f"wf_peer_open failed ({rc}): {lib.wf_last_error().decode()}")
        return ptr.value


def rms_norm_rope_(x, weight, eps: float, rope=None):
    """In place on a bf16 [rows, D] view; rope: float64 [rows, 64, 2] or None."""
    rows, D = x.shape
    assert x.dtype == torch.bfloat16 and x.stride(1) == 1
    assert weight.dtype == torch.float32 and weight.numel() == D
    assert rope is None or (rope.dtype == torch.float64 and rope.shape == (rows, 64, 2) and rope.is_contiguous())
    _call("wf_rms_norm_rope", _p(x), x.stride(0), _p(weight), _p(rope), rows, D, eps, _stream())
    return x


def patchify(hidden, cols):
    C_, F_, H, W = hidden.shape
    assert hidden.dtype == torch.bfloat16 and hidden.is_contiguous()
    assert cols.dtype == torch.bfloat16 and cols.is_contiguous() and cols.shape == (F_ * (H // 2) * (W // 2), 4 * C_)
    _call("wf_patchify", _p(hidden), _p(cols), C_, F_, H, W, _stream())
    return cols


def dit_head(x, scale, shift, w, b, out, grid, eps: float, tok_offset: int = 0, rows_per_group: int = 0, round_bf16: bool = False):
    L, D = x.shape
    F_, GH, GW = grid
    cout = w.shape[0] // 4
    assert x.stride(1) == 1 and w.is_contiguous() and w.dtype == torch.float32
    assert out.dtype == torch.float32 and out.is_contiguous() and out.shape == (cout, F_, 2 * GH, 2 * GW)
    _call("wf_dit_head", _p(x), _is_bf16(x), x.stride(0), L, D, _p(scale), _p(shift), _p(w), _p(b), cout, _p(out), F_, GH, GW,
          eps, tok_offset, rows_per_group, int(round_bf16), _stream())
    return out


def gemv_f32(w, x, b, out, silu_in=False, silu_out=False):
    N, K = w.shape
    assert w.dtype == torch.float32 and w.is_contiguous() and x.dtype == torch.float32 and x.numel() == K
    assert out.dtype == torch.float32 and out.numel() == N
    _call("wf_gemv_f32", _p(w), _p(x), _p(b), _p(out), N, K, int(silu_in), int(silu_out), _stream())
    return out


def gelu_erf_bf16_(x):
    assert x.dtype == torch.bfloat16 and x.is_contiguous()
    _call("wf_gelu_erf_bf16", _p(x), x.numel(), _stream())
    return x


def rms_norm_head_rope_(x, gain, eps: float, rope=None):
    """In place on a bf16 [rows, heads*128] view; gain bf16 [128]; rope fp32 [rows, 64, 2] or None."""
    rows, W = x.shape
    assert x.dtype == torch.bfloat16 and x.stride(1) == 1 and W % 128 == 0
    assert gain.dtype == torch.bfloat16 and gain.numel() == 128
    assert rope is None or (rope.dtype == torch.float32 and rope.shape == (rows, 64, 2) and rope.is_contiguous())
    _call("wf_rms_norm_head_rope", _p(x), x.stride(0), _p(gain), _p(rope), rows, W // 128, eps, _stream())
    return x


def attention_small(q, k, v, out, heads: int, head_dim: int, mode: int = 1, scale: float = 1.0, n_valid: Optional[int] = None,
                    bias_emb=None, bias_bucket=None):
    """Attention of the encoders (head_dim 64 / 80, <= 1024 keys); see wf_attention_small_bf16 in include/wf_b200.h."""
    for t in (q, k, v, out):
        assert t.dtype == torch.bfloat16 and t.dim() == 2 and t.stride(1) == 1
    Lq, Lk = q.shape[0], k.shape[0]
    if bias_emb is not None:
        assert bias_emb.dtype == torch.bfloat16 and bias_emb.is_contiguous() and bias_emb.shape[1] == heads
        assert bias_bucket.dtype == torch.int32 and bias_bucket.numel() == 2 * Lk - 1 and bias_bucket.is_contiguous()
    _call("wf_attention_small_bf16", _p(q), q.stride(0), _p(k), k.stride(0), _p(v), v.stride(0), _p(out), out.stride(0),
          _p(bias_emb), _p(bias_bucket), Lq, Lk, Lk if n_valid is None else n_valid, heads, head_dim, float(scale), mode, _stream())
    return out


def geglu_bf16(inp, out):
    """inp bf16 [rows, 2F] = [gate | fc1] -> out [rows, F] = fc1 * GELU_tanh(gate) (T5FeedForward, bf16 intermediates)."""
    rows, F2 = inp.shape
    assert inp.dtype == torch.bfloat16 and inp.is_contiguous() and out.is_contiguous() and out.shape == (rows, F2 // 2)
    _call("wf_geglu_bf16", _p(inp), _p(out), rows, F2 // 2, _stream())
    return out


def swiglu_bf16(inp, out):
    rows, F2 = inp.shape
    assert inp.dtype == torch.bfloat16 and inp.is_contiguous() and out.is_contiguous() and out.shape == (rows, F2 // 2)
    _call("wf_swiglu_bf16", _p(inp), _p(out), rows, F2 // 2, _stream())
    return out


def timestep_embedding_f32(t, out):
    assert t.dtype == torch.float32 and out.dtype == torch.float32 and out.is_contiguous() and out.shape[0] == t.numel()
    _call("wf_timestep_embedding_f32", _p(t.contiguous()), _p(out), t.numel(), out.shape[1], _stream())
    return out


def small_gemm_f32(x, w, b, out, silu_in=False):
    T, K = x.shape
    R = w.shape[0]
    assert x.dtype == torch.float32 and x.is_contiguous() and w.dtype == torch.bfloat16 and w.is_contiguous() and w.shape[1] == K
    assert out.dtype == torch.float32 and out.is_contiguous() and out.shape == (T, R)
    _call("wf_small_gemm_f32", _p(x), _p(w), _p(b), _p(out), T, R, K, int(silu_in), _stream())
    return out


def time_sinusoid(timestep, out):
    assert timestep.dtype == torch.int64 and out.dtype == torch.float32 and out.is_contiguous()
    _call("wf_time_sinusoid", _p(timestep), _p(out), out.numel(), _stream())
    return out


def add_bcast_f32(a, b, out):
    assert a.dtype == b.dtype == out.dtype == torch.float32 and a.is_contiguous() and b.is_contiguous() and out.is_contiguous()
    inner = b.numel()
    assert a.numel() % inner == 0 and out.numel() == a.numel()
    _call("wf_add_bcast_f32", _p(a), _p(b), _p(out), a.numel() // inner, inner, _stream())
    return out


# -------------------------------------------------------------------------- sampler kernels

def _flat_ok(*ts):
    n = ts[0].numel()
    for t in ts:
        assert t.is_contiguous() and t.numel() == n
    assert n % 4 == 0, "element count must be a multiple of 4"
    return n


def cfg_combine(cond, uncond, scale: float):
    assert cond.dtype == uncond.dtype
    out = torch.empty_like(cond)
    n = _flat_ok(cond, uncond)
    _call("wf_cfg_combine", _p(cond), _p(uncond), _p(out), _is_bf16(cond), float(scale), n, _stream())
    return out


def x0_convert(sample, v, sigma: float):
    n = _flat_ok(sample, v)
    dt = torch.bfloat16 if (sample.dtype == torch.bfloat16 and v.dtype == torch.bfloat16) else torch.float32
    out = torch.empty(sample.shape, dtype=dt, device=sample.device)
    _call("wf_x0_convert", _p(sample), _is_bf16(sample), _p(v), _is_bf16(v), _p(out), float(sigma), n, _stream())
    return out


def unip_update(x, m0, m1, order: int, c_x: float, c_m0: float, rk: float, rk_is_reciprocal: bool, c_res: float):
    n = _flat_ok(x, m0) if m1 is None else _flat_ok(x, m0, m1)
    out = torch.empty_like(x)
    _call("wf_unip_update", _p(x), _is_bf16(x), _p(m0), _is_bf16(m0), _p(m1), _is_bf16(m1) if m1 is not None else 0,
          _p(out), order, float(c_x), float(c_m0), float(rk), int(rk_is_reciprocal), float(c_res), n, _stream())
    return out


def renoise(x0, noise, one_minus_sigma: float, sigma: float):
    n = _flat_ok(x0, noise)
    assert noise.dtype == torch.float32
    out = torch.empty(x0.shape, dtype=torch.float32, device=x0.device)
    _call("wf_renoise", _p(x0), _is_bf16(x0), _p(noise), _p(out), float(one_minus_sigma), float(sigma), n, _stream())
    return out


_ws = {}


def _workspace(key: str, nbytes: int, device):
    k = (key, device)
    if k not in _ws or _ws[k].numel() < nbytes:
        _ws[k] = torch.empty(nbytes, dtype=torch.uint8, device=device)
    return _ws[k]


def dsg(g, w, omega: float, stats=None):
    assert g.dtype == w.dtype
    n = _flat_ok(g, w)
    out = torch.empty_like(g)
    ws = _workspace("dsg", load().wf_dsg_workspace_bytes(), g.device)
    _call("wf_dsg", _p(g), _p(w), _p(out), _is_bf16(g), float(omega), n, _p(ws), _p(stats), _stream())
    return out


def flf_blend(decoded, ref, mask):
    """decoded/ref [1,3,F,H,W] fp32, mask [1,1,F,H,W] fp32."""
    assert decoded.dtype == ref.dtype == mask.dtype == torch.float32
    assert decoded.is_contiguous() and ref.is_contiguous() and mask.is_contiguous()
    ch = decoded.shape[1]
    plane = mask.numel()
    assert decoded.numel() == ch * plane == ref.numel()
    out = torch.empty_like(decoded)
    _call("wf_flf_blend", _p(decoded), _p(ref), _p(mask), _p(out), ch, plane, _stream())
    return out


def latent_denorm(x0, mean_host, inv_std_host):
    """x0 [1,C,f,h,w]; mean/inv_std: host float32 arrays already rounded to x0's dtype."""
    assert x0.is_contiguous()
    ch = x0.shape[1]
    out = torch.empty(x0.shape, dtype=torch.float32, device=x0.device)
    m = (C.c_float * ch)(*mean_host); s = (C.c_float * ch)(*inv_std_host)
    _call("wf_latent_denorm", _p(x0), _is_bf16(x0), _p(out), C.cast(m, _vp), C.cast(s, _vp), ch, x0.numel() // ch, _stream())
    return out


def latent_norm_replace(enc, x0, mean_host, inv_std_host, channels):
    assert enc.dtype == torch.float32 and enc.is_contiguous() and x0.is_contiguous() and enc.shape == x0.shape
    ch = x0.shape[1]
    mask = 0
    for c in channels:
        if 0 <= c < ch:
            mask |= 1 << c
    out = torch.empty_like(x0)
    m = (C.c_float * ch)(*mean_host); s = (C.c_float * ch)(*inv_std_host)
    _call("wf_latent_norm_replace", _p(enc), _p(x0), _is_bf16(x0), _p(out), C.cast(m, _vp), C.cast(s, _vp), mask, ch,
          x0.numel() // ch, _stream())
    return out


def quantise_u8(x, mode: int = 0, out=None):
    assert x.is_contiguous()
    out = torch.empty(x.shape, dtype=torch.uint8, device=x.device) if out is None else out
    ws = _workspace("quant", load().wf_quantise_workspace_bytes(), x.device)
    _call("wf_quantise_u8", _p(x), _is_bf16(x), _p(out), x.numel(), mode, _p(ws), _stream())
    return out


def farneback_u8(clips_u8, winsize: int = 15, iterations: int = 3):
    """uint8 [clips, T, H, W] -> fp32 flows [clips, T-1, H, W, 2] (dx, dy) between consecutive frames (single-level Farneback)."""
    assert clips_u8.dtype == torch.uint8 and clips_u8.is_contiguous() and clips_u8.dim() == 4
    n, T, H, W = clips_u8.shape
    flow = torch.empty(n, T - 1, H, W, 2, dtype=torch.float32, device=clips_u8.device)
    ws = _workspace("farneback", load().wf_farneback_workspace_bytes(n, T, H, W), clips_u8.device)
    _call("wf_farneback_u8", _p(clips_u8), n, T, H, W, winsize, iterations, _p(flow), _p(ws), _stream())
    return flow


def flow_metrics(flow_ref, flow_cand, outlier_or: bool = False):
    """[C, ..., 2] fp32 flow fields -> fp32 [C, 3]: mean EPE, mean outlier fraction (Wan: AND of the two tests, LongCat: OR),
    mean angular error (degrees) per channel."""
    assert flow_ref.shape == flow_cand.shape and flow_ref.dtype == flow_cand.dtype == torch.float32
    assert flow_ref.is_contiguous() and flow_cand.is_contiguous()
    C_ = flow_ref.shape[0]
    out = torch.empty(C_, 3, dtype=torch.float32, device=flow_ref.device)
    _call("wf_flow_metrics", _p(flow_ref), _p(flow_cand), _p(out), C_, flow_ref.numel() // (2 * C_), int(outlier_or), _stream())
    return out


def cfg_zero(cond, uncond, scale: float, stats=None):
    assert cond.dtype == uncond.dtype == torch.float32
    n = _flat_ok(cond, uncond)
    out = torch.empty_like(cond)
    ws = _workspace("dsg", load().wf_dsg_workspace_bytes(), cond.device)
    _call("wf_cfg_zero", _p(cond), _p(uncond), _p(out), float(scale), n, _p(ws), _p(stats), _stream())
    return out


def refine_upsample(video_u8, F2: int, H: int, W: int, pad_front: int = 0, pad_back: int = 0):
    """uint8 [F,H0,W0,3] on the device -> fp32 [3, pad_front+F2+pad_back, H, W] in [-1,1] (bf16-representable values)."""
    assert video_u8.dtype == torch.uint8 and video_u8.is_contiguous() and video_u8.dim() == 4 and video_u8.shape[3] == 3
    F, H0, W0, _ = video_u8.shape
    out = torch.empty(3, pad_front + F2 + pad_back, H, W, dtype=torch.float32, device=video_u8.device)
    _call("wf_refine_upsample", _p(video_u8), F, H0, W0, _p(out), F2, H, W, pad_front, pad_back, _stream())
    return out


# ------------------------------------------------------------------------------ VAE kernels

def conv_tf32(inp, weights, bias, taps, out, *, T, H, W, Cout, t_stride=1, t_off=0, t_mul=1, c_split=None, sy=1, sx=1,
              oy=0, ox=0, resid=None, planar_clamp=False, tile_w=16, out_hw=None, ldc=None, round_out=False,
              norm_gamma=None, norm_out=None, norm_silu=True):
    """inp: channels-last fp32 [in_T, in_H, in_W, Cin]; weights fp32 [ntaps*Cout, Cin]; taps: list of (dt,dy,dx).
    out: channels-last fp32 [frames, out_H, out_W, ldc] (or planar [c, frames, out_H, out_W] with planar_clamp).
    The tensor core truncates its fp32 operands to tf32: pass operands already ROUNDED to tf32 (round_tf32 / round_out of
    the producers) to get cuDNN's round-to-nearest arithmetic.  round_out: store ``out`` rounded to tf32.
    norm_gamma: fuse the RMS-norm (+SiLU) that follows this convolution into its epilogue (Cout <= 192): the normalised,
    tf32-rounded activation goes to ``norm_out``, or replaces the raw result in ``out`` when ``norm_out`` is None."""
    assert inp.dtype == torch.float32 and inp.is_contiguous() and inp.dim() == 4
    in_T, in_H, in_W, Cin = inp.shape
    ntaps = len(taps)
    assert weights.dtype == torch.float32 and weights.is_contiguous() and weights.shape == (ntaps * Cout, Cin)
    assert out.dtype == torch.float32 and out.is_contiguous()
    tb = (C.c_byte * (3 * ntaps))(*[v for t3 in taps for v in t3])
    if planar_clamp:
        cs = out.shape[0]
        oH, oW = out.shape[2], out.shape[3]
        cstride = out.stride(0)
        ld = 0
    else:
        oH, oW = (out.shape[1], out.shape[2]) if out_hw is None else out_hw
        ld = out.shape[3] if ldc is None else ldc
        cs = Cout if c_split is None else c_split
        cstride = 0
    _call("wf_conv_tf32", _p(inp), in_T, in_H, in_W, Cin, _p(weights), _p(bias), Cout, ntaps, C.cast(tb, _vp), T, H, W,
          t_stride, t_off, _p(out), ld, oH, oW, t_mul, cs, sy, sx, oy, ox, _p(resid), int(planar_clamp), cstride, tile_w,
          int(round_out), _p(norm_gamma), _p(norm_out), int(norm_silu), _stream())
    return out


def rms_norm_cl(x, gamma, out=None, silu=True, round_tf32=True):
    """RMS-norm (+SiLU) over the channels of every pixel.  Its result always feeds a convolution in the VAE, so by default
    it is stored rounded to tf32 (cuDNN's operand conversion; see conv_tf32)."""
    C_ = x.shape[-1]
    assert x.dtype == torch.float32 and x.is_contiguous() and gamma.numel() == C_
    out = torch.empty_like(x) if out is None else out
    _call("wf_rms_norm_cl", _p(x), C_, _p(out), C_, _p(gamma), x.numel() // C_, C_, int(silu), int(round_tf32), _stream())
    return out


def planar_to_cl(src, Cp, round_tf32=False):
    """[C, ...] planar fp32 -> [..., Cp] channels-last with zero-padded channels (optionally rounded to tf32)."""
    assert src.dtype == torch.float32 and src.is_contiguous()
    C_ = src.shape[0]
    n = src.numel() // C_
    dst = torch.empty(*src.shape[1:], Cp, dtype=torch.float32, device=src.device)
    _call("wf_planar_to_cl", _p(src), _p(dst), n, C_, Cp, int(round_tf32), _stream())
    return dst


def round_tf32(src, dst=None):
    """src rounded to the nearest tf32 (what cuDNN feeds the tensor cores), as a new tensor or into ``dst``."""
    assert src.dtype == torch.float32 and src.is_contiguous() and src.numel() % 4 == 0
    dst = torch.empty_like(src) if dst is None else dst
    _call("wf_round_tf32", _p(src), _p(dst), src.numel(), _stream())
    return dst


def cl_to_planar(src, C_):
    assert src.dtype == torch.float32 and src.is_contiguous()
    ld = src.shape[-1]
    n = src.numel() // ld
    dst = torch.empty(C_, *src.shape[:-1], dtype=torch.float32, device=src.device)
    _call("wf_cl_to_planar", _p(src), _p(dst), n, C_, ld, _stream())
    return dst


def space_to_depth(src):
    T, H, W, C_ = src.shape
    assert src.dtype == torch.float32 and src.is_contiguous()
    dst = torch.empty(T, H // 2, W // 2, 4 * C_, dtype=torch.float32, device=src.device)
    _call("wf_space_to_depth", _p(src), _p(dst), T, H, W, C_, _stream())
    return dst


def softmax_rows_(x, scale: float):
    assert x.dtype == torch.float32 and x.dim() == 2 and x.stride(1) == 1
    _call("wf_softmax_rows", _p(x), x.shape[0], x.shape[1], x.stride(0), float(scale), _stream())
    return x


def transpose_f32(src, dst):
    assert src.dtype == torch.float32 and src.dim() == 2 and src.stride(1) == 1 and dst.is_contiguous()
    R, C_ = src.shape
    assert dst.shape == (C_, R)
    _call("wf_transpose_f32", _p(src), _p(dst), R, C_, src.stride(0), dst.stride(0), _stream())
    return dst


def split_tf32(x, hi=None, lo=None):
    """x = hi + lo with hi exactly representable in tf32 (what the tensor core reads of x) and lo the exact remainder."""
    assert x.dtype == torch.float32 and x.is_contiguous() and x.numel() % 4 == 0
    hi = torch.empty_like(x) if hi is None else hi
    lo = torch.empty_like(x) if lo is None else lo
    _call("wf_split_tf32", _p(x), _p(hi), _p(lo), x.numel(), _stream())
    return hi, lo
