"""ctypes binding of libunivid_b200.so (include/univid_b200.h) and thin tensor-level wrappers.

PyTorch is used for device memory and streams only; every compute call goes through the C ABI.
There is NO fallback: if the library is missing or the device is not sm_100 the call raises.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libunivid_b200.so")

EXPORTS = (
    "uvb_version", "uvb_last_error", "uvb_qk_norm_rope", "uvb_head_scatter_bf16",
    "uvb_fmha_fwd_bf16", "uvb_fmha_workspace_bytes", "uvb_xattn_fwd_bf16", "uvb_debug_fmha_timeline",
    "uvb_qk_norm_rope_sp", "uvb_head_scatter_sp", "uvb_fmha_fwd_sp_bf16", "uvb_sp_buffer_alloc",
    "uvb_sp_buffer_free", "uvb_sp_ipc_export", "uvb_sp_ipc_import", "uvb_sp_ipc_close", "uvb_sp_signal",
    "uvb_sp_wait", "uvb_block_glue", "uvb_linear_bf16", "uvb_unipc_step", "uvb_set_knob", "uvb_get_knob",
    "uvb_linear_bf16_sp", "uvb_sp_signal_wait",
)
ABI_VERSION = 110
KNOBS = {"fmha_pair": 0, "fmha_split": 1, "gemm_ctas": 2, "gemm_bn": 3, "gemm_small": 4, "prologue_pair": 5,
         "fmha_poly": 6, "sp_wait_timeout_s": 7, "xattn_pair": 8}

UVB_BF16, UVB_F32 = 0, 1
_c = ctypes
_vp, _i, _i64, _f = _c.c_void_p, _c.c_int, _c.c_int64, _c.c_float
_lib = None


class UnipcCoef(_c.Structure):
    """uvb_unipc_coef (include/univid_b200.h)."""
    _fields_ = [("guide_scale", _f), ("sigma", _f), ("corrector_order", _c.c_int32), ("c_a", _f), ("c_b", _f),
                ("c_ab", _f), ("c_rk", _f), ("c_rho0", _f), ("c_rho_last", _f), ("predictor_order", _c.c_int32),
                ("p_a", _f), ("p_b", _f), ("p_ab", _f), ("p_rk", _f), ("p_rho0", _f), ("history_bf16", _c.c_int32)]


launch_count = 0   # kernels launched through this module (bench.py reports it as gpu_launches)


def lib():
    """Load the shared library (once). Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m univid_b200.build` "
            "(nvcc, sm_100a). univid_b200 has no fallback path.")
    L = _c.CDLL(LIB_PATH)
    L.uvb_version.restype = _i
    L.uvb_last_error.restype = _c.c_char_p
    L.uvb_qk_norm_rope.restype = _i
    L.uvb_qk_norm_rope.argtypes = [_vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i,
                                   _vp, _i, _f, _i, _i64, _i64, _i64, _vp]
    L.uvb_head_scatter_bf16.restype = _i
    L.uvb_head_scatter_bf16.argtypes = [_vp, _vp, _i, _i, _i, _i, _i64, _i64, _i64, _vp]
    L.uvb_fmha_fwd_bf16.restype = _i
    L.uvb_fmha_fwd_bf16.argtypes = [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _vp, _f,
                                    _vp, _i64, _vp]
    L.uvb_fmha_workspace_bytes.restype = _i64
    L.uvb_fmha_workspace_bytes.argtypes = []
    L.uvb_xattn_fwd_bf16.restype = _i
    L.uvb_xattn_fwd_bf16.argtypes = [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i,
                                     _vp, _vp, _vp, _vp, _f, _vp, _i64, _vp]
    _u32, _pp = _c.c_uint32, _c.POINTER(_c.c_void_p)
    L.uvb_qk_norm_rope_sp.restype = _i
    L.uvb_qk_norm_rope_sp.argtypes = [_vp, _vp, _i, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i,
                                      _vp, _i, _f, _i, _i64, _i64, _i64, _vp]
    L.uvb_head_scatter_sp.restype = _i
    L.uvb_head_scatter_sp.argtypes = [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i64, _i64, _i64, _vp]
    L.uvb_sp_signal_wait.restype = _i
    L.uvb_sp_signal_wait.argtypes = [_vp, _i, _u32, _vp, _vp]
    L.uvb_fmha_fwd_sp_bf16.restype = _i
    L.uvb_fmha_fwd_sp_bf16.argtypes = [_vp, _vp, _vp, _vp, _i, _i, _i, _vp, _i, _i, _i, _i, _vp, _vp, _vp, _f,
                                       _vp, _i64, _vp]
    L.uvb_sp_buffer_alloc.restype = _i
    L.uvb_sp_buffer_alloc.argtypes = [_i64, _pp]
    L.uvb_sp_buffer_free.restype = _i
    L.uvb_sp_buffer_free.argtypes = [_vp]
    L.uvb_sp_ipc_export.restype = _i
    L.uvb_sp_ipc_export.argtypes = [_vp, _vp]
    L.uvb_sp_ipc_import.restype = _i
    L.uvb_sp_ipc_import.argtypes = [_vp, _pp]
    L.uvb_sp_ipc_close.restype = _i
    L.uvb_sp_ipc_close.argtypes = [_vp]
    L.uvb_sp_signal.restype = _i
    L.uvb_sp_signal.argtypes = [_vp, _i, _u32, _vp]
    L.uvb_sp_wait.restype = _i
    L.uvb_sp_wait.argtypes = [_vp, _i, _u32, _vp]
    L.uvb_block_glue.restype = _i
    L.uvb_block_glue.argtypes = [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i64, _i64, _vp, _f, _i, _vp]
    L.uvb_linear_bf16.restype = _i
    L.uvb_linear_bf16.argtypes = [_vp, _vp, _vp, _vp, _i, _i, _i, _i64, _i64, _i64, _i, _vp]
    L.uvb_linear_bf16_sp.restype = _i
    L.uvb_linear_bf16_sp.argtypes = [_vp, _vp, _vp, _vp, _i, _i, _i, _i, _i64, _i64, _i64, _i, _vp]
    L.uvb_unipc_step.restype = _i
    L.uvb_unipc_step.argtypes = [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, _i64, _c.POINTER(UnipcCoef), _vp]
    L.uvb_set_knob.restype = _i
    L.uvb_set_knob.argtypes = [_i, _i]
    L.uvb_get_knob.restype = _i
    L.uvb_get_knob.argtypes = [_i]
    if L.uvb_version() != ABI_VERSION:
        raise RuntimeError(f"{LIB_PATH} has ABI version {L.uvb_version()}, expected {ABI_VERSION}: rebuild it "
                           "with `python -m univid_b200.build --force`")
    _lib = L
    return L


def set_knob(name, value):
    """Explicit tuning knob of the library (uvb_set_knob; names in KNOBS).  Returns the previous value."""
    old = lib().uvb_get_knob(KNOBS[name])
    _check(lib().uvb_set_knob(KNOBS[name], int(value)))
    return old


def _check(rc):
    if rc != 0:
        raise RuntimeError(f"univid_b200 error {rc}: {lib().uvb_last_error().decode()}")


def _ptr(t):
    return None if t is None else t.data_ptr()


def _stream(t):
    """The current stream of t's device.  The C ABI launches on the CURRENT device (cudaGetDevice) and caches the SM
    count / kernel attributes per current device, so the tensors must live there: refuse anything else loudly instead
    of launching on the wrong GPU with a foreign stream handle."""
    idx = t.device.index
    if idx is not None and idx != torch.cuda.current_device():
        raise RuntimeError(f"univid_b200: tensors are on cuda:{idx} but the current device is "
                           f"cuda:{torch.cuda.current_device()}; call torch.cuda.set_device / use torch.cuda.device()")
    return torch.cuda.current_stream(t.device).cuda_stream


def _require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise RuntimeError("univid_b200 kernels need CUDA tensors; there is no CPU path")


def _no_grad_only(*tensors):
    """The hot path is forward-only: refuse to run under autograd silently."""
    if torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in tensors):
        raise RuntimeError(
            "univid_b200 attention kernels are forward-only; run under torch.no_grad() "
            "(training through the DiT is outside this path's contract)")


def _strides3(t):
    """(batch, token, head) element strides of a [B, L, N, 128] tensor as a ctypes int64[3]."""
    if t.stride(3) != 1:
        raise RuntimeError("head_dim must be the contiguous dimension")
    return (_i64 * 3)(t.stride(0), t.stride(1), t.stride(2))


_GRID_CACHE = {}


def _grid_array(grid_sizes):
    key = tuple(map(tuple, grid_sizes)) if not torch.is_tensor(grid_sizes) else tuple(map(tuple, grid_sizes.tolist()))
    arr = _GRID_CACHE.get(key)
    if arr is None:
        flat = [int(v) for row in key for v in row]
        arr = (_c.c_int32 * len(flat))(*flat)
        _GRID_CACHE[key] = arr
    return arr, key


def _grouped_out(o, src, groups, B, L, hpg, device):
    """Output buffer [B, L, N, 128] (groups == 1) or [groups, B, L, N/groups, 128]; a caller-provided
    buffer may be a strided view (e.g. one slot of a fused send buffer) as long as its two innermost
    dimensions are dense."""
    if src is None:
        return None
    shape = (B, L, hpg, 128) if groups == 1 else (groups, B, L, hpg, 128)
    if o is None:
        return torch.empty(shape, dtype=torch.bfloat16, device=device)
    if tuple(o.shape) != shape or o.dtype != torch.bfloat16 or o.stride(-1) != 1 or o.stride(-2) != 128:
        raise RuntimeError(f"bad output buffer: need bf16 {shape} with dense (head, dim)")
    return o


def _grouped_strides(o, groups):
    """(out_sb, out_sl, out_sg) element strides of a buffer returned by _grouped_out."""
    if groups == 1:
        return o.stride(0), o.stride(1), 0
    return o.stride(1), o.stride(2), o.stride(0)


def ptr_array(ptrs):
    """ctypes void*[n] from a list of integer device addresses (kept alive by the caller)."""
    return (_c.c_void_p * len(ptrs))(*[int(p) for p in ptrs])


def qk_norm_rope(q_in, k_in, wq, wk, eps, num_heads, cos_sin=None, grid_sizes=None, tok_offset=0,
                 row_scale=None, pre_bias=None, groups=1, q_out=None, k_out=None, peers=None):
    """Fused RMSNorm (+RoPE) of q and/or k: [B, L, dim] -> bf16 [B, L, N, 128] (groups == 1) or the
    Ulysses send layout [groups, B, L, N/groups, 128].  See uvb_qk_norm_rope in the header.
    peers = (q_ptrs, k_ptrs, out_sb, out_sl): head group j is stored through q_ptrs[j] / k_ptrs[j]
    (ctypes void*[groups] of peer-mapped device addresses, uvb_qk_norm_rope_sp) and nothing is returned."""
    global launch_count
    ref = q_in if q_in is not None else k_in
    _require_cuda(q_in, k_in, wq, wk, cos_sin, row_scale, pre_bias)
    _no_grad_only(q_in, k_in)
    B, L, dim = ref.shape
    N = num_heads
    if dim != N * 128:
        raise NotImplementedError(f"head_dim {dim // N} is not supported (128 only)")
    if N % groups != 0:
        raise ValueError(f"num_heads {N} not divisible by {groups}")
    dt = ref.dtype
    if dt not in (torch.bfloat16, torch.float32):
        raise NotImplementedError(f"qk_norm_rope input dtype {dt} (bf16 / fp32 only)")
    for t in (q_in, k_in):
        if t is not None and (t.dtype != dt or not t.is_contiguous() or t.shape != ref.shape):
            raise RuntimeError("q_in / k_in must be contiguous, same shape and dtype")
    hpg = N // groups
    if peers is not None:
        q_ptrs, k_ptrs, out_sb, out_sl = peers
        grid = None
        if cos_sin is not None:
            grid, key = _grid_array(grid_sizes)
            if len(key) != B:
                raise ValueError("grid_sizes must have one (f, h, w) row per sample")
        f32 = lambda t: None if t is None else (t if t.dtype == torch.float32 else t.float()).contiguous()
        wq, wk, row_scale, pre_bias = f32(wq), f32(wk), f32(row_scale), f32(pre_bias)
        _check(lib().uvb_qk_norm_rope_sp(
            _ptr(q_in), _ptr(k_in), UVB_BF16 if dt == torch.bfloat16 else UVB_F32, _ptr(wq), _ptr(wk),
            _ptr(cos_sin), _ptr(row_scale), _ptr(pre_bias), None, None,
            None if q_in is None else _c.cast(q_ptrs, _vp), None if k_in is None else _c.cast(k_ptrs, _vp),
            groups, B, L, N, None if grid is None else _c.cast(grid, _vp), int(tok_offset), float(eps), hpg,
            int(out_sb), int(out_sl), 0, _stream(ref)))
        launch_count += 1
        return None, None
    q_out = _grouped_out(q_out, q_in, groups, B, L, hpg, ref.device)
    k_out = _grouped_out(k_out, k_in, groups, B, L, hpg, ref.device)
    strides = _grouped_strides(q_out if q_out is not None else k_out, groups)
    if q_out is not None and k_out is not None and _grouped_strides(k_out, groups) != strides:
        raise RuntimeError("q_out and k_out must share strides")
    grid = None
    if cos_sin is not None:
        grid, key = _grid_array(grid_sizes)
        if len(key) != B:
            raise ValueError("grid_sizes must have one (f, h, w) row per sample")
    f32 = lambda t: None if t is None else (t if t.dtype == torch.float32 else t.float()).contiguous()
    wq, wk, row_scale, pre_bias = f32(wq), f32(wk), f32(row_scale), f32(pre_bias)
    _check(lib().uvb_qk_norm_rope(
        _ptr(q_in), _ptr(k_in), UVB_BF16 if dt == torch.bfloat16 else UVB_F32, _ptr(wq), _ptr(wk),
        _ptr(cos_sin), _ptr(row_scale), _ptr(pre_bias), _ptr(q_out), _ptr(k_out), B, L, N,
        None if grid is None else _c.cast(grid, _vp), int(tok_offset), float(eps), hpg,
        strides[0], strides[1], strides[2], _stream(ref)))
    launch_count += 1
    return q_out, k_out


def head_scatter(v, groups, out=None, peers=None):
    """v [B, L, N, 128] bf16 -> [groups, B, L, N/groups, 128] (Ulysses send layout); with
    peers = (ptrs, out_sb, out_sl) head group j is stored through ptrs[j] (uvb_head_scatter_sp)."""
    global launch_count
    _require_cuda(v)
    B, L, N, D = v.shape
    if D != 128 or v.dtype != torch.bfloat16 or not v.is_contiguous():
        raise RuntimeError("head_scatter expects contiguous bf16 [B, L, N, 128]")
    hpg = N // groups
    if peers is not None:
        ptrs, out_sb, out_sl = peers
        _check(lib().uvb_head_scatter_sp(_ptr(v), None, _c.cast(ptrs, _vp), groups, B, L, N, hpg, int(out_sb),
                                         int(out_sl), 0, _stream(v)))
        launch_count += 1
        return None
    out = _grouped_out(out, v, groups, B, L, hpg, v.device)
    sb, sl, sg = _grouped_strides(out, groups)
    _check(lib().uvb_head_scatter_bf16(_ptr(v), _ptr(out), B, L, N, hpg, sb, sl, sg, _stream(v)))
    launch_count += 1
    return out


def _pad128(t, lk, fill=1.0):
    if t is None:
        return None
    t = t.float().contiguous()
    n = (lk + 127) // 128 * 128
    if t.numel() == n:
        return t
    if t.numel() != lk:
        raise ValueError("per-key vector must have Lk entries")
    out = torch.full((n,), fill, dtype=torch.float32, device=t.device)
    out[:lk] = t
    return out


_WORKSPACES = {}


def _fmha_workspace(device, stream):
    """Scratch buffer for the split of remainder query blocks over the key axis (uvb_fmha_workspace_bytes):
    one per (device, stream), zero-filled once; the kernels hand it back zero-filled."""
    key = (device.index if device.index is not None else torch.cuda.current_device(), stream)
    ws = _WORKSPACES.get(key)
    if ws is None:
        with torch.cuda.device(device):
            n = int(lib().uvb_fmha_workspace_bytes())
        if n <= 0:
            raise RuntimeError("uvb_fmha_workspace_bytes failed: " + lib().uvb_last_error().decode())
        ws = torch.zeros(n, dtype=torch.uint8, device=device)
        _WORKSPACES[key] = ws
    return ws


def fmha_fwd(q, k, v, k_lens=None, softmax_scale=None, out=None, key_logit_scale=None,
             key_pv_weight=None, out_bias=None, split_units=True):
    """softmax(q k^T * scale) v on [B, L, N, 128] bf16 tensors (any strides with contiguous head_dim).
    k_lens: int32 CUDA tensor [B] or None.  The per-key modifiers select uvb_xattn_fwd_bf16."""
    global launch_count
    _require_cuda(q, k, v, k_lens, key_logit_scale, key_pv_weight, out_bias)
    _no_grad_only(q, k, v)
    if q.dim() != 4 or k.dim() != 4 or v.dim() != 4:
        raise ValueError("q, k, v must be [B, L, N, D]")
    B, Lq, N, D = q.shape
    Lk = k.shape[1]
    if D != 128 or v.shape[3] != 128:
        raise NotImplementedError(f"head_dim {D} is not supported by the sm_100a kernel (128 only)")
    if k.shape != (B, Lk, N, D) or v.shape != (B, Lk, N, D):
        raise NotImplementedError("grouped-query shapes (Nk != Nq) are not supported")
    for t in (q, k, v):
        if t.dtype != torch.bfloat16:
            raise NotImplementedError(f"fmha_fwd computes in bf16; got {t.dtype}")
    if out is None:
        out = torch.empty((B, Lq, N, D), dtype=torch.bfloat16, device=q.device)
    if k_lens is not None and (k_lens.dtype != torch.int32 or k_lens.numel() != B):
        raise ValueError("k_lens must be int32 [B] on the device")
    scale = float(D ** -0.5 if softmax_scale is None else softmax_scale)
    stream = _stream(q)
    # split_units=False passes workspace = NULL: the schedule that never splits a query block over CTAs
    ws = _fmha_workspace(q.device, stream) if split_units else None
    args = (B, Lq, Lk, N, _c.cast(_strides3(q), _vp), _c.cast(_strides3(k), _vp),
            _c.cast(_strides3(v), _vp), _c.cast(_strides3(out), _vp), scale,
            None if ws is None else ws.data_ptr(), 0 if ws is None else ws.numel(), stream)
    if key_logit_scale is None and key_pv_weight is None and out_bias is None:
        _check(lib().uvb_fmha_fwd_bf16(_ptr(q), _ptr(k), _ptr(v), _ptr(out), _ptr(k_lens), *args))
    else:
        kls, pvw = _pad128(key_logit_scale, Lk), _pad128(key_pv_weight, Lk)
        ob = None if out_bias is None else out_bias.float().contiguous()
        if ob is not None and ob.numel() != N * D:
            raise ValueError("out_bias must have N*128 entries")
        _check(lib().uvb_xattn_fwd_bf16(_ptr(q), _ptr(k), _ptr(v), _ptr(out), _ptr(k_lens), _ptr(kls),
                                        _ptr(pvw), _ptr(ob), *args))
    launch_count += 1
    return out


def fmha_fwd_sp(q, k, v, o_ptrs, n_peers, head_offset, total_heads, k_lens=None, softmax_scale=None):
    """Attention on a head shard q/k/v [B, L, n, 128] whose output rows are TMA-stored into the ranks that own
    them: rows [j*L/p, (j+1)*L/p) -> o_ptrs[j] = [B, L/p, total_heads, 128] at heads [head_offset, +n)
    (uvb_fmha_fwd_sp_bf16)."""
    global launch_count
    _require_cuda(q, k, v, k_lens)
    _no_grad_only(q, k, v)
    B, Lq, N, D = q.shape
    Lk = k.shape[1]
    if D != 128 or k.shape != (B, Lk, N, D) or v.shape != (B, Lk, N, D):
        raise NotImplementedError("fmha_fwd_sp: q/k/v must be [B, L, n, 128] with equal head counts")
    for t in (q, k, v):
        if t.dtype != torch.bfloat16:
            raise NotImplementedError(f"fmha_fwd_sp computes in bf16; got {t.dtype}")
    if k_lens is not None and (k_lens.dtype != torch.int32 or k_lens.numel() != B):
        raise ValueError("k_lens must be int32 [B] on the device")
    scale = float(D ** -0.5 if softmax_scale is None else softmax_scale)
    stream = _stream(q)
    ws = _fmha_workspace(q.device, stream)
    _check(lib().uvb_fmha_fwd_sp_bf16(
        _ptr(q), _ptr(k), _ptr(v), _c.cast(o_ptrs, _vp), int(n_peers), int(head_offset), int(total_heads),
        _ptr(k_lens), B, Lq, Lk, N, _c.cast(_strides3(q), _vp), _c.cast(_strides3(k), _vp),
        _c.cast(_strides3(v), _vp), scale, ws.data_ptr(), ws.numel(), stream))
    launch_count += 1


def sp_signal(flag_ptrs, n, value, stream):
    global launch_count
    _check(lib().uvb_sp_signal(_c.cast(flag_ptrs, _vp), int(n), int(value) & 0xffffffff, stream))
    launch_count += 1


def sp_signal_wait(flag_ptrs, n, value, wait_ptr, stream):
    global launch_count
    _check(lib().uvb_sp_signal_wait(_c.cast(flag_ptrs, _vp), int(n), int(value) & 0xffffffff, int(wait_ptr), stream))
    launch_count += 1


def sp_wait(flags_ptr, n, value, stream):
    global launch_count
    _check(lib().uvb_sp_wait(int(flags_ptr), int(n), int(value) & 0xffffffff, stream))
    launch_count += 1


GLUE_DIMS = (256, 512, 1024, 1536, 2048, 3072, 4096, 5120)


def block_glue(x, y=None, gate=None, ln=None, scale=None, shift=None, eps=1e-6, want_h=True, inplace=False,
               index=None, ln_round_bf16=False):
    """Fused residual / LayerNorm / adaLN glue of a WanAttentionBlock (uvb_block_glue):
        x' = x + y * gate   (if y is given; a new tensor unless inplace)   h = LN(x')[* w + b] * (1 + scale) + shift -> bf16
    x fp32 [B, L, dim] contiguous; y bf16 [B, L, dim]; gate / scale / shift fp32 [B, 1 or L, dim] views with
    unit stride over dim (chunks of the modulation tensor); ln = None | (weight, bias).
    index: int32 [B, L] or None -- token (b, l) uses modulation row index[b, l]; the chunks are then [1 or B, U, dim].
    ln_round_bf16: round the LayerNorm result to bf16 before the modulation (WanLayerNorm's .type_as(x) for a bf16 x).
    Returns (x', h)."""
    global launch_count
    _require_cuda(x, y, gate, scale, shift, index)
    _no_grad_only(x, y, gate, scale, shift)
    B, L, dim = x.shape
    if x.dtype != torch.float32 or not x.is_contiguous():
        raise RuntimeError("block_glue: x must be contiguous fp32 [B, L, dim]")
    if dim not in GLUE_DIMS:
        raise NotImplementedError(f"block_glue: dim {dim} is not supported")
    if y is not None and (y.dtype != torch.bfloat16 or y.shape != x.shape or not y.is_contiguous()):
        raise RuntimeError("block_glue: y must be contiguous bf16 with x's shape")
    if index is not None and (index.dtype != torch.int32 or tuple(index.shape) != (B, L) or not index.is_contiguous()):
        raise RuntimeError("block_glue: index must be a contiguous int32 [B, L] tensor")
    sb = sl = None
    for m in (gate, scale, shift):
        if m is None:
            continue
        if m.dtype != torch.float32 or m.dim() != 3 or m.size(2) != dim or m.stride(2) != 1:
            raise RuntimeError("block_glue: modulation chunks must be fp32 [B, 1|L, dim] with unit dim stride")
        if index is None:
            if m.size(0) != B or m.size(1) not in (1, L):
                raise RuntimeError("block_glue: modulation chunks must be fp32 [B, 1|L, dim] with unit dim stride")
            msb = m.stride(0) if B > 1 else 0
            msl = m.stride(1) if m.size(1) == L and L > 1 else 0
        else:
            if m.size(0) not in (1, B):
                raise RuntimeError("block_glue: indexed modulation chunks must be fp32 [1|B, U, dim]")
            msb = m.stride(0) if m.size(0) > 1 else 0
            msl = m.stride(1)
        if sb is None:
            sb, sl = msb, msl
        elif (sb, sl) != (msb, msl):
            raise RuntimeError("block_glue: gate / scale / shift must share strides")
    lw = lb = None
    if ln is not None:
        lw, lb = ln
        lw = None if lw is None else lw.float().contiguous()
        lb = None if lb is None else lb.float().contiguous()
    h = torch.empty((B, L, dim), dtype=torch.bfloat16, device=x.device) if want_h else None
    x_new = x if (y is None or inplace) else torch.empty_like(x)
    _check(lib().uvb_block_glue(_ptr(x), _ptr(y), _ptr(gate), _ptr(x_new) if y is not None else None, _ptr(lw),
                                _ptr(lb), _ptr(scale), _ptr(shift), _ptr(h), B, L, dim, sb or 0, sl or 0,
                                _ptr(index) if sb is not None else None, float(eps), 1 if ln_round_bf16 else 0,
                                _stream(x)))
    launch_count += 1
    return x_new, h


ACT_NONE, ACT_GELU_TANH = 0, 1


def linear(x, weight, bias=None, act=ACT_NONE, out=None, peers=None):
    """y = act(x @ weight.T + bias) (uvb_linear_bf16): x bf16 [..., K] with dense rows, weight bf16 [N, K]
    (nn.Linear layout), bias fp32 [N] (already rounded to bf16 values when mirroring autocast) or None.
    Returns bf16 [..., N].
    peers = (ptrs, n, ld): column group j of the result ([M, N/n]) is stored through ptrs[j] with leading dimension
    ld instead (uvb_linear_bf16_sp: the projection's epilogue writes the Ulysses send layout into the peers'
    buffers); nothing is returned."""
    global launch_count
    _require_cuda(x, weight, bias)
    _no_grad_only(x, weight, bias)
    if x.dtype != torch.bfloat16 or weight.dtype != torch.bfloat16:
        raise NotImplementedError(f"linear computes in bf16; got x {x.dtype}, weight {weight.dtype}")
    if weight.dim() != 2 or x.size(-1) != weight.size(1):
        raise ValueError(f"linear: x [..., {x.size(-1)}] does not match weight {tuple(weight.shape)}")
    N, K = weight.shape
    if N % 8 != 0 or K % 8 != 0:
        raise NotImplementedError(f"linear: N={N} and K={K} must be multiples of 8")
    if weight.stride(1) != 1:
        weight = weight.contiguous()
    x2 = x.reshape(-1, K)
    if x2.stride(1) != 1 or x2.stride(0) % 8 != 0 or x2.data_ptr() % 16 != 0:
        x2 = x2.contiguous()
    M = x2.size(0)
    if bias is not None:
        if bias.numel() != N:
            raise ValueError("linear: bias must have N entries")
        bias = (bias if bias.dtype == torch.float32 else bias.float()).contiguous()
    if peers is not None:
        ptrs, n_peers, ld = peers
        if M > 0:
            _check(lib().uvb_linear_bf16_sp(_ptr(x2), _ptr(weight), _ptr(bias), _c.cast(ptrs, _vp), int(n_peers), M, N, K,
                                            x2.stride(0), weight.stride(0), int(ld), int(act), _stream(x)))
            launch_count += 1
        return None
    if out is None:
        out = torch.empty((M, N), dtype=torch.bfloat16, device=x.device)
    elif out.dtype != torch.bfloat16 or out.numel() != M * N or not out.is_contiguous():
        raise RuntimeError("linear: out must be a contiguous bf16 tensor of M*N elements")
    if M > 0:
        _check(lib().uvb_linear_bf16(_ptr(x2), _ptr(weight), _ptr(bias), _ptr(out), M, N, K, x2.stride(0),
                                     weight.stride(0), N, int(act), _stream(x)))
        launch_count += 1
    return out.view(*x.shape[:-1], N)


def unipc_step(cond, uncond, x, last, m0, m1, coef):
    """Fused CFG combine + UniPC step (uvb_unipc_step).  All tensors fp32 CUDA with x's number of elements; coef is
    a UnipcCoef.  Returns (m_t, x_corrected, x_next) as new tensors shaped like x."""
    global launch_count
    _require_cuda(cond, uncond, x, last, m0, m1)
    _no_grad_only(cond, uncond, x)
    n = x.numel()
    ins = []
    for t in (cond, uncond, x, last, m0, m1):
        if t is None:
            ins.append(None)
            continue
        if t.dtype != torch.float32 or t.numel() != n:
            raise RuntimeError("unipc_step: every tensor must be fp32 with the sample's number of elements")
        ins.append(t.contiguous())
    m_t, x_c, x_n = torch.empty_like(ins[2]), torch.empty_like(ins[2]), torch.empty_like(ins[2])
    _check(lib().uvb_unipc_step(*[_ptr(t) for t in ins], _ptr(m_t), _ptr(x_c), _ptr(x_n), n, _c.byref(coef), _stream(x)))
    launch_count += 1
    return m_t, x_c, x_n
