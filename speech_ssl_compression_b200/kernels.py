"""Tensor-level wrappers over the C ABI (one Python function per ``mh_*`` entry point).

PyTorch is used for device memory and streams only; every function here launches exactly
the hand-written sm_100a kernels in ``csrc/`` on ``torch.cuda.current_stream()``.
"""
import ctypes
from ctypes import byref, c_float, c_int, c_longlong, c_uint32, c_uint64, c_void_p

import torch

from . import lib as L
from .lib import EPI_ADD, EPI_BF16, EPI_DELTA, EPI_DGELU, EPI_F32, EPI_GELU, EPI_RES  # noqa: F401

bf16 = torch.bfloat16


def _p(t):
    return None if t is None else c_void_p(t.data_ptr())


def _s():
    return c_void_p(torch.cuda.current_stream().cuda_stream)


def _chk2d(t, name, dtype=None):
    if t.dim() != 2 or t.stride(1) != 1:
        raise ValueError(f"{name}: expected a 2-D tensor with unit inner stride, got {tuple(t.shape)} / {t.stride()}")
    if dtype is not None and t.dtype != dtype:
        raise ValueError(f"{name}: expected {dtype}, got {t.dtype}")


def gemm(a, b, out, *, a_mn=False, b_mn=False, epilogue=EPI_BF16, bias=None, aux_in=None, aux_out=None,
         mask=None, p_drop=0.0, seed=0, site=0, block_n=0, split_k=0, delta=None, delta_T=0):
    """out[M,N] (op)= A[M,K] @ B[N,K]^T.  ``a``: [M,K] (or [K,M] when a_mn), ``b``: [N,K] (or
    [K,N] when b_mn); bf16, unit inner stride.  ``out`` bf16 (or fp32 for EPI_F32, accumulated)."""
    _chk2d(a, "a", bf16)
    _chk2d(b, "b", bf16)
    _chk2d(out, "out")
    M, K = (a.shape[1], a.shape[0]) if a_mn else (a.shape[0], a.shape[1])
    N, Kb = (b.shape[1], b.shape[0]) if b_mn else (b.shape[0], b.shape[1])
    if K != Kb or out.shape[0] != M or out.shape[1] != N:
        raise ValueError(f"gemm shape mismatch: A {tuple(a.shape)} B {tuple(b.shape)} out {tuple(out.shape)}")
    want = torch.float32 if epilogue == EPI_F32 else bf16
    if out.dtype != want:
        raise ValueError(f"gemm: out dtype {out.dtype}, expected {want}")
    args = L.GemmArgs()
    args.M, args.N, args.K = M, N, K
    args.A, args.lda, args.a_mn = a.data_ptr(), a.stride(0), int(a_mn)
    args.B, args.ldb, args.b_mn = b.data_ptr(), b.stride(0), int(b_mn)
    args.D, args.ldd = out.data_ptr(), out.stride(0)
    args.epilogue = epilogue
    if bias is not None:
        if bias.dtype != torch.float32 or bias.numel() != N:
            raise ValueError("gemm: bias must be fp32 [N]")
        args.bias = bias.data_ptr()
    ld_aux = 0
    for t, nm in ((aux_in, "aux_in"), (aux_out, "aux_out")):
        if t is not None:
            _chk2d(t, nm, bf16)
            if tuple(t.shape) != (M, N):
                raise ValueError(f"gemm: {nm} shape {tuple(t.shape)} != {(M, N)}")
            if ld_aux and ld_aux != t.stride(0):
                raise ValueError("gemm: aux_in / aux_out leading dims differ")
            ld_aux = t.stride(0)
    args.aux_in = None if aux_in is None else aux_in.data_ptr()
    args.aux_out = None if aux_out is None else aux_out.data_ptr()
    args.ld_aux = ld_aux
    if mask is not None:
        if mask.dtype not in (torch.uint8, torch.bool) or tuple(mask.shape) != (M, N) or mask.stride(0) != out.stride(0):
            raise ValueError("gemm: mask must be u8/bool [M,N] with the same leading dim as out")
        args.mask = mask.data_ptr()
    args.p_drop, args.seed, args.site = float(p_drop), int(seed), int(site)
    args.block_n, args.split_k = block_n, split_k
    if epilogue == EPI_DELTA:
        if delta is None or delta.dtype != torch.float32 or not delta.is_contiguous() or delta_T <= 0 or M % delta_T \
                or N % 64 or delta.numel() != M * (N // 64):
            raise ValueError("gemm: EPI_DELTA needs delta fp32 [M / T, N / 64, T] and delta_T dividing M")
        args.delta, args.delta_T = delta.data_ptr(), int(delta_T)
    L.check(L.lib().mh_gemm(byref(args), _s()), "mh_gemm")
    return out


def _f(x):
    return c_float(float(x))


def _call(name, *args):
    L.check(getattr(L.lib(), name)(*args), name)


# ------------------------------------------------------------------------------ attention
def attn_fwd(qkv, kv_len, B, T, heads, *, causal=False, p_drop=0.0, seed=0, site=0, want_keep=True):
    """qkv: bf16 [B*T, 3*64*heads]; kv_len: int32 [B] (or None).
    Returns (out [B*T, E], lse [B,H,T], keep bits (opaque u8 [B,H,T,16*ceil(T/128)] buffer, see mh_b200.h) or None when p_drop == 0)."""
    E = 64 * heads
    if qkv.dtype != bf16 or tuple(qkv.shape) != (B * T, 3 * E) or not qkv.is_contiguous():
        raise ValueError(f"attn_fwd: qkv must be contiguous bf16 [{B * T}, {3 * E}], got {tuple(qkv.shape)}")
    out = torch.empty(B * T, E, device=qkv.device, dtype=bf16)
    lse = torch.empty(B, heads, T, device=qkv.device, dtype=torch.float32)
    keep = None
    if p_drop > 0.0 and want_keep:
        keep = torch.empty(B, heads, T, 16 * ((T + 127) // 128), device=qkv.device, dtype=torch.uint8)
    _call("mh_attn_fwd", _p(qkv), _p(kv_len), _p(out), _p(lse), _p(keep), c_int(B), c_int(T), c_int(heads),
          c_int(int(causal)), _f(p_drop), c_uint64(seed), c_uint32(site), _s())
    return out, lse, keep


def attn_bwd(qkv, kv_len, out, dout, lse, keep, B, T, heads, *, causal=False, p_drop=0.0, seed=0, site=0, dq_acc=None,
             delta=None, bias_grad=None):
    """``dq_acc``: optional fp32 [B*T, E] workspace that the caller has ALREADY zeroed (ops zeroes it on a side stream
    under the FFN backward); without it the call zeroes a fresh one itself.  ``delta``: optional precomputed
    rowsum(dO * O) [B, heads, T] (the out_proj dgrad GEMM's EPI_DELTA epilogue); without it a kernel computes it.
    ``bias_grad``: optional fp32 [3E] that receives (+=) the column sums of the returned dqkv."""
    E = 64 * heads
    if dout.dtype != bf16 or not dout.is_contiguous() or tuple(dout.shape) != (B * T, E):
        raise ValueError("attn_bwd: dout must be contiguous bf16 [B*T, E]")
    dqkv = torch.empty_like(qkv)
    flags = 0
    if delta is None:
        delta = torch.empty(B, heads, T, device=qkv.device, dtype=torch.float32)
    else:
        if delta.dtype != torch.float32 or delta.numel() != B * heads * T or not delta.is_contiguous():
            raise ValueError("attn_bwd: delta must be contiguous fp32 [B, heads, T]")
        flags |= 2
    if dq_acc is None:
        dq_acc = torch.empty(B * T, E, device=qkv.device, dtype=torch.float32)
    else:
        if dq_acc.dtype != torch.float32 or dq_acc.numel() != B * T * E or not dq_acc.is_contiguous():
            raise ValueError("attn_bwd: dq_acc must be a contiguous fp32 workspace of B*T*E elements")
        flags |= 1
    if bias_grad is not None:
        # ``bias_grad``: fp32 [3E], the stacked q / k / v bias gradients: the k / v parts come out of the attention backward
        # itself (column sums of its fp32 dK / dV accumulators), the q part out of the pass that converts dQ to bf16
        if bias_grad.dtype != torch.float32 or bias_grad.numel() != 3 * E or not bias_grad.is_contiguous():
            raise ValueError("attn_bwd: bias_grad must be contiguous fp32 [3E]")
        _call("mh_attn_bwd_bias", _p(qkv), _p(kv_len), _p(out), _p(dout), _p(lse), _p(keep), _p(delta), _p(dq_acc), _p(dqkv),
              _p(bias_grad), c_int(B), c_int(T), c_int(heads), c_int(int(causal)), _f(p_drop), c_uint64(seed), c_uint32(site),
              c_int(flags), _s())
        return dqkv
    _call("mh_attn_bwd_ex", _p(qkv), _p(kv_len), _p(out), _p(dout), _p(lse), _p(keep), _p(delta), _p(dq_acc), _p(dqkv),
          c_int(B), c_int(T), c_int(heads), c_int(int(causal)), _f(p_drop), c_uint64(seed), c_uint32(site), c_int(flags), _s())
    return dqkv


def dq_finish_colsum(dq_acc, dqkv, colsum):
    """dqkv[:, 0:E] = bf16(dq_acc); colsum[0:3E] += column sums of dqkv (finishes a mh_attn_bwd_ex call made with flags & 4)."""
    rows, E = dq_acc.shape
    _call("mh_dq_finish_colsum", _p(dq_acc), _p(dqkv), _p(colsum), c_int(rows), c_int(E), _s())


# ------------------------------------------------------------------------------ layer norm
def layernorm_fwd(x, gamma, beta, eps=1e-5, *, p_drop=0.0, seed=0, site=0):
    rows, cols = x.shape
    y = torch.empty_like(x)
    mean = torch.empty(rows, device=x.device, dtype=torch.float32)
    rstd = torch.empty(rows, device=x.device, dtype=torch.float32)
    _call("mh_layernorm_fwd", _p(x), _p(gamma), _p(beta), _p(y), _p(mean), _p(rstd), c_int(rows), c_int(cols), _f(eps),
          _f(p_drop), c_uint64(seed), c_uint32(site), _s())
    return y, mean, rstd


def layernorm_bwd(dy, x, gamma, mean, rstd, dgamma, dbeta, *, want_drop=False, p_in=0.0, seed_in=0, site_in=0,
                  p_out=0.0, seed_out=0, site_out=0, colsum_out=None):
    """Returns (dx, dx_drop or None).  dgamma / dbeta (fp32 [cols]) are accumulated in place; so is ``colsum_out``
    (optional fp32 [cols]): the column sums of dx_drop (of dx when there is no dx_drop)."""
    rows, cols = x.shape
    dx = torch.empty_like(x)
    dx_drop = torch.empty_like(x) if want_drop else None
    if colsum_out is None:
        _call("mh_layernorm_bwd", _p(dy), _p(x), _p(gamma), _p(mean), _p(rstd), _p(dx), _p(dx_drop), _p(dgamma), _p(dbeta),
              c_int(rows), c_int(cols), _f(p_in), c_uint64(seed_in), c_uint32(site_in), _f(p_out), c_uint64(seed_out),
              c_uint32(site_out), _s())
    else:
        if colsum_out.dtype != torch.float32 or colsum_out.numel() != cols or not colsum_out.is_contiguous():
            raise ValueError("layernorm_bwd: colsum_out must be contiguous fp32 [cols]")
        _call("mh_layernorm_bwd_colsum", _p(dy), _p(x), _p(gamma), _p(mean), _p(rstd), _p(dx), _p(dx_drop), _p(dgamma),
              _p(dbeta), _p(colsum_out), c_int(rows), c_int(cols), _f(p_in), c_uint64(seed_in), c_uint32(site_in), _f(p_out),
              c_uint64(seed_out), c_uint32(site_out), _s())
    return dx, dx_drop


def dropout_apply(x, p_drop, seed, site):
    y = torch.empty_like(x)
    _call("mh_dropout_apply", _p(x), _p(y), c_int(x.shape[0]), c_int(x.shape[1]), _f(p_drop), c_uint64(seed),
          c_uint32(site), _s())
    return y


def colsum_add(x, out):
    """out[n] += sum_m x[m, n]  (x bf16 [rows, cols], out fp32 [cols])"""
    _call("mh_colsum", _p(x), c_longlong(x.stride(0)), _p(out), c_int(x.shape[0]), c_int(x.shape[1]), _s())


# ------------------------------------------------------------------------------ positional conv
def posconv_weight_prep(v, g, want_bwd=True):
    """v f32 [C, 48, 128], g f32 [..., 128] -> (w_fwd, w_bwd or None, norm[128]) operand tensors."""
    C, cg, taps = v.shape
    dev = v.device
    w_fwd = torch.empty(C * cg * taps, device=dev, dtype=bf16)
    w_bwd = torch.empty(C * cg * taps, device=dev, dtype=bf16) if want_bwd else None
    ws = torch.empty(taps, device=dev, dtype=torch.float32)
    norm = torch.empty(taps, device=dev, dtype=torch.float32)
    _call("mh_posconv_weight_prep", _p(v), _p(g), _p(w_fwd), _p(w_bwd), _p(ws), _p(norm), c_int(C), c_int(C // cg),
          c_int(taps), _s())
    return w_fwd, w_bwd, norm


def posconv_fwd(x, w_fwd, bias, B, T, groups=16, taps=128, want_z=True):
    """x bf16 [B*T, C] -> (y = x + gelu(conv(x) + bias), z = conv(x) + bias or None)"""
    C = x.shape[1]
    y = torch.empty_like(x)
    z = torch.empty_like(x) if want_z else None
    _call("mh_posconv_fwd", _p(x), _p(w_fwd), _p(bias), _p(z), _p(y), c_int(B), c_int(T), c_int(C), c_int(groups),
          c_int(taps), _s())
    return y, z


def posconv_dgrad(dz, w_bwd, dy, B, T, groups=16, taps=128):
    """dx = dy + conv^T(dz)"""
    dx = torch.empty_like(dz)
    _call("mh_posconv_dgrad", _p(dz), _p(w_bwd), _p(dy), _p(dx), c_int(B), c_int(T), c_int(dz.shape[1]), c_int(groups),
          c_int(taps), _s())
    return dx


def gelu_bwd_mul(dy, z):
    dz = torch.empty_like(dy)
    _call("mh_gelu_bwd_mul", _p(dy), _p(z), _p(dz), c_longlong(dy.numel()), _s())
    return dz


def posconv_wgrad(dz, x, dw, B, T, groups=16, taps=128):
    """dw f32 [C, 48, 128] += correlation of dz with the input window"""
    _call("mh_posconv_wgrad", _p(dz), _p(x), _p(dw), c_int(B), c_int(T), c_int(x.shape[1]), c_int(groups), c_int(taps),
          _s())


def posconv_weight_bwd(dw, v, g, norm, dv, dg):
    C, cg, taps = v.shape
    ws = torch.empty(taps, device=v.device, dtype=torch.float32)
    _call("mh_posconv_weight_bwd", _p(dw), _p(v), _p(g), _p(norm), _p(ws), _p(dv), _p(dg), c_int(C), c_int(C // cg),
          c_int(taps), _s())


# ------------------------------------------------------------------------------ prep / movement
def weight_prep(src, mask, dst, dst_t=None):
    """src fp32 [rows, cols] (contiguous), mask bool/u8 or None; dst bf16 view [rows, cols] (any ld);
    dst_t bf16 view [cols, rows] or None."""
    rows, cols = src.shape
    _call("mh_weight_prep", _p(src), _p(mask), _p(dst), c_longlong(dst.stride(0)), _p(dst_t),
          c_longlong(dst_t.stride(0) if dst_t is not None else 0), c_int(rows), c_int(cols), _s())


def bias_prep(src, mask, dst):
    _call("mh_bias_prep", _p(src), _p(mask), _p(dst), c_int(src.numel()), _s())


def to_bf16(src, dst=None):
    src = src.contiguous()
    if dst is None:
        dst = torch.empty(src.shape, device=src.device, dtype=bf16)
    _call("mh_cast_f32_to_bf16", _p(src), _p(dst), c_longlong(src.numel()), _s())
    return dst


def to_f32(src, dst=None):
    src = src.contiguous()
    if dst is None:
        dst = torch.empty(src.shape, device=src.device, dtype=torch.float32)
    _call("mh_cast_bf16_to_f32", _p(src), _p(dst), c_longlong(src.numel()), _s())
    return dst


def mask_rows_to_bf16(src, zero_row, dst=None):
    rows, cols = src.shape
    if dst is None:
        dst = torch.empty(rows, cols, device=src.device, dtype=bf16)
    _call("mh_mask_rows_f32_to_bf16", _p(src), _p(zero_row), _p(dst), c_int(rows), c_int(cols), _s())
    return dst


def zero_rows_(x, zero_row):
    _call("mh_zero_rows_bf16", _p(x), _p(zero_row), c_int(x.shape[0]), c_int(x.shape[1]), _s())
    return x


def select_rows(sel, idx=None, count=None):
    rows = sel.numel()
    if idx is None:
        idx = torch.empty(rows, device=sel.device, dtype=torch.int32)
    if count is None:
        count = torch.empty(1, device=sel.device, dtype=torch.int32)
    _call("mh_select_rows", _p(sel), _p(idx), _p(count), c_int(rows), _s())
    return idx, count


def gather_rows(src, idx, n_idx, dst=None):
    cols = src.shape[1]
    if dst is None:
        dst = torch.empty(n_idx, cols, device=src.device, dtype=bf16)
    _call("mh_gather_rows", _p(src), _p(idx), _p(dst), c_int(n_idx), c_int(cols), _s())
    return dst


def scatter_rows_add_(src, idx, n_idx, dst):
    _call("mh_scatter_rows_add", _p(src), _p(idx), _p(dst), c_int(n_idx), c_int(src.shape[1]), _s())
    return dst


def gather_labels(label, idx, n_idx):
    dst = torch.empty(n_idx, device=label.device, dtype=torch.int64)
    _call("mh_gather_labels", _p(label), _p(idx), _p(dst), c_int(n_idx), _s())
    return dst


# ------------------------------------------------------------------------------ criteria
def ce_fwd(logits, labels, acc, n_valid=None, row_loss=None):
    _call("mh_ce_fwd", _p(logits), _p(labels), _p(n_valid), _p(row_loss), _p(acc), c_int(logits.shape[0]),
          c_int(logits.shape[1]), _s())


def ce_bwd(logits, labels, grad_scale, n_valid=None):
    d = torch.empty_like(logits)
    _call("mh_ce_bwd", _p(logits), _p(labels), _p(n_valid), _p(grad_scale), _p(d), c_int(logits.shape[0]),
          c_int(logits.shape[1]), _s())
    return d


def ce_finalize(acc, weight, loss, grad_scale):
    _call("mh_ce_finalize", _p(acc), _f(weight), _p(loss), _p(grad_scale), _s())


def kd_fwd(s_logits, t_logits, labels, T, acc, n_valid=None):
    _call("mh_kd_fwd", _p(s_logits), _p(t_logits), _p(labels), _p(n_valid), _f(T), _p(acc), c_int(s_logits.shape[0]),
          c_int(s_logits.shape[1]), _s())


def kd_bwd(s_logits, t_logits, labels, T, w_hard, w_soft, n_valid=None):
    d = torch.empty_like(s_logits)
    _call("mh_kd_bwd", _p(s_logits), _p(t_logits), _p(labels), _p(n_valid), _f(T), _p(w_hard), _p(w_soft), _p(d),
          c_int(s_logits.shape[0]), c_int(s_logits.shape[1]), _s())
    return d


def kd_finalize(acc, alpha, out, w_hard, w_soft):
    _call("mh_kd_finalize", _p(acc), _f(alpha), _p(out), _p(w_hard), _p(w_soft), _s())


def l1cos_fwd(pred, target, acc):
    _call("mh_l1cos_fwd", _p(pred), _p(target), _p(acc), c_int(pred.shape[0]), c_int(pred.shape[1]), _s())


def l1cos_bwd(pred, target, w_l1, w_cos):
    d = torch.empty_like(pred)
    _call("mh_l1cos_bwd", _p(pred), _p(target), _p(w_l1), _p(w_cos), _p(d), c_int(pred.shape[0]), c_int(pred.shape[1]), _s())
    return d


# ------------------------------------------------------------------------------ pruning objects
def abs_kth_smallest(tensors, k):
    """Exact k-th smallest |w| (1-based) over the concatenation of fp32 tensors.
    Returns a device uint64[3]: (threshold bit pattern, #strictly below, #equal)."""
    dev = tensors[0].device
    n = len(tensors)
    ptrs = (c_void_p * n)(*[t.data_ptr() for t in tensors])
    sizes = (c_longlong * n)(*[t.numel() for t in tensors])
    ws = torch.zeros(2048 + 8, device=dev, dtype=torch.int64)
    res = torch.zeros(3, device=dev, dtype=torch.int64)
    _call("mh_abs_kth_smallest", ptrs, sizes, c_int(n), c_longlong(k), _p(ws), _p(res), _s())
    return res


def apply_threshold_masks(tensors, masks, res, k):
    """Clear mask bytes of the k smallest |w| (ties at the threshold resolved by flat position)."""
    tie = torch.zeros(1, device=tensors[0].device, dtype=torch.int64)
    for t, m in zip(tensors, masks):
        _call("mh_apply_threshold_mask", _p(t), _p(m), c_longlong(t.numel()), _p(res), c_longlong(k), _p(tie), _s())


def row_abs_sums(w):
    out = torch.empty(w.shape[0], device=w.device, dtype=torch.float64)
    _call("mh_row_abs_sums", _p(w), c_longlong(w.stride(0)), _p(out), c_int(w.shape[0]), c_int(w.shape[1]), _s())
    return out


def col_abs_sums(w):
    out = torch.empty(w.shape[1], device=w.device, dtype=torch.float64)
    _call("mh_col_abs_sums", _p(w), c_longlong(w.stride(0)), _p(out), c_int(w.shape[0]), c_int(w.shape[1]), _s())
    return out


# ------------------------------------------------------------------------------ optimizer
def sumsq_add(x, out, mask=None):
    if mask is None:
        _call("mh_sumsq", _p(x), c_longlong(x.numel()), _p(out), _s())
    else:
        _call("mh_sumsq_masked", _p(x), _p(mask), c_longlong(x.numel()), _p(out), _s())


def adam_step(param, grad, exp_avg, exp_avg_sq, step, *, lr, beta1, beta2, eps, weight_decay=0.0, grad_scale=1.0,
              max_norm=0.0, sumsq=None, zero_grad=True, shadow=None, mask=None, effective=None):
    if mask is None:
        _call("mh_adam_step", _p(param), _p(grad), _p(exp_avg), _p(exp_avg_sq), c_longlong(param.numel()), _f(lr), _f(beta1),
              _f(beta2), _f(eps), _f(weight_decay), _p(step), _f(grad_scale), _f(max_norm), _p(sumsq), c_int(int(zero_grad)),
              _p(shadow), _s())
    else:
        _call("mh_adam_step_masked", _p(param), _p(grad), _p(exp_avg), _p(exp_avg_sq), c_longlong(param.numel()), _f(lr),
              _f(beta1), _f(beta2), _f(eps), _f(weight_decay), _p(step), _f(grad_scale), _f(max_norm), _p(sumsq),
              c_int(int(zero_grad)), _p(shadow), _p(mask), _p(effective), _s())


def flat_effective(param, mask, shadow, effective):
    """shadow (bf16) / effective (fp32) <- param * mask over a whole flat buffer (mask None: plain copies)."""
    _call("mh_flat_effective", _p(param), _p(mask), _p(shadow), _p(effective), c_longlong(param.numel()), _s())


# ------------------------------------------------------------------------------ peer-memory gradient exchange
def _ptr_array(ptrs):
    import ctypes

    return (ctypes.c_void_p * len(ptrs))(*[ctypes.c_void_p(int(x)) for x in ptrs])


def peer_reduce_scatter(grad_ptrs, flag_ptrs, state, start, count, rank, world, ctas=0):
    _call("mh_peer_reduce_scatter", _ptr_array(grad_ptrs), _ptr_array(flag_ptrs), _p(state), c_longlong(start),
          c_longlong(count), c_int(rank), c_int(world), c_int(ctas), _s())


def peer_all_gather(grad_ptrs, flag_ptrs, state, start, count, rank, world, ctas=0):
    _call("mh_peer_all_gather", _ptr_array(grad_ptrs), _ptr_array(flag_ptrs), _p(state), c_longlong(start),
          c_longlong(count), c_int(rank), c_int(world), c_int(ctas), _s())


def peer_reduce_scatter_ce(grad_ptrs, flag_ptrs, state, staging, start, count, rank, world):
    _call("mh_peer_reduce_scatter_ce", _ptr_array(grad_ptrs), _ptr_array(flag_ptrs), _p(state), _p(staging),
          c_longlong(staging.numel()), c_longlong(start), c_longlong(count), c_int(rank), c_int(world), _s())


def peer_all_gather_ce(grad_ptrs, flag_ptrs, state, start, count, rank, world):
    _call("mh_peer_all_gather_ce", _ptr_array(grad_ptrs), _ptr_array(flag_ptrs), _p(state), c_longlong(start),
          c_longlong(count), c_int(rank), c_int(world), _s())


def peer_barrier_sum(grad_ptrs, flag_ptrs, mail_ptrs, state, vals, rank, world):
    n = 0 if vals is None else vals.numel()
    _call("mh_peer_barrier_sum", _ptr_array(grad_ptrs), _ptr_array(flag_ptrs), _ptr_array(mail_ptrs), _p(state), _p(vals),
          c_int(n), c_int(rank), c_int(world), _s())


def set_dropout_offset(counter):
    """Register a device uint64 counter that is mixed into every dropout seed (None disables)."""
    _call("mh_set_dropout_offset_ptr", _p(counter))


def counter_add(counter, v=1):
    _call("mh_counter_add", _p(counter), c_uint64(v), _s())


# ------------------------------------------------------------------------------ front-end
def fbank(wave, n_samples, mel_weights, *, mean=None, inv_std=None, frame_len=400, frame_shift=160, scale=32768.0,
          preemph=0.97, window_type=0):
    """wave: f32 [B, L] zero-padded waveforms; n_samples: int32 [B]; mel_weights: f32 [n_mel, 257].
    Returns f32 [B, max_frames, n_mel] log-mel features (normalised when mean / inv_std are given), frames past
    an utterance's end are 0.  ``max_frames`` follows from L (snip_edges)."""
    if wave.dtype != torch.float32 or wave.dim() != 2 or wave.stride(1) != 1:
        raise ValueError("fbank: wave must be f32 [B, L] with unit inner stride")
    if n_samples.dtype != torch.int32 or n_samples.numel() != wave.shape[0]:
        raise ValueError("fbank: n_samples must be int32 [B]")
    if mel_weights.dtype != torch.float32 or mel_weights.dim() != 2 or mel_weights.shape[1] != 257 or not mel_weights.is_contiguous():
        raise ValueError("fbank: mel_weights must be contiguous f32 [n_mel, 257]")
    B, L_ = wave.shape
    if L_ < frame_len:
        raise ValueError(f"fbank: waveforms shorter than one frame ({L_} < {frame_len})")
    max_frames = 1 + (L_ - frame_len) // frame_shift
    n_mel = mel_weights.shape[0]
    out = torch.empty(B, max_frames, n_mel, device=wave.device, dtype=torch.float32)
    _call("mh_fbank", _p(wave), c_longlong(wave.stride(0)), _p(n_samples), c_int(B), _p(mel_weights), _p(mean), _p(inv_std),
          _p(out), c_int(max_frames), c_int(n_mel), c_int(frame_len), c_int(frame_shift), _f(scale), _f(preemph),
          c_int(window_type), _s())
    return out
