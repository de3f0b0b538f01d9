"""Thin Python wrappers over the C ABI (include/hfl.h).  Tensors in, tensors
out; no math happens here.  Every function raises HflError on failure."""
from __future__ import annotations

from typing import Optional

import torch

from . import native as N

_p = N.ptr


def _s():
    return N.stream()


def gather_gemm(A: torch.Tensor, W: torch.Tensor, *, idx: Optional[torch.Tensor] = None,
                M: Optional[int] = None, KD: int = 1, bias=None, res=None, act: int = 0,
                out_v_f32=None, out_v_bf16=None, ln=None, relu: bool = False,
                y_mapped: bool = False, out_y_f32=None, out_y_bf16=None, out_rows=None):
    """See hfl_gather_gemm.  A: [rows, Cin] bf16; W: [N, KD*Cin] bf16; idx: [M, KD] int32."""
    assert A.dtype == torch.bfloat16 and W.dtype == torch.bfloat16
    Cin = A.shape[1]
    Nn = W.shape[0]
    assert W.shape[1] == KD * Cin, (W.shape, KD, Cin)
    if M is None:
        M = idx.shape[0] if idx is not None else A.shape[0]
    if idx is not None:
        assert idx.dtype == torch.int32 and idx.shape[1] == KD
    g, b = (ln if ln is not None else (None, None))
    N.check(N.lib().hfl_gather_gemm(
        _p(A), _p(idx), _p(W), M, Nn, KD, Cin, _p(bias), _p(res), act, _p(out_v_f32),
        _p(out_v_bf16), _p(g), _p(b), int(relu), int(y_mapped), _p(out_y_f32), _p(out_y_bf16),
        _p(out_rows), _s()))


def mlp_fused(A, W1, b1, W2, b2, *, res, out_f32, out_bf16=None, out_rows=None, M=None):
    """x[orow] = res[orow] + fc2(GELU(fc1(A))) (hfl_mlp_fused)."""
    C = A.shape[1]
    assert W1.shape == (4 * C, C) and W2.shape == (C, 4 * C)
    N.check(N.lib().hfl_mlp_fused(_p(A), _p(W1), _p(b1), _p(W2), _p(b2), A.shape[0] if M is None else M,
                                  C, _p(res), _p(out_f32), _p(out_bf16), _p(out_rows), _s()))


def proj_mlp_fused(O, Wp, bp, ln_g, ln_b, W1, b1, W2, b2, *, res, out_f32, out_bf16=None, out_rows=None,
                   M=None):
    """s = res[orow] + O Wp^T + bp ; x[orow] = s + fc2(GELU(fc1(LN(s)))) (hfl_proj_mlp_fused)."""
    C = O.shape[1]
    assert Wp.shape == (C, C) and W1.shape == (4 * C, C) and W2.shape == (C, 4 * C)
    N.check(N.lib().hfl_proj_mlp_fused(_p(O), _p(Wp), _p(bp), _p(ln_g), _p(ln_b), _p(W1), _p(b1), _p(W2),
                                       _p(b2), O.shape[0] if M is None else M, C, _p(res), _p(out_f32),
                                       _p(out_bf16), _p(out_rows), _s()))


def window_attn(qkv, out, xyzb, rpe, n_win, H, C, K, dil, hat, bnd, scale):
    N.check(N.lib().hfl_window_attn(_p(qkv), _p(out), _p(xyzb), _p(rpe), n_win, H, C, K, dil,
                                    int(hat), bnd, float(scale), _s()))


def regroup_qkv(W: torch.Tensor, b: torch.Tensor):
    """nn.Linear(C, 3C) weight / bias -> the per-4-head [q | k | v] row order hfl_qkv_attn reads."""
    C = W.shape[1]
    idx = torch.cat([torch.arange(64) + part * C + g * 64 for g in range(C // 64) for part in range(3)])
    idx = idx.to(W.device)
    return W[idx].contiguous(), b[idx].contiguous()


def qkv_attn_supported(H, C, K, dil, hat, bnd) -> bool:
    return bool(N.lib().hfl_qkv_attn_supported(H, C, K, dil, int(hat), bnd))


def qkv_attn_codes(xyzb, n_win, K, dil, hat, bnd, use_rpe=True):
    """Pair codes of a level for qkv_attn (block-invariant; hfl_qkv_attn_codes)."""
    nb = int(N.lib().hfl_qkv_attn_codes_bytes(n_win, K, int(hat)))
    codes = torch.empty(max(nb // 4, 1), dtype=torch.int32, device=xyzb.device)
    N.check(N.lib().hfl_qkv_attn_codes(_p(xyzb), n_win, K, dil, int(hat), bnd, int(use_rpe), _p(codes), _s()))
    return codes


def qkv_attn(y, Wg, bias_g, out, xyzb, rpe, n_win, H, C, K, dil, hat, bnd, scale, codes=None):
    """out = window attention of (y Wqkv^T + b), qkv never materialised (hfl_qkv_attn)."""
    N.check(N.lib().hfl_qkv_attn(_p(y), _p(Wg), _p(bias_g), _p(out), _p(xyzb), _p(rpe), n_win, y.shape[0],
                                 H, C, K, dil, int(hat), bnd, float(scale), _p(codes), _s()))


def varlen_attn(qkv, out, cu, ids, B, max_len, H, C, scale):
    N.check(N.lib().hfl_varlen_attn(_p(qkv), _p(out), _p(cu), _p(ids), B, max_len, H, C,
                                    float(scale), _s()))


def stem_conv(leaf_pts, ne, n, depth, w, g, b, out):
    N.check(N.lib().hfl_stem_conv(_p(leaf_pts), _p(ne), n, depth, _p(w), _p(g), _p(b), _p(out), _s()))


def cpe_ln(x, xb, ne, w, g_cpe, b_cpe, g1, b1, y1, cpe_out, n, rows, C, K):
    N.check(N.lib().hfl_cpe_ln(_p(x), _p(xb), _p(ne), _p(w), _p(g_cpe), _p(b_cpe), _p(g1), _p(b1),
                               _p(y1), _p(cpe_out), n, rows, C, K, _s()))


def ln_rows(x, rows, m, C, g, b, y):
    N.check(N.lib().hfl_ln_rows(_p(x), _p(rows), m, C, _p(g), _p(b), _p(y), _s()))


def rt_init(x, src, xyzb, n, n_win, K, C, depth, mode, w1, b1, h, stats_out=None):
    N.check(N.lib().hfl_rt_init(_p(x), _p(src), _p(xyzb), n, n_win, K, C, depth, mode, _p(w1),
                                _p(b1), _p(h), _p(stats_out), _s()))


def hat_rows(out, n, K, offset=0):
    N.check(N.lib().hfl_hat_rows(_p(out), n, K, offset, _s()))


def remap_hat(src, out, n, K):
    N.check(N.lib().hfl_remap_hat(_p(src), _p(out), n, K, _s()))


def f32_to_bf16(src, out):
    N.check(N.lib().hfl_f32_to_bf16(_p(src), _p(out), src.numel(), _s()))


def attn_pool(logits, x, xb, tok_off, stat, out, B, kq, ldl, K, C, ktot, q_off, scale):
    N.check(N.lib().hfl_attn_pool(_p(logits), _p(x), _p(xb), _p(tok_off), _p(stat), _p(out), B, kq, ldl, K,
                                  C, ktot, q_off, float(scale), _s()))


def mixer_tail(x, wc, bc, wr, br, out, B, kin, kout, C, od, normalize):
    N.check(N.lib().hfl_mixer_tail(_p(x), _p(wc), _p(bc), _p(wr), _p(br), _p(out), B, kin, kout, C,
                                   od, int(normalize), _s()))


def gem_pool(x, tok_off, B, K, C, pw, eps, out, ld_out, col_off):
    N.check(N.lib().hfl_gem_pool(_p(x), _p(tok_off), B, K, C, float(pw), float(eps), _p(out),
                                 ld_out, col_off, _s()))


def gem_head(pooled, w, bn_g, bn_b, bn_mean, bn_var, bn_eps, normalize, out):
    B, in_dim = pooled.shape
    N.check(N.lib().hfl_gem_head(_p(pooled), B, in_dim, _p(w), _p(bn_g), _p(bn_b), _p(bn_mean),
                                 _p(bn_var), float(bn_eps), w.shape[0], int(normalize), _p(out), _s()))


def prepare_clouds(pts: torch.Tensor, off: torch.Tensor, *, norm: bool, zero_mean: bool = True,
                   scale_factor=None, norm_range: float = 1.0, cyl: bool = False):
    """Device-side Normalize / range masks / CylindricalCoordinates + compaction of a batch of raw fp32 clouds
    (hfl_prepare_clouds).  Returns (points [n_in, 3] fp32 of which the first off_out[-1] rows are valid,
    off_out [B + 1] int32, total [1] int32 device tensor)."""
    assert pts.dtype == torch.float32 and off.dtype == torch.int32
    n_in, B = pts.shape[0], off.numel() - 1
    dev = pts.device
    tmp = torch.empty_like(pts)
    out = torch.empty_like(pts)
    cnt = torch.empty(B, dtype=torch.int32, device=dev)
    off_out = torch.empty(B + 1, dtype=torch.int32, device=dev)
    total = torch.empty(1, dtype=torch.int32, device=dev)
    N.check(N.lib().hfl_prepare_clouds(_p(pts), _p(off), B, n_in, int(norm), int(zero_mean),
                                       float(scale_factor) if scale_factor is not None else 0.0,
                                       float(norm_range if norm_range is not None else 1.0), int(cyl), _p(tmp),
                                       _p(cnt), _p(out), _p(off_out), _p(total), _s()))
    return out, off_out, total


def knn_topk(q: torch.Tensor, db: torch.Tensor, k: int = 25, idx_offset: int = 0):
    """Exact L2 top-k of every query row against a database shard; returns
    (squared distances [nq,k] fp32, indices [nq,k] int32) sorted by (dist, idx)."""
    assert q.dtype == torch.float32 and db.dtype == torch.float32
    nq, dim = q.shape
    od = torch.empty((nq, k), dtype=torch.float32, device=q.device)
    oi = torch.empty((nq, k), dtype=torch.int32, device=q.device)
    nb = int(N.lib().hfl_knn_workspace_bytes(nq, db.shape[0], k))
    ws = torch.empty((max(nb, 8),), dtype=torch.uint8, device=q.device)
    N.check(N.lib().hfl_knn_topk_ws(_p(q.contiguous()), nq, _p(db.contiguous()), db.shape[0], dim, k,
                                    idx_offset, _p(od), _p(oi), _p(ws), nb, _s()))
    return od, oi


def topk_merge(parts_d: torch.Tensor, parts_i: torch.Tensor):
    """[parts, nq, k] partial lists -> global [nq, k]."""
    P, nq, k = parts_d.shape
    od = torch.empty((nq, k), dtype=torch.float32, device=parts_d.device)
    oi = torch.empty((nq, k), dtype=torch.int32, device=parts_d.device)
    N.check(N.lib().hfl_topk_merge(_p(parts_d.contiguous()), _p(parts_i.contiguous()), P, nq, k,
                                   _p(od), _p(oi), _s()))
    return od, oi
