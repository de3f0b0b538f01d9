"""GPU numerics of every kernel behind include/hfl.h against a plain PyTorch fp32
statement of the same op (on the same bf16-rounded operands).  Tolerances are
written next to each check."""
import math
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _ops():
    from hotformerloc_b200 import ops
    return ops


def _bf(x):
    return x.to(torch.bfloat16)


DEV = 'cuda'


@pytest.mark.parametrize('M,N,K', [(128, 128, 128), (1000, 256, 256), (4100, 384, 128),
                                   (777, 768, 256), (3000, 1024, 256), (513, 256, 1024),
                                   (130, 64, 256), (50000, 256, 256)])
def test_gemm_dense_bias(M, N, K):
    torch.manual_seed(0)
    A = _bf(torch.randn(M, K, device=DEV))
    W = _bf(torch.randn(N, K, device=DEV) / math.sqrt(K))
    bias = torch.randn(N, device=DEV)
    out = torch.empty(M, N, device=DEV)
    outb = torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
    _ops().gather_gemm(A, W, bias=bias, out_v_f32=out, out_v_bf16=outb)
    ref = A.float() @ W.float().t() + bias
    assert torch.allclose(out, ref, atol=2e-3, rtol=1e-3), (out - ref).abs().max()
    assert torch.allclose(outb.float(), ref, atol=3e-2, rtol=1e-2)


def test_gemm_gelu_residual_ln_rowmap():
    torch.manual_seed(1)
    M, N, K, R = 5000, 256, 512, 9000
    A = _bf(torch.randn(M, K, device=DEV))
    W = _bf(torch.randn(N, K, device=DEV) / math.sqrt(K))
    bias = torch.randn(N, device=DEV)
    g, b = torch.randn(N, device=DEV), torch.randn(N, device=DEV)
    rows = torch.randperm(R, device=DEV)[:M].to(torch.int32)
    rows[17] = -1
    x = torch.randn(R, N, device=DEV)
    x0 = x.clone()
    y = torch.zeros(M, N, device=DEV, dtype=torch.bfloat16)
    xb = torch.zeros(R, N, device=DEV, dtype=torch.bfloat16)
    _ops().gather_gemm(A, W, bias=bias, res=x, out_v_f32=x, out_v_bf16=xb, ln=(g, b),
                       out_y_bf16=y, out_rows=rows)
    v = A.float() @ W.float().t() + bias
    live = rows >= 0
    ref_x = x0.clone()
    ref_x[rows[live].long()] = x0[rows[live].long()] + v[live]
    assert torch.allclose(x, ref_x, atol=2e-3, rtol=1e-3)
    assert torch.allclose(xb.float()[rows[live].long()], ref_x[rows[live].long()], atol=5e-2, rtol=1e-2)
    ref_y = F.layer_norm(ref_x[rows.clamp(min=0).long()], (N,), g, b, 1e-5)
    assert torch.allclose(y.float()[live], ref_y[live], atol=5e-2, rtol=2e-2)
    # GELU + relu-after-LN variants
    out = torch.empty(M, N, device=DEV, dtype=torch.bfloat16)
    _ops().gather_gemm(A, W, bias=bias, act=1, out_v_bf16=out)
    assert torch.allclose(out.float(), F.gelu(v), atol=3e-2, rtol=1e-2)
    yf = torch.empty(M, N, device=DEV)
    _ops().gather_gemm(A, W, ln=(g, b), relu=True, out_y_f32=yf)
    assert torch.allclose(yf, F.relu(F.layer_norm(A.float() @ W.float().t(), (N,), g, b, 1e-5)),
                          atol=5e-3, rtol=1e-3)


@pytest.mark.parametrize('KD,Cin,N', [(27, 64, 64), (8, 32, 64), (27, 128, 128), (8, 256, 256),
                                      (8, 64, 128)])
def test_gemm_gather_is_octree_conv(KD, Cin, N):
    torch.manual_seed(2)
    rows_a, M = 3000, 2500
    A = _bf(torch.randn(rows_a, Cin, device=DEV))
    W3 = torch.randn(KD, Cin, N, device=DEV) / math.sqrt(KD * Cin)
    Wt = _bf(W3.flatten(0, 1).t().contiguous())
    idx = torch.randint(-1, rows_a, (M, KD), device=DEV, dtype=torch.int32)
    idx[idx % 3 == 0] = -1
    out = torch.empty(M, N, device=DEV)
    _ops().gather_gemm(A, Wt, idx=idx, KD=KD, out_v_f32=out)
    buf = A.float()[idx.clamp(min=0).long()] * (idx >= 0).unsqueeze(-1)
    ref = buf.flatten(1) @ Wt.float().t()
    assert torch.allclose(out, ref, atol=3e-3, rtol=1e-3), (out - ref).abs().max()


@pytest.mark.parametrize('C,M,mapped', [(256, 5000, False), (128, 3001, False), (256, 777, True),
                                        (256, 300000, False), (256, 100, False), (128, 1, False),
                                        (128, 128 * 148 + 5, True), (256, 128 * 148 * 2, False)])
def test_mlp_fused(C, M, mapped):
    torch.manual_seed(8)
    y = _bf(torch.randn(M, C, device=DEV))
    W1 = _bf(torch.randn(4 * C, C, device=DEV) / math.sqrt(C))
    W2 = _bf(torch.randn(C, 4 * C, device=DEV) / math.sqrt(4 * C))
    b1, b2 = torch.randn(4 * C, device=DEV) * 0.1, torch.randn(C, device=DEV) * 0.1
    R = M + 500 if mapped else M
    rows = torch.randperm(R, device=DEV)[:M].to(torch.int32) if mapped else None
    x = torch.randn(R, C, device=DEV)
    x0 = x.clone()
    xb = torch.zeros(R, C, device=DEV, dtype=torch.bfloat16)
    _ops().mlp_fused(y, W1, b1, W2, b2, res=x, out_f32=x, out_bf16=xb, out_rows=rows)
    h = _bf(F.gelu(y.float() @ W1.float().t() + b1)).float()          # hidden rounded to bf16 as on chip
    upd = h @ W2.float().t() + b2
    ref = x0.clone()
    if mapped:
        ref[rows.long()] += upd
    else:
        ref += upd
    assert torch.allclose(x, ref, atol=5e-3, rtol=2e-3), (x - ref).abs().max()
    sel = rows.long() if mapped else slice(None)
    assert torch.allclose(xb.float()[sel], ref[sel], atol=5e-2, rtol=1e-2)


@pytest.mark.parametrize('C,M,mapped', [(256, 5000, False), (128, 3001, False), (256, 777, True),
                                        (256, 300000, False), (256, 100, False), (128, 1, False),
                                        (128, 128 * 148 + 5, True), (256, 128 * 148 * 2, False)])
def test_proj_mlp_fused(C, M, mapped):
    """x = x + proj(o); x = x + mlp(norm2(x)) in one kernel vs fp32 torch."""
    torch.manual_seed(9)
    o = _bf(torch.randn(M, C, device=DEV))
    Wp = _bf(torch.randn(C, C, device=DEV) / math.sqrt(C))
    W1 = _bf(torch.randn(4 * C, C, device=DEV) / math.sqrt(C))
    W2 = _bf(torch.randn(C, 4 * C, device=DEV) / math.sqrt(4 * C))
    bp, b1, b2 = (torch.randn(n, device=DEV) * 0.1 for n in (C, 4 * C, C))
    g, b = torch.rand(C, device=DEV) + 0.5, torch.randn(C, device=DEV) * 0.1
    R = M + 500 if mapped else M
    rows = torch.randperm(R, device=DEV)[:M].to(torch.int32) if mapped else None
    x = torch.randn(R, C, device=DEV) * 2 + 0.5
    x0 = x.clone()
    xb = torch.zeros(R, C, device=DEV, dtype=torch.bfloat16)
    _ops().proj_mlp_fused(o, Wp, bp, g, b, W1, b1, W2, b2, res=x, out_f32=x, out_bf16=xb, out_rows=rows)
    sel = rows.long() if mapped else slice(None)
    s = x0[sel] + o.float() @ Wp.float().t() + bp
    y = _bf(F.layer_norm(s, (C,), g, b, 1e-5)).float()                  # operand rounded to bf16 as on chip
    h = _bf(F.gelu(y @ W1.float().t() + b1)).float()
    ref = x0.clone()
    ref[sel] = s + h @ W2.float().t() + b2
    assert torch.allclose(x, ref, atol=1e-2, rtol=2e-3), (x - ref).abs().max()
    assert torch.allclose(xb.float()[sel], ref[sel], atol=5e-2, rtol=1e-2)


def _tokens(n_pad, n, B, K):
    xyz = torch.randint(0, 128, (n_pad, 3), dtype=torch.int16)
    bid = torch.sort(torch.randint(0, B, (n_pad,))).values.to(torch.int16)
    bid[n:] = B
    xyz[n:] = 0
    return torch.cat([xyz, bid[:, None]], 1).contiguous()


def _ref_window_attn(qkv, tok, rpe, K, dil, hat, H, bnd):
    rows, C3 = qkv.shape
    C = C3 // 3
    n_pad = tok.shape[0]
    t = torch.arange(n_pad)
    if hat:
        win_tok = t.view(-1, K)
        W = win_tok.shape[0]
        rowidx = (torch.arange(W)[:, None] * (K + 1) + torch.arange(K + 1)[None])
        ids = torch.cat([tok[win_tok[:, :1], 3], tok[win_tok, 3]], 1).long()
        xyz = torch.cat([torch.zeros(W, 1, 3, dtype=torch.long), tok[win_tok][..., :3].long()], 1)
    else:
        rowidx = t.view(-1, K) if dil == 1 else t.view(-1, K, dil).transpose(1, 2).reshape(-1, K)
        ids = tok[rowidx, 3].long()
        xyz = tok[rowidx][..., :3].long()
    x = qkv.float().cpu()[rowidx]                                        # (W,L,3C)
    W_, L = rowidx.shape
    q, k, v = x.view(W_, L, 3, H, 16).permute(2, 0, 3, 1, 4)
    att = (q @ k.transpose(-1, -2)) * 0.25
    bias = torch.zeros(W_, 1, L, L)
    bias.masked_fill_((ids[:, :, None] != ids[:, None, :])[:, None], float('-inf'))
    if rpe is not None:
        num = 2 * bnd + 1
        rel = (xyz[:, :, None] - xyz[:, None, :]).clamp(-bnd, bnd) + bnd
        r = sum(rpe.cpu()[rel[..., a] + a * num] for a in range(3)).permute(0, 3, 1, 2)
        if hat:
            r[:, :, 0, :] = 0
            r[:, :, :, 0] = 0
        bias = bias + r
    o = ((att + bias).softmax(-1) @ v).transpose(1, 2).reshape(W_, L, C)
    out = torch.zeros(rows, C)
    out[rowidx.reshape(-1)] = o.reshape(-1, C)
    return out


@pytest.mark.parametrize('K,dil,hat,H', [(48, 1, False, 8), (48, 4, False, 8), (48, 1, True, 16),
                                         (64, 1, True, 16), (64, 4, False, 8), (32, 1, True, 8),
                                         (96, 1, True, 16), (96, 4, False, 8), (96, 1, False, 16)])
def test_window_attention(K, dil, hat, H):
    torch.manual_seed(3)
    C = 16 * H
    n_pad, n, B = K * 4 * 5, K * 4 * 5 - 37, 3
    tok = _tokens(n_pad, n, B, K)
    n_win = n_pad // K
    rows = n_win * (K + 1) if hat else n_pad
    qkv = _bf(torch.randn(rows, 3 * C, device=DEV))
    bnd = int(0.8 * K * dil ** 0.5)
    rpe = (torch.randn(3 * (2 * bnd + 1), H) * 0.5).to(DEV)
    out = torch.zeros(rows, C, device=DEV, dtype=torch.bfloat16)
    _ops().window_attn(qkv, out, tok.to(DEV), rpe, n_win, H, C, K, dil, hat, bnd, 0.25)
    ref = _ref_window_attn(qkv, tok, rpe, K, dil, hat, H, bnd)
    err = (out.float().cpu() - ref).abs().max().item()
    assert err < 3e-2, err            # bf16 probabilities / outputs


@pytest.mark.parametrize('K,dil,hat,H', [(48, 1, False, 8), (48, 4, False, 8), (48, 1, True, 16),
                                         (64, 4, False, 8), (32, 1, True, 8), (32, 1, False, 16),
                                         (64, 1, False, 16), (16, 1, True, 8)])
@pytest.mark.parametrize('n_groups', [5, 301])
def test_qkv_attn_fused(K, dil, hat, H, n_groups):
    """fused projection + tensor-core window attention vs fp32 torch on the bf16-rounded qkv."""
    torch.manual_seed(13)
    C = 16 * H
    if not _ops().qkv_attn_supported(H, C, K, dil, hat, int(0.8 * K * dil ** 0.5)):
        pytest.skip('configuration handled by the unfused path')
    n_pad, B = K * 4 * n_groups, 7
    n = n_pad - 37
    tok = _tokens(n_pad, n, B, K)
    n_win = n_pad // K
    rows = n_win * (K + 1) if hat else n_pad
    y = _bf(torch.randn(rows, C, device=DEV))
    W = _bf(torch.randn(3 * C, C, device=DEV) / math.sqrt(C))
    b = torch.randn(3 * C, device=DEV) * 0.2
    bnd = int(0.8 * K * dil ** 0.5)
    rpe = (torch.randn(3 * (2 * bnd + 1), H) * 0.5).to(DEV)
    Wg, bg = _ops().regroup_qkv(W, b)
    out = torch.zeros(rows, C, device=DEV, dtype=torch.bfloat16)
    _ops().qkv_attn(y, Wg, bg, out, tok.to(DEV), rpe, n_win, H, C, K, dil, hat, bnd, 0.25)
    qkv = _bf(y.float() @ W.float().t() + b)
    ref = _ref_window_attn(qkv, tok, rpe, K, dil, hat, H, bnd)
    err = (out.float().cpu() - ref).abs().max().item()
    assert err < 3e-2, err
    # pair codes made once per level (hfl_qkv_attn_codes) instead of per tile: bit-identical output
    codes = _ops().qkv_attn_codes(tok.to(DEV), n_win, K, dil, hat, bnd, True)
    out2 = torch.zeros_like(out)
    _ops().qkv_attn(y, Wg, bg, out2, tok.to(DEV), rpe, n_win, H, C, K, dil, hat, bnd, 0.25, codes=codes)
    assert torch.equal(out2, out)
    # no RPE (disable_RPE)
    out.zero_()
    _ops().qkv_attn(y, Wg, bg, out, tok.to(DEV), None, n_win, H, C, K, dil, hat, bnd, 0.25)
    ref = _ref_window_attn(qkv, tok, None, K, dil, hat, H, bnd)
    err = (out.float().cpu() - ref).abs().max().item()
    assert err < 3e-2, err


def test_varlen_attention():
    torch.manual_seed(4)
    H, C = 16, 256
    lens = [168, 33, 1, 300, 80, 81]
    cu = torch.tensor([0] + list(np.cumsum(lens)), dtype=torch.int32)
    tot = int(cu[-1])
    ids = torch.cat([torch.full((l,), i, dtype=torch.int32) for i, l in enumerate(lens)])
    ids[-5:] = len(lens)                     # padding relay tokens of the last submap
    qkv = _bf(torch.randn(tot, 3 * C, device=DEV))
    out = torch.zeros(tot, C, device=DEV, dtype=torch.bfloat16)
    _ops().varlen_attn(qkv, out, cu.to(DEV), ids.to(DEV), len(lens), max(lens), H, C, 0.25)
    x = qkv.float().cpu()
    for b, l in enumerate(lens):
        s = int(cu[b])
        q, k, v = x[s:s + l].view(l, 3, H, 16).permute(1, 2, 0, 3)
        att = (q @ k.transpose(-1, -2)) * 0.25
        i = ids[s:s + l]
        att.masked_fill_((i[:, None] != i[None, :])[None], float('-inf'))
        ref = (att.softmax(-1) @ v).transpose(0, 1).reshape(l, C)
        err = (out[s:s + l].float().cpu() - ref).abs().max().item()
        assert err < 3e-2, (b, err)


@pytest.mark.parametrize('C,K', [(128, 0), (256, 48), (256, 64)])
def test_cpe_ln(C, K):
    torch.manual_seed(5)
    n = 1000
    n_pad = -(-n // (4 * (K or 48))) * 4 * (K or 48)
    rows = n_pad // K * (K + 1) if K else n_pad
    tok_row = (torch.arange(n_pad) + torch.arange(n_pad) // K + 1) if K else torch.arange(n_pad)
    x = torch.zeros(rows, C)
    x[tok_row[:n]] = torch.randn(n, C)
    if K:
        x[::K + 1] = torch.randn(rows // (K + 1), C)
    ne = torch.randint(-1, n, (n, 27), dtype=torch.int32)
    ne[torch.rand(n, 27) < 0.6] = -1
    w = (torch.randn(27, C) / 5).to(torch.bfloat16).float()
    g_c, b_c, g1, b1 = (torch.randn(C) for _ in range(4))
    xd = x.to(DEV)
    xb = _bf(xd)
    y1 = torch.zeros(rows, C, device=DEV, dtype=torch.bfloat16)
    _ops().cpe_ln(xd, xb, ne.to(DEV), _bf(w.to(DEV)), g_c.to(DEV), b_c.to(DEV), g1.to(DEV), b1.to(DEV),
                  y1, None, n, rows, C, K)
    src = xb.float().cpu()[tok_row[:n]]
    buf = src[ne.clamp(min=0).long()] * (ne >= 0).unsqueeze(-1)
    dw = torch.einsum('ikc,kc->ic', buf, w)
    ref = x.clone()
    ref[tok_row[:n]] += F.layer_norm(dw, (C,), g_c, b_c, 1e-5)
    assert torch.allclose(xd.cpu(), ref, atol=2e-4, rtol=1e-4), (xd.cpu() - ref).abs().max()
    ref_y = F.layer_norm(ref, (C,), g1, b1, 1e-5)
    live = torch.zeros(rows, dtype=torch.bool)
    live[tok_row[:n]] = True
    if K:
        live[::K + 1] = True
    assert torch.allclose(y1.float().cpu()[live], ref_y[live], atol=6e-2, rtol=2e-2)
    # cpe-only mode
    out = torch.zeros(n, C, device=DEV)
    _ops().cpe_ln(xd, xb, ne.to(DEV), _bf(w.to(DEV)), g_c.to(DEV), b_c.to(DEV), None, None, None, out,
                  n, rows, C, K)
    assert torch.allclose(out.cpu(), F.layer_norm(dw, (C,), g_c, b_c, 1e-5), atol=2e-4, rtol=1e-4)


def test_stem_conv_and_ln_rows():
    torch.manual_seed(6)
    n, D = 3000, 9
    pts = torch.rand(n, 3) * 2 ** D
    ne = torch.randint(-1, n, (n, 27), dtype=torch.int32)
    ne[torch.rand(n, 27) < 0.5] = -1
    w = torch.randn(27, 3, 32) / 9
    g, b = torch.randn(32), torch.randn(32)
    out = torch.zeros(n, 32, device=DEV, dtype=torch.bfloat16)
    _ops().stem_conv(pts.to(DEV), ne.to(DEV), n, D, w.flatten(0, 1).contiguous().to(DEV),
                     g.to(DEV), b.to(DEV), out)
    f = pts * 2 ** (1 - D) - 1
    buf = f[ne.clamp(min=0).long()] * (ne >= 0).unsqueeze(-1)
    ref = F.relu(F.layer_norm(buf.flatten(1) @ w.flatten(0, 1), (32,), g, b, 1e-5))
    assert torch.allclose(out.float().cpu(), ref, atol=3e-2, rtol=1e-2)
    x = torch.randn(500, 256)
    rows = torch.randperm(500)[:200].to(torch.int32)
    g, b = torch.randn(256), torch.randn(256)
    y = torch.zeros(200, 256, device=DEV, dtype=torch.bfloat16)
    _ops().ln_rows(x.to(DEV), rows.to(DEV), 200, 256, g.to(DEV), b.to(DEV), y)
    assert torch.allclose(y.float().cpu(), F.layer_norm(x[rows.long()], (256,), g, b, 1e-5),
                          atol=6e-2, rtol=2e-2)


def test_knn_matches_bruteforce():
    torch.manual_seed(7)
    db = F.normalize(torch.randn(3000, 256), dim=1).to(DEV)
    q = F.normalize(torch.randn(700, 256), dim=1).to(DEV)
    d, i = _ops().knn_topk(q, db, 25)
    ref = torch.cdist(q.double(), db.double()) ** 2
    rd, ri = ref.topk(25, dim=1, largest=False)
    assert torch.allclose(d.double(), rd, atol=1e-5)
    # same neighbours up to fp32 near-ties: every returned index is within 1e-6 of the true k-th
    true_d = ref.gather(1, i.long())
    assert bool((true_d <= rd[:, -1:] + 1e-6).all())
    assert (i.long() == ri).float().mean() > 0.999
    assert bool((d[:, 1:] >= d[:, :-1]).all())
    # sharded search + merge == global search
    parts = [(_ops().knn_topk(q, db[s:s + 1000], 25, idx_offset=s)) for s in (0, 1000, 2000)]
    md, mi = _ops().topk_merge(torch.stack([p[0] for p in parts]), torch.stack([p[1] for p in parts]))
    assert torch.equal(mi, i) and torch.equal(md, d)


def test_knn_full_size_with_ties_and_small_query_sets():
    """BASELINE.json configs[3] size (8192 x 8192 x 256) incl. exact duplicates in the database (index ties must
    resolve to the lower index, as a stable sort of the reference's distances does), a query set smaller than one
    tile, a database smaller than k, and the split / unsplit paths agreeing bit for bit."""
    torch.manual_seed(11)
    db = F.normalize(torch.randn(8192, 256), dim=1)
    db[4096:4096 + 512] = db[:512]                       # duplicates: distance ties at different indices
    db = db.to(DEV)
    q = F.normalize(db[torch.randperm(8192)[:8192].to(DEV)] + 0.05 * torch.randn(8192, 256, device=DEV), dim=1)
    d, i = _ops().knn_topk(q, db, 25)
    ref = torch.cdist(q.double(), db.double()) ** 2
    rd, _ = ref.topk(25, dim=1, largest=False)
    assert torch.allclose(d.double(), rd, atol=2e-5)
    true_d = ref.gather(1, i.long())
    assert bool((true_d <= rd[:, -1:] + 1e-6).all())
    assert bool((d[:, 1:] >= d[:, :-1]).all())
    same = d[:, 1:] == d[:, :-1]
    assert bool((i[:, 1:][same] > i[:, :-1][same]).all())  # ties ordered by index
    assert all(len(set(r.tolist())) == 25 for r in i[:64].cpu())
    # unsplit path (no workspace) gives the identical lists
    from hotformerloc_b200 import native as N
    od = torch.empty_like(d); oi = torch.empty_like(i)
    N.check(N.lib().hfl_knn_topk(q.data_ptr(), 8192, db.data_ptr(), 8192, 256, 25, 0, od.data_ptr(), oi.data_ptr(),
                                 torch.cuda.current_stream().cuda_stream))
    assert torch.equal(od, d) and torch.equal(oi, i)
    # 5 queries (one partial tile, many database splits) and a database smaller than k
    d5, i5 = _ops().knn_topk(q[:5], db, 25)
    assert torch.equal(d5, d[:5]) and torch.equal(i5, i[:5])
    d3, i3 = _ops().knn_topk(q[:70], db[:10], 25)
    assert bool((i3[:, 10:] == -1).all()) and bool(torch.isinf(d3[:, 10:]).all())
    assert torch.equal(i3[:, :10].long().sort(1).values, torch.arange(10, device=DEV).expand(70, 10))


# ---------------------------------------------------------------------------
# relay-token / pooling-head kernels in isolation
# ---------------------------------------------------------------------------
def _hat(n_tok, K):
    """hat-layout rows of tokens 0..n_tok-1 (relay token first in every window)."""
    t = torch.arange(n_tok)
    return t + t // K + 1


@pytest.mark.parametrize('name,K', [('cswp_b6_stress', 64), ('oxford_b4_init', 48)])
def test_rt_init_vs_reference_octree_t(name, K):
    """hfl_rt_init: masked window mean + ADaPE window statistics + fc1/GELU against the values the
    reference's own OctreeT produced (tests/golden/octree_t.npz: rt_init_mask, window_stats) --
    includes windows shared by several submaps and pure-padding windows."""
    from oracle.make_golden import CASES
    from tests.common import GOLDEN, case_clouds
    from hotformerloc_b200.octree import build_batch
    cfg, depth, spec, seed, mode = CASES[name]
    gold = np.load(os.path.join(GOLDEN, 'octree_t.npz'))
    octree = build_batch(case_clouds(name), depth, 2, 'cuda').finalize()
    C, d0 = 256, depth - 2
    torch.manual_seed(14)
    w1, b1 = torch.randn(C, 9, device=DEV) * 0.5, torch.randn(C, device=DEV) * 0.1
    for d in (d0 - 1, d0 - 2, d0 - 3):
        n = octree.n(d)
        npad = int(gold[f'{name}_nnum_a_{d}'])
        n_win = npad // K
        tok = octree.tokens(d, npad)
        x = torch.zeros(n_win * (K + 1), C, device=DEV)
        feat = torch.randn(n, C, device=DEV)
        x[_hat(n, K).to(DEV)] = feat
        h = torch.zeros(n_win, C, device=DEV, dtype=torch.bfloat16)
        stats = torch.zeros(n_win, 9, device=DEV)
        _ops().rt_init(x, None, tok, n, n_win, K, C, d, 9, w1, b1, h, stats_out=stats)
        # reference: mean over the tokens the reference's rt_init_mask keeps (padding tokens are zeros)
        mask = np.unpackbits(gold[f'{name}_rt_init_mask_{d}'])[:n_win * K].reshape(n_win, K).astype(bool)
        keep = torch.from_numpy(~mask).to(DEV)
        xp = torch.cat([feat, feat.new_zeros(npad - n, C)]).view(n_win, K, C)
        ref_rt = (xp * keep.unsqueeze(-1)).sum(1) / keep.sum(1, keepdim=True).clamp(min=1)
        got_rt = x[::K + 1]
        assert torch.allclose(got_rt, ref_rt, atol=1e-5, rtol=1e-5), (d, (got_rt - ref_rt).abs().max())
        ref_stats = torch.from_numpy(gold[f'{name}_window_stats_{d}']).to(DEV)
        assert torch.allclose(stats, ref_stats, atol=2e-5, rtol=1e-4), (d, (stats - ref_stats).abs().max())
        ref_h = F.gelu(ref_stats @ w1.t() + b1)
        assert torch.allclose(h.float(), ref_h, atol=3e-2, rtol=2e-2)


def test_attn_pool_vs_torch():
    """hfl_attn_pool (AdaptivePooling, salsa.py:25-55 restricted to each submap's tokens) against
    softmax(q x^T / sqrt(C)) x per submap, including an EMPTY submap."""
    torch.manual_seed(15)
    B, K, C, kq, ktot, q_off = 5, 48, 256, 74, 128, 18
    counts = [700, 33, 0, 1500, 259]
    n = sum(counts)
    n_win = -(-n // (4 * K)) * 4
    rows = n_win * (K + 1)
    x = torch.zeros(rows, C, device=DEV)
    feat = torch.randn(n, C, device=DEV)
    x[_hat(n, K).to(DEV)] = feat
    xb = _bf(x)
    q = torch.randn(kq, C, device=DEV)
    npd = (kq + 63) // 64 * 64
    qp = torch.zeros(npd, C, device=DEV, dtype=torch.bfloat16)
    qp[:kq] = _bf(q)
    logits = torch.empty(rows, npd, device=DEV)
    _ops().gather_gemm(xb, qp, out_v_f32=logits)
    tok_off = torch.tensor([0] + list(np.cumsum(counts)), dtype=torch.int32, device=DEV)
    stat = torch.zeros(B, kq, 2, device=DEV)
    out = torch.full((B, ktot, C), 7.0, device=DEV)
    _ops().attn_pool(logits, x, xb, tok_off, stat, out, B, kq, npd, K, C, ktot, q_off, C ** -0.5)
    fb, qb = xb.float()[_hat(n, K).to(DEV)], qp[:kq].float()
    o = 0
    for b, c in enumerate(counts):
        seg = fb[o:o + c]
        o += c
        if c == 0:
            assert torch.isfinite(out[b]).all()
            continue
        ref = ((qb @ seg.t()) * C ** -0.5).softmax(-1) @ seg
        got = out[b, q_off:q_off + kq]
        assert torch.allclose(got, ref, atol=2e-2, rtol=2e-2), (b, (got - ref).abs().max())
    assert bool((out[:, :q_off] == 7.0).all()) and bool((out[:, q_off + kq:] == 7.0).all())


@pytest.mark.parametrize('ktot,kout,od', [(128, 32, 8), (256, 64, 4)])
def test_mixer_tail_vs_torch(ktot, kout, od):
    """hfl_mixer_tail: channel_proj over the token axis, row_proj, flatten, L2-normalise (salsa.py:103-111)."""
    torch.manual_seed(16)
    B, C = 9, 256
    x = torch.randn(B, ktot, C, device=DEV)
    wc, bc = torch.randn(kout, ktot, device=DEV) / math.sqrt(ktot), torch.randn(kout, device=DEV) * 0.1
    wr, br = torch.randn(od, C, device=DEV) / math.sqrt(C), torch.randn(od, device=DEV) * 0.1
    for normalize in (True, False):
        out = torch.zeros(B, kout * od, device=DEV)
        _ops().mixer_tail(x, wc, bc, wr, br, out, B, ktot, kout, C, od, normalize)
        y = F.linear(x.permute(0, 2, 1), wc, bc).permute(0, 2, 1)
        ref = F.linear(y, wr, br).flatten(1)
        if normalize:
            ref = F.normalize(ref, dim=1)
        assert torch.allclose(out, ref, atol=2e-4, rtol=1e-3), (out - ref).abs().max()


def test_gem_pool_and_head_vs_torch():
    """hfl_gem_pool + hfl_gem_head (PyramidOctGeMWrapper.forward, eval mode, pooling.py:87-103)."""
    torch.manual_seed(17)
    B, K, C, L = 4, 48, 256, 3
    pooled = torch.zeros(B, L * C, device=DEV)
    refs = []
    for j, (counts, pw) in enumerate(zip(([300, 0, 77, 1000], [64, 1, 5, 200], [9, 9, 9, 9]), (3.0, 2.5, 1.7))):
        n = sum(counts)
        n_win = -(-max(n, 1) // (4 * K)) * 4
        x = torch.zeros(n_win * (K + 1), C, device=DEV)
        feat = torch.randn(n, C, device=DEV)
        x[_hat(n, K).to(DEV)] = feat
        tok_off = torch.tensor([0] + list(np.cumsum(counts)), dtype=torch.int32, device=DEV)
        _ops().gem_pool(x, tok_off, B, K, C, pw, 1e-6, pooled, L * C, j * C)
        t = feat.clamp(min=1e-6).pow(pw)
        ref = torch.zeros(B, C, device=DEV)
        o = 0
        for b, c in enumerate(counts):
            if c:
                ref[b] = t[o:o + c].mean(0)
            o += c
        refs.append(ref.pow(1.0 / pw))
    ref_pooled = torch.cat(refs, 1)
    assert torch.allclose(pooled, ref_pooled, atol=1e-4, rtol=1e-3), (pooled - ref_pooled).abs().max()
    w = torch.randn(256, L * C, device=DEV) / math.sqrt(L * C)
    g, b, mu, var = (torch.rand(256, device=DEV) + 0.5, torch.randn(256, device=DEV) * 0.1,
                     torch.randn(256, device=DEV) * 0.1, torch.rand(256, device=DEV) + 0.5)
    out = torch.zeros(B, 256, device=DEV)
    _ops().gem_head(pooled, w, g, b, mu, var, 1e-5, True, out)
    ref = F.normalize((pooled @ w.t() - mu) / torch.sqrt(var + 1e-5) * g + b, dim=1)
    assert torch.allclose(out, ref, atol=2e-4, rtol=1e-3), (out - ref).abs().max()


def test_hat_rows_and_remap():
    K = 48
    n = 1000
    r = torch.zeros(n, device=DEV, dtype=torch.int32)
    _ops().hat_rows(r, n, K, 5)
    assert torch.equal(r.cpu(), (_hat(n, K) + 5).to(torch.int32))
    src = torch.randint(-1, n, (300, 8), dtype=torch.int32)
    out = torch.zeros_like(src, device=DEV)
    _ops().remap_hat(src.to(DEV), out, src.numel(), K)
    ref = torch.where(src < 0, src, (src + src // K + 1))
    assert torch.equal(out.cpu(), ref)
