"""ORACLE (test infrastructure, not product code).

Functional CPU (torch fp32) restatement of the reference's embedding hot path
``HOTFormerLoc.forward(batch)`` for the shipped configurations -- everything
*above* the ocnn boundary (which is ``oracle/octree_ref.py``).  It is written
against a plain ``state_dict`` (reference key layout) and a
:class:`oracle.octree_ref.RefOctree`, using index arithmetic (SURVEY.md
Appendix B) instead of the reference's dense mask tensors.

Parity status: PINNED against the reference's own, unmodified Python
(``/root/reference/models/*.py`` run over ``oracle/ocnn_standin.py``) by
``oracle/make_golden.py`` in the authoring container; the resulting
descriptors are frozen in ``tests/golden/descriptors_*.npz`` and re-checked by
``tests/test_oracle_model.py`` (max-abs 1e-5 in fp32).

Reference anchors (file:line under /root/reference):
  forward              models/hotformerloc.py:33-59
  backbone             models/hotformerloc_backbone.py:702-723, 574-635, 540-572
  blocks               models/octformer_backbone.py:52-93, 251-299, 451-477
                       models/hotformerloc_backbone.py:83-119, 197-236, 275-295, 345-363
  window bookkeeping   models/octree.py:73-75, 130-184, 229-344
  RT concat / split    models/relay_token_utils.py:12-79
  layers               models/layers/octformer_layers.py:38-59, 80-98, 122-170, 177-210
  pooling head         models/layers/pooling.py:183-233, models/layers/salsa.py:25-55, 103-111

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this module.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import numpy as np
import torch
import torch.nn.functional as F

from . import octree_ref as R


@dataclass
class HParams:
    """The [MODEL] keys of models/*_cfg.txt that reach the forward pass."""
    channels: Sequence[int] = (128, 256)
    num_blocks: Sequence[int] = (4, 10)
    num_heads: Sequence[int] = (8, 16)
    num_pyramid_levels: int = 3
    num_octf_levels: int = 1
    patch_size: int = 48
    dilation: int = 4
    stem_down: int = 2
    ADaPE_mode: Optional[str] = 'cov'
    k_pooled_tokens: Sequence[int] = (74, 36, 18)
    output_dim: int = 256
    normalize_embeddings: bool = True
    pooling: str = 'PyramidAttnPoolMixer'

    @staticmethod
    def from_cfg(path: str) -> 'HParams':
        import configparser
        cp = configparser.ConfigParser()
        cp.read(path)
        p = cp['MODEL']
        tup = lambda s: tuple(int(e) for e in s.split(','))
        adape = p.get('ADaPE_mode', None)
        return HParams(
            channels=tup(p['channels']), num_blocks=tup(p['num_blocks']),
            num_heads=tup(p['num_heads']),
            num_pyramid_levels=p.getint('num_pyramid_levels', 3),
            num_octf_levels=p.getint('num_octf_levels', 1),
            patch_size=p.getint('patch_size', 32), dilation=p.getint('dilation', 4),
            stem_down=p.getint('num_input_downsamples', 2),
            ADaPE_mode=None if adape in (None, 'None') else adape,
            k_pooled_tokens=tup(p.get('k_pooled_tokens', '64')),
            output_dim=p.getint('output_dim', 256),
            normalize_embeddings=p.getboolean('normalize_embeddings', False),
            pooling=p.get('pooling', 'OctGeM'))


# ----------------------------------------------------------------------------
# window bookkeeping (models/octree.py)
# ----------------------------------------------------------------------------
@dataclass
class Level:
    depth: int
    n: int                       # non-empty nodes (all submaps)
    n_pad: int                   # nnum_a
    bid: torch.Tensor            # (n_pad,) submap id, padding = B
    xyz: torch.Tensor            # (n_pad,3) integer cell coords, padding = 0
    counts: np.ndarray           # (B,) nodes per submap
    num_windows: Optional[np.ndarray] = None   # (B,) RTs owned per submap


def make_level(oct: R.RefOctree, d: int, K: int, dil: int) -> Level:
    B = oct.batch_size
    n = int(oct.nnum_nempty[d])
    blk = K * dil
    n_pad = -(-n // blk) * blk
    key = oct.key(d, nempty=True)
    x, y, z, b = R.key2xyz(key, d)
    bid = np.full(n_pad, B, dtype=np.int64)
    bid[:n] = b
    xyz = np.zeros((n_pad, 3), dtype=np.int64)
    xyz[:n] = np.stack([x, y, z], 1)
    counts = np.asarray(oct.batch_nnum_nempty[d], dtype=np.int64)
    cum = np.cumsum(counts)
    cum[-1] += n_pad - n
    boundary = -(-cum // K)                                   # ceil
    nw = np.diff(np.concatenate([[0], boundary]))
    return Level(d, n, n_pad, torch.from_numpy(bid), torch.from_numpy(xyz), counts, nw)


def window_rows(n_pad: int, K: int, dil: int) -> torch.Tensor:
    """(N_win, K) row indices into the padded token array for each window."""
    t = torch.arange(n_pad)
    if dil == 1:
        return t.view(-1, K)
    return t.view(-1, K, dil).transpose(1, 2).reshape(-1, K)


# ----------------------------------------------------------------------------
# layers
# ----------------------------------------------------------------------------
def _ln(x, sd, p):
    return F.layer_norm(x, (x.shape[-1],), sd[p + '.weight'], sd[p + '.bias'], 1e-5)


def _lin(x, sd, p):
    return F.linear(x, sd[p + '.weight'], sd.get(p + '.bias'))


def _mlp(x, sd, p):
    return _lin(F.gelu(_lin(x, sd, p + '.fc1')), sd, p + '.fc2')


def _gather(data, neigh):
    idx = neigh.clamp(min=0)
    buf = data[idx]                                         # (rows,kdim,C)
    return buf * (neigh >= 0).unsqueeze(-1).to(data.dtype)


def octree_conv(data, neigh, w, bias=None):
    out = _gather(data, neigh).flatten(1) @ w.flatten(0, 1)
    return out if bias is None else out + bias


def octree_dwconv(data, neigh, w):
    return torch.einsum('ikc,kc->ic', _gather(data, neigh), w[:, 0, :])


class _Ctx:
    def __init__(self, sd, oct: R.RefOctree, hp: HParams):
        self.sd, self.oct, self.hp = sd, oct, hp
        self._neigh = {}

    def neigh(self, d, kernel, stride):
        k = (d, kernel, stride)
        if k not in self._neigh:
            self._neigh[k] = torch.from_numpy(self.oct.get_neigh(d, kernel, stride, nempty=True))
        return self._neigh[k]

    def conv_norm_relu(self, x, p, d, kernel, stride):
        y = octree_conv(x, self.neigh(d, kernel, stride), self.sd[p + '.conv.weights'])
        return F.relu(_ln(y, self.sd, p + '.norm'))

    def downsample(self, x, p, d):
        y = octree_conv(x, self.neigh(d, '222', 2), self.sd[p + '.conv.weights'],
                        self.sd[p + '.conv.bias'])
        return _ln(y, self.sd, p + '.norm')

    def cpe(self, x, p, d):
        return _ln(octree_dwconv(x, self.neigh(d, '333', 1), self.sd[p + '.conv.weights']),
                   self.sd, p + '.norm')


def _rpe_bias(sd, p, xyz_w, K, dil):
    """(N_win,H,K,K) relative position bias (octformer_layers.py:156-170)."""
    table = sd[p + '.rpe_table']
    bnd = int(0.8 * K * dil ** 0.5)
    num = 2 * bnd + 1
    rel = (xyz_w.unsqueeze(2) - xyz_w.unsqueeze(1)).clamp(-bnd, bnd) + bnd   # (N,K,K,3)
    out = 0
    for a in range(3):
        out = out + table[rel[..., a] + a * num]               # (N,K,K,H)
    return out.permute(0, 3, 1, 2)


def _attention(x, ids, sd, p, H, rpe=None):
    """x: (N,L,C), ids: (N,L) tokens attend iff ids equal. rpe: (N,H,L,L)|None."""
    N, L, C = x.shape
    qkv = _lin(x, sd, p + '.qkv').view(N, L, 3, H, C // H).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0], qkv[1], qkv[2]
    bias = torch.zeros(N, 1, L, L)
    bias.masked_fill_((ids.unsqueeze(2) != ids.unsqueeze(1)).unsqueeze(1), -1000.0)
    if rpe is not None:
        bias = bias + rpe
    att = (q @ k.transpose(-1, -2)) * ((C // H) ** -0.5) + bias
    out = (att.softmax(-1) @ v).transpose(1, 2).reshape(N, L, C)
    return _lin(out, sd, p + '.proj')


def octformer_block(cx: _Ctx, x, p, lv: Level, H, K, dil):
    """OctFormerBlock.forward, use_rt=False (octformer_backbone.py:251-299)."""
    sd = cx.sd
    x = x + cx.cpe(x, p + '.cpe', lv.depth)
    rows = window_rows(lv.n_pad, K, dil)
    xp = torch.cat([x, x.new_zeros(lv.n_pad - lv.n, x.shape[1])])
    w = xp[rows]                                              # (N,K,C)
    rpe = _rpe_bias(sd, p + '.attention.rpe', lv.xyz[rows], K, dil)
    w = w + _attention(_ln(w, sd, p + '.norm1'), lv.bid[rows], sd, p + '.attention', H, rpe)
    w = w + _mlp(_ln(w, sd, p + '.norm2'), sd, p + '.mlp')
    xp = torch.empty_like(xp)
    xp[rows.reshape(-1)] = w.reshape(-1, x.shape[1])
    return xp[:lv.n]


def hosa_block(cx: _Ctx, x, rt, p, lv: Level, H, K):
    """HOTFormerBlock.forward (hotformerloc_backbone.py:197-236)."""
    sd = cx.sd
    C = x.shape[1]
    x = x + cx.cpe(x, p + '.cpe', lv.depth)
    xp = torch.cat([x, x.new_zeros(lv.n_pad - lv.n, C)]).view(-1, K, C)
    bid = lv.bid.view(-1, K)
    w = torch.cat([rt.unsqueeze(1), xp], 1)                   # (N,K+1,C)
    ids = torch.cat([bid[:, :1], bid], 1)                     # RT id = min = first token's
    rpe = _rpe_bias(sd, p + '.attention.rpe', lv.xyz.view(-1, K, 3), K, 1)
    rpe = F.pad(rpe, (1, 0, 1, 0))
    w = w + _attention(_ln(w, sd, p + '.norm1'), ids, sd, p + '.attention', H, rpe)
    w = w + _mlp(_ln(w, sd, p + '.norm2'), sd, p + '.mlp')
    return w[:, 1:].reshape(-1, C)[:lv.n], w[:, 0]


def rt_layout(levels: List[Level], B: int):
    """Per-submap RT gather plan (relay_token_utils.py:12-40, octree.py:229-265).
    Returns index (B,Nmax) into the level-concatenated RT array (-1 = pad) and
    the attention ids (B,Nmax)."""
    nws = [lv.num_windows for lv in levels]
    tot = np.sum(nws, 0)
    Nmax = int(tot.max())
    index = np.full((B, Nmax), -1, dtype=np.int64)
    ids = np.full((B, Nmax), 10000, dtype=np.int64)
    base = 0
    fill = np.zeros(B, dtype=np.int64)
    for lv, nw in zip(levels, nws):
        start = np.cumsum(nw) - nw
        for b in range(B):
            index[b, fill[b]:fill[b] + nw[b]] = base + start[b] + np.arange(nw[b])
            ids[b, fill[b]:fill[b] + nw[b]] = b
        fill += nw
        base += int(nw.sum())
    prev = 0
    for lv, nw in zip(levels, nws):                     # padding windows of the last submap
        K = lv.n_pad // int(nw.sum())
        npad_tok = int((lv.bid.view(-1, K)[:, 0] >= B).sum())
        rel = int(nw[-1])
        if npad_tok > 0:
            end = prev + rel
            ids[-1, np.arange(Nmax)[end - npad_tok:end]] = B
        prev += rel
    return torch.from_numpy(index), torch.from_numpy(ids)


def rtsa_block(cx: _Ctx, rts: List[torch.Tensor], p, index, ids, H):
    """RelayTokenTransformerBlock.forward (hotformerloc_backbone.py:275-295)."""
    sd = cx.sd
    cat = torch.cat(rts)
    C = cat.shape[1]
    x = cat[index.clamp(min=0)] * (index >= 0).unsqueeze(-1)
    x = x + _attention(_ln(x, sd, p + '.norm1'), ids, sd, p + '.rt_attention', H)
    x = x + _mlp(_ln(x, sd, p + '.norm2'), sd, p + '.mlp')
    out = torch.empty_like(cat)
    m = index >= 0
    out[index[m]] = x[m]
    return list(out.split([r.shape[0] for r in rts]))


def rt_init(x, lv: Level, K):
    """RelayTokenInitialiser.forward (hotformerloc_backbone.py:345-363)."""
    C = x.shape[1]
    xp = torch.cat([x, x.new_zeros(lv.n_pad - lv.n, C)]).view(-1, K, C)
    bid = lv.bid.view(-1, K)
    valid = (bid == bid[:, :1]).unsqueeze(-1).to(x.dtype)
    return (xp * valid).sum(1) / valid.sum(1)


def window_stats(lv: Level, K, mode='cov'):
    """OctreeT.compute_window_stats (models/octree.py:285-344)."""
    d = lv.depth
    pts = lv.xyz.to(torch.float32) * (2 ** (1 - d)) - 1.0
    pts[lv.n:] = 0.0                                           # padded AFTER rescale
    pts = pts.view(-1, K, 3)
    bid = lv.bid.view(-1, K)
    valid = (bid == bid[:, :1])
    cnt = valid.sum(1, keepdim=True).float()
    v = valid.unsqueeze(-1).float()
    mu = (pts * v).sum(1) / cnt.clamp(min=1.0)
    if mode == 'pos':
        return mu
    cen = (pts - mu.unsqueeze(1)) * v
    den = (cnt - 1).clamp(min=1.0)
    ok = (cnt >= 2).float()
    if mode == 'var':
        return torch.cat([mu, (cen ** 2).sum(1) / den * ok], 1)
    cov = torch.bmm(cen.transpose(1, 2), cen) / den.unsqueeze(-1) * ok.unsqueeze(-1)
    iu = torch.triu_indices(3, 3)
    return torch.cat([mu, cov[:, iu[0], iu[1]]], 1)


def attn_pool_mixer(cx: _Ctx, feats: List[torch.Tensor], levels: List[Level]):
    """PyramidAttnPoolWrapper.forward + Mixer (pooling.py:183-233, salsa.py)."""
    sd, B = cx.sd, cx.oct.batch_size
    P = 'pooling.pooling'
    toks = []
    for j, (x, lv) in enumerate(zip(feats, levels)):
        q = sd[f'{P}.attpool.{j}.query']
        outs = []
        for seg in x.split([int(c) for c in lv.counts]):
            att = (q @ seg.t()) * (q.shape[1] ** -0.5)
            outs.append(att.softmax(-1) @ seg)
        toks.append(torch.stack(outs))
    x = torch.cat(toks, 1)                                     # (B,ktot,C)
    D = f'{P}.descriptor_extractor'
    l = 0
    while f'{D}.mix.{l}.mix.0.weight' in sd:
        h = _ln(x, sd, f'{D}.mix.{l}.mix.0')
        x = x + _lin(F.gelu(_lin(h, sd, f'{D}.mix.{l}.mix.1')), sd, f'{D}.mix.{l}.mix.3')
        l += 1
    x = _lin(x.permute(0, 2, 1), sd, f'{D}.channel_proj').permute(0, 2, 1)
    x = _lin(x, sd, f'{D}.row_proj')
    return x.flatten(1)


def pyramid_gem(cx: _Ctx, feats, levels, eps=1e-6):
    """PyramidOctGeMWrapper.forward, eval mode (pooling.py:87-103)."""
    sd, B = cx.sd, cx.oct.batch_size
    P = 'pooling.pooling'
    descs = []
    for j, (x, lv) in enumerate(zip(feats, levels)):
        p = sd[f'{P}.p'][j]
        t = x.clamp(min=eps).pow(p)
        bid = lv.bid[:lv.n]
        s = t.new_zeros(B, t.shape[1]).index_add_(0, bid, t)
        cnt = torch.bincount(bid, minlength=B).clamp(min=1).to(t.dtype)
        descs.append((s / cnt[:, None]).pow(1.0 / p))
    g = F.linear(torch.cat(descs, -1), sd[f'{P}.linear_bn.0.weight'])
    bn = f'{P}.linear_bn.1'
    return (g - sd[bn + '.running_mean']) / torch.sqrt(sd[bn + '.running_var'] + 1e-5) \
        * sd[bn + '.weight'] + sd[bn + '.bias']


# ----------------------------------------------------------------------------
# full forward
# ----------------------------------------------------------------------------
@torch.inference_mode()
def forward(sd: Dict[str, torch.Tensor], oct: R.RefOctree, hp: HParams,
            return_intermediates: bool = False):
    """HOTFormerLoc.forward -> (B, output_dim) float32 tensor."""
    cx = _Ctx(sd, oct, hp)
    B, D, K = oct.batch_size, oct.depth, hp.patch_size
    BB = 'backbone.backbone'
    inter = {}
    x = torch.from_numpy(oct.input_feature_P())
    # PatchEmbed (octformer_backbone.py:451-461)
    d = D
    for i in range(hp.stem_down):
        x = cx.conv_norm_relu(x, f'{BB}.patch_embed.convs.{i}', d, '333', 1)
        x = cx.conv_norm_relu(x, f'{BB}.patch_embed.downsamples.{i}', d, '222', 2)
        d -= 1
    x = cx.conv_norm_relu(x, f'{BB}.patch_embed.proj', d, '333', 1)
    inter['stem'] = x
    # OctFormer stage(s) + downsample (hotformerloc_backbone.py:715-718)
    for i in range(hp.num_octf_levels):
        lv = make_level(oct, d, K, hp.dilation)
        for blk in range(hp.num_blocks[i]):
            dil = 1 if blk % 2 == 0 else hp.dilation
            x = octformer_block(cx, x, f'{BB}.octf_stage.{i}.blocks.{blk}', lv,
                                hp.num_heads[i], K, dil)
        inter[f'octf{i}'] = x
        x = cx.downsample(x, f'{BB}.downsample.{i}', d)
        d -= 1
    # HOTFormer stage (hotformerloc_backbone.py:574-635)
    HS = f'{BB}.hotf_stage'
    H = hp.num_heads[-1]
    L = hp.num_pyramid_levels
    levels = [make_level(oct, d - j, K, hp.dilation) for j in range(L)]
    feats, rts = [x], []
    for j, lv in enumerate(levels):
        src = feats[j]
        if hp.ADaPE_mode is None:
            src = cx.cpe(src, f'{HS}.relay_tokeniser.cpe', lv.depth)
        rt = rt_init(src, lv, K)
        if hp.ADaPE_mode is not None:
            rt = rt + _mlp(window_stats(lv, K, hp.ADaPE_mode), sd, f'{HS}.rt_adape.mlp')
        rts.append(rt)
        if j < L - 1:
            feats.append(cx.downsample(feats[j], f'{HS}.downsamples.{j}', lv.depth))
    inter['rt_init'] = [r.clone() for r in rts]
    index, ids = rt_layout(levels, B)
    for i in range(hp.num_blocks[-1]):
        rts = rtsa_block(cx, rts, f'{HS}.rtsa_blocks.{i}', index, ids, H)
        for j, lv in enumerate(levels):
            feats[j], rts[j] = hosa_block(cx, feats[j], rts[j],
                                          f'{HS}.hosa_blocks.{j}.{i}', lv, H, K)
    inter['feats'] = feats
    inter['rts'] = rts
    if hp.pooling == 'PyramidAttnPoolMixer':
        g = attn_pool_mixer(cx, feats, levels)
    elif hp.pooling == 'PyramidOctGeM':
        g = pyramid_gem(cx, feats, levels)
    else:
        raise NotImplementedError(hp.pooling)
    if hp.normalize_embeddings:
        g = F.normalize(g, dim=1)
    return (g, inter) if return_intermediates else g


# ----------------------------------------------------------------------------
# deterministic synthetic inputs / weights shared by oracle, tests and bench
# ----------------------------------------------------------------------------
# the synthetic cloud generator lives with the other workload generators (shared by bench.py, the
# tools and the tests); re-exported here for the parity tests
from hotformerloc_b200.datasets.synthetic import lidar_cloud  # noqa: E402,F401


def synthetic_state_dict(shapes: Dict[str, Sequence[int]], seed: int = 0,
                         mode: str = 'init'):
    """Name-keyed deterministic weights (independent of module construction
    order so that both sides of a parity test can regenerate them): each
    tensor is drawn from its own generator seeded by crc32(name).

    mode='init'   : the reference's initialisation *distributions*
                    (hotformerloc_backbone.py:817-843 Linear trunc_normal .02 /
                    bias 0; octformer_layers.py:153-154 RPE trunc_normal .02;
                    ocnn conv xavier_uniform; LayerNorm 1/0; salsa.py:21 randn).
    mode='stress' : unit-gain weights, non-zero biases and non-unit norm gains,
                    so that every term of every layer is exercised."""
    import zlib
    sd = {}
    for name, shape in shapes.items():
        g = torch.Generator().manual_seed((zlib.crc32(name.encode()) + seed) & 0x7fffffff)
        shape = tuple(shape)
        leaf = name.rsplit('.', 1)[-1]
        if name.endswith('num_batches_tracked'):
            sd[name] = torch.zeros(shape, dtype=torch.long)
            continue
        t = torch.randn(shape, generator=g)
        stress = mode == 'stress'
        if leaf == 'weights':                                  # octree conv / dwconv
            if stress:
                kdim, cin = shape[0], shape[1]
                t = t * (1.0 / math.sqrt(kdim * cin))
            else:
                fan_in, fan_out = shape[1] * shape[2], shape[0] * shape[2]
                bound = math.sqrt(6.0 / (fan_in + fan_out))
                t = (torch.rand(shape, generator=g) * 2 - 1) * bound
        elif leaf == 'query':
            pass
        elif leaf == 'rpe_table':
            t = t.clamp(-2, 2) * (0.2 if stress else 0.02)
        elif leaf == 'p':
            t = 3.0 + (0.1 * t if stress else 0 * t)
        elif leaf == 'running_var':
            t = 1.0 + (0.1 * t.abs() if stress else 0 * t)
        elif leaf == 'running_mean':
            t = 0.05 * t if stress else 0 * t
        elif len(shape) >= 2:                                  # linear weight (out,in)
            t = t * (1.0 / math.sqrt(shape[1])) if stress else t.clamp(-2, 2) * 0.02
        elif leaf == 'weight':                                 # norm gain
            t = 1.0 + (0.1 * t if stress else 0 * t)
        else:                                                  # biases
            t = 0.05 * t if stress else 0 * t
        sd[name] = t.contiguous()
    return sd
