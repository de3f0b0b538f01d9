"""ORACLE (test infrastructure, not product code).

CPU/numpy restatement of the octree arithmetic that the reference delegates to
the third-party package ``ocnn==2.2.2`` (pinned in reference
``requirements.txt:7``; *not* vendored in /root/reference, not installable
offline).  The algorithm is restated from SURVEY.md Appendix A and anchored on
the reference's own call sites:

* ``Octree.build_octree``     <- eval/pnv_evaluate.py:173-175, datasets/dataset_utils.py:90-92
* ``merge_octrees``           <- eval/pnv_evaluate.py:123
* ``construct_all_neigh``     <- misc/torch_utils.py:49-51
* ``get_neigh``               <- libs/dwconv/dwconv/nn.py:59, models/octree.py:95-110
* ``xyz2key/key2xyz``         <- models/octree.py:273-275, 298

Parity status: PINNED.  ``tests/test_oracle_octree.py`` checks every function
below bit-for-bit against the reference's golden fixtures
``libs/dwconv/test/data/octree/test_00{1..5}.npz`` (key, child, nnum,
nnum_nempty) and ``libs/dwconv/test/data/batch_45.npz`` (merged key, child,
nnum, nnum_nempty and the 27-neighbour table), re-packed into
``tests/golden/octree_fixtures.npz`` by ``oracle/make_golden.py``.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this module.
"""
from __future__ import annotations

import numpy as np

BATCH_SHIFT = 48
KEY_MASK = (1 << BATCH_SHIFT) - 1


# ----------------------------------------------------------------------------
# Morton keys
# ----------------------------------------------------------------------------
def xyz2key(x, y, z, b=None, depth: int = 16) -> np.ndarray:
    """Interleave the low ``depth`` bits of x,y,z (x most significant of each
    triple); batch id goes to bits >= 48.  (ocnn.octree.shuffled_key.xyz2key)"""
    x = np.asarray(x, dtype=np.int64)
    y = np.asarray(y, dtype=np.int64)
    z = np.asarray(z, dtype=np.int64)
    key = np.zeros_like(x)
    for i in range(depth):
        key |= ((x >> i) & 1) << (3 * i + 2)
        key |= ((y >> i) & 1) << (3 * i + 1)
        key |= ((z >> i) & 1) << (3 * i + 0)
    if b is not None:
        key = key | (np.asarray(b, dtype=np.int64) << BATCH_SHIFT)
    return key


def key2xyz(key, depth: int = 16):
    """Inverse of :func:`xyz2key` -> (x, y, z, b)."""
    key = np.asarray(key, dtype=np.int64)
    b = key >> BATCH_SHIFT
    k = key & KEY_MASK
    x = np.zeros_like(k)
    y = np.zeros_like(k)
    z = np.zeros_like(k)
    for i in range(depth):
        x |= ((k >> (3 * i + 2)) & 1) << i
        y |= ((k >> (3 * i + 1)) & 1) << i
        z |= ((k >> (3 * i + 0)) & 1) << i
    return x, y, z, b


# ----------------------------------------------------------------------------
# Neighbour look-up tables (ocnn.octree.Octree.construct_neigh LUTs)
# ----------------------------------------------------------------------------
def _make_luts():
    d = np.array([(i, j, k) for i in (-1, 0, 1) for j in (-1, 0, 1) for k in (-1, 0, 1)],
                 dtype=np.int64)                       # (27,3) x-major
    c = np.array([(i, j, k) for i in (2, 3) for j in (2, 3) for k in (2, 3)],
                 dtype=np.int64)                       # (8,3) child position + 2
    s = c[:, None, :] + d[None, :, :]                  # (8,27,3) in 1..4
    lut_parent = ((s // 2) * np.array([9, 3, 1])).sum(-1)   # index into parent's 27-neigh
    lut_child = ((s % 2) * np.array([4, 2, 1])).sum(-1)     # child slot in that neighbour
    return d, lut_parent.astype(np.int64), lut_child.astype(np.int64)


NEIGH_OFFSETS, LUT_PARENT, LUT_CHILD = _make_luts()
LUT_KERNEL = {
    '222': [13, 14, 16, 17, 22, 23, 25, 26],
    '311': [4, 13, 22], '131': [10, 13, 16], '113': [12, 13, 14],
    '331': [1, 4, 7, 10, 13, 16, 19, 22, 25],
    '313': [3, 4, 5, 12, 13, 14, 21, 22, 23],
    '133': [9, 10, 11, 12, 13, 14, 15, 16, 17],
    '333': list(range(27)),
}


# ----------------------------------------------------------------------------
# Octree container
# ----------------------------------------------------------------------------
class RefOctree:
    """Mirror of the ``ocnn.octree.Octree`` state the reference model reads."""

    def __init__(self, depth: int, full_depth: int = 2, batch_size: int = 1):
        self.depth = depth
        self.full_depth = full_depth
        self.batch_size = batch_size
        n = depth + 1
        self.keys = [None] * n
        self.children = [None] * n
        self.neighs = [None] * n
        self.points = [None] * n          # only index `depth` is filled
        self.nnum = np.zeros(n, dtype=np.int64)
        self.nnum_nempty = np.zeros(n, dtype=np.int64)
        self.batch_nnum = None
        self.batch_nnum_nempty = None

    # -- ocnn.octree.Octree.build_octree -------------------
    def build_octree(self, points: np.ndarray) -> np.ndarray:
        D, F = self.depth, self.full_depth
        pts = (points.astype(np.float32) + np.float32(1.0)) * np.float32(2 ** (D - 1))
        ijk = pts.astype(np.int64)                      # truncation toward zero
        key = xyz2key(ijk[:, 0], ijk[:, 1], ijk[:, 2], depth=D)
        node_key, idx, counts = np.unique(key, return_inverse=True, return_counts=True)
        for d in range(F + 1):
            n = 8 ** d
            self.nnum[d] = self.nnum_nempty[d] = n
            self.keys[d] = np.arange(n, dtype=np.int64)
            self.children[d] = np.arange(n, dtype=np.int32)
        nk = node_key
        for d in range(D, F, -1):
            pkey = nk >> 3
            first = np.ones(len(pkey), dtype=bool)
            first[1:] = pkey[1:] != pkey[:-1]
            upkey = pkey[first]
            pidx = np.cumsum(first) - 1
            self.keys[d] = ((upkey[:, None] << 3) + np.arange(8, dtype=np.int64)).reshape(-1)
            self.nnum[d] = 8 * len(upkey)
            self.nnum_nempty[d] = len(nk)
            child = np.full(self.nnum[d], -1, dtype=np.int32)
            child[(pidx << 3) | (nk & 7)] = np.arange(len(nk), dtype=np.int32)
            self.children[d] = child
            nk = upkey
        child = np.full(8 ** F, -1, dtype=np.int32)
        child[nk] = np.arange(len(nk), dtype=np.int32)
        self.children[F] = child
        self.nnum_nempty[F] = len(nk)
        # leaf average: sequential fp32 accumulation in original point order
        acc = np.zeros((len(node_key), 3), dtype=np.float32)
        np.add.at(acc, idx, pts)
        self.points[D] = acc / counts[:, None].astype(np.float32)
        self.batch_nnum = self.nnum.copy()[:, None]
        self.batch_nnum_nempty = self.nnum_nempty.copy()[:, None]
        return idx

    # -- helpers mirroring ocnn.octree.Octree ---------------------------------
    def nempty_mask(self, d):
        return self.children[d] >= 0

    def key(self, d, nempty=False):
        k = self.keys[d]
        return k[self.nempty_mask(d)] if nempty else k

    def xyzb(self, d, nempty=False):
        return key2xyz(self.key(d, nempty), d)

    def batch_id(self, d, nempty=False):
        return self.key(d, nempty) >> BATCH_SHIFT

    # -- ocnn.octree.Octree.construct_neigh -----------------------------------
    def construct_neigh(self, d: int):
        B = self.batch_size
        if d <= self.full_depth:
            n = 8 ** d
            x, y, z, _ = key2xyz(np.arange(n, dtype=np.int64), d)
            xyz = np.stack([x, y, z], 1)[:, None, :] + NEIGH_OFFSETS[None]     # (n,27,3)
            xyz = xyz.reshape(-1, 3)
            neigh = xyz2key(xyz[:, 0], xyz[:, 1], xyz[:, 2], depth=d)
            neigh = neigh[None, :] + (np.arange(B, dtype=np.int64) * n)[:, None]
            bad = np.any((xyz < 0) | (xyz >= 2 ** d), axis=1)
            neigh[:, bad] = -1
            self.neighs[d] = neigh.reshape(B * n, 27)
        else:
            child_p = self.children[d - 1]
            P = self.neighs[d - 1][child_p >= 0]                  # (n_{d-1},27)
            Pn = P[:, LUT_PARENT]                                 # (n_{d-1},8,27)
            C = child_p[Pn].astype(np.int64)                      # -1 wraps; masked below
            neigh = C * 8 + LUT_CHILD[None]
            neigh[(C < 0) | (Pn < 0)] = -1
            self.neighs[d] = neigh.reshape(-1, 27)

    def construct_all_neigh(self):
        for d in range(1, self.depth + 1):
            self.construct_neigh(d)

    # -- ocnn.octree.Octree.get_neigh -----------------------------------------
    def get_neigh(self, d: int, kernel: str = '333', stride: int = 1, nempty: bool = False):
        if stride == 1:
            neigh = self.neighs[d]
        elif stride == 2:
            neigh = self.neighs[d][::8].copy()
        else:
            raise ValueError('stride must be 1 or 2')
        if nempty:
            child = self.children[d]
            if stride == 1:
                neigh = neigh[child >= 0]
            valid = neigh >= 0
            out = np.full_like(neigh, -1)
            out[valid] = child[neigh[valid]]
            neigh = out
        if kernel != '333':
            neigh = neigh[:, LUT_KERNEL[kernel]]
        return neigh

    def input_feature_P(self):
        """ocnn.modules.InputFeature('P', nempty=True): models/hotformerloc.py:28-31."""
        D = self.depth
        return self.points[D] * np.float32(2 ** (1 - D)) - np.float32(1.0)


def build_octree(points: np.ndarray, depth: int, full_depth: int = 2) -> RefOctree:
    o = RefOctree(depth, full_depth)
    o.build_octree(points)
    return o


def merge_octrees(octrees) -> RefOctree:
    """ocnn.octree.merge_octrees."""
    o0 = octrees[0]
    D, B = o0.depth, len(octrees)
    out = RefOctree(D, o0.full_depth, batch_size=B)
    out.batch_nnum = np.stack([o.nnum for o in octrees], 1)
    out.batch_nnum_nempty = np.stack([o.nnum_nempty for o in octrees], 1)
    out.nnum = out.batch_nnum.sum(1)
    out.nnum_nempty = out.batch_nnum_nempty.sum(1)
    cum = np.cumsum(out.batch_nnum_nempty, 1) - out.batch_nnum_nempty    # exclusive
    for d in range(D + 1):
        keys, children = [], []
        for i, o in enumerate(octrees):
            keys.append((o.keys[d] & KEY_MASK) | (np.int64(i) << BATCH_SHIFT))
            c = o.children[d].copy()
            c[c >= 0] += np.int32(cum[d, i])
            children.append(c)
        out.keys[d] = np.concatenate(keys)
        out.children[d] = np.concatenate(children)
    out.points[D] = np.concatenate([o.points[D] for o in octrees], 0)
    return out


def build_batch(clouds, depth: int, full_depth: int = 2, neigh: bool = True) -> RefOctree:
    """The reference's per-batch host path: eval/pnv_evaluate.py:155-185
    (per-submap build, merge, construct_all_neigh)."""
    o = merge_octrees([build_octree(c, depth, full_depth) for c in clouds])
    if neigh:
        o.construct_all_neigh()
    return o
