"""GPU parity: libhfl_b200.so batched octree build vs the pinned oracle
(oracle/octree_ref.py) and vs the reference's own known-answer fixtures.
Bit-exact for every integer table; leaf means exact (same summation order)."""
import os

import numpy as np
import pytest
import torch

from oracle import octree_ref as R
from oracle import model_ref as M

pytestmark = pytest.mark.gpu


def _native(clouds, depth, full_depth=2):
    from hotformerloc_b200.octree import build_batch
    return build_batch(clouds, depth, full_depth, 'cuda', want_point_leaf=True).finalize()


def _check_against_oracle(clouds, depth, full_depth=2, check_full_neigh=True):
    o = _native(clouds, depth, full_depth)
    r = R.build_batch(clouds, depth, full_depth)
    assert np.array_equal(o.nnum.numpy(), r.nnum)
    assert np.array_equal(o.nnum_nempty.numpy(), r.nnum_nempty)
    assert np.array_equal(o.batch_nnum_nempty.numpy(), r.batch_nnum_nempty)
    assert np.array_equal(o.batch_nnum.numpy(), r.batch_nnum)
    keys, children = o.keys, o.children
    for d in range(depth + 1):
        assert np.array_equal(keys[d].cpu().numpy(), r.keys[d]), f'keys[{d}]'
        assert np.array_equal(children[d].cpu().numpy(), r.children[d]), f'children[{d}]'
    assert np.array_equal(o.points[depth].cpu().numpy(), r.points[depth])
    for d in range(full_depth, depth + 1):
        assert np.array_equal(o.ne_table(d).cpu().numpy(), r.get_neigh(d, '333', 1, True)), f'ne[{d}]'
        assert np.array_equal(o.child_table(d).cpu().numpy().reshape(-1, 8),
                              r.get_neigh(d, '222', 2, True)) or d <= full_depth
    if check_full_neigh:
        neighs = o.neighs
        for d in range(1, depth + 1):
            assert np.array_equal(neighs[d].cpu().numpy(), r.neighs[d]), f'neighs[{d}]'
        assert np.array_equal(o.get_neigh(depth, '222', 2, True).cpu().numpy(),
                              r.get_neigh(depth, '222', 2, True))
        assert np.array_equal(o.get_neigh(depth - 1, '333', 1, False).cpu().numpy(),
                              r.get_neigh(depth - 1, '333', 1, False))
    # per-token tables used by the attention kernels
    for d in range(max(full_depth, depth - 5), depth + 1):
        n = int(r.nnum_nempty[d])
        n_pad = -(-n // 192) * 192
        t = o.tokens(d, n_pad).cpu().numpy()
        x, y, z, b = R.key2xyz(r.key(d, True), d)
        assert np.array_equal(t[:n], np.stack([x, y, z, b], 1).astype(np.int16))
        assert (t[n:, :3] == 0).all() and (t[n:, 3] == len(clouds)).all()
    return o, r


def test_reference_fixtures(golden_dir):
    fx = np.load(os.path.join(golden_dir, 'octree_fixtures.npz'))
    for i in range(1, 6):
        o = _native([fx[f't{i}_points']], int(fx[f't{i}_depth']), int(fx[f't{i}_full_depth']))
        assert np.array_equal(torch.cat(o.keys).cpu().numpy(), fx[f't{i}_key'])
        assert np.array_equal(torch.cat(o.children).cpu().numpy(), fx[f't{i}_child'])
        assert np.array_equal(o.nnum.numpy(), fx[f't{i}_nnum'])
        assert np.array_equal(o.nnum_nempty.numpy(), fx[f't{i}_nnum_nempty'])
    o = _native([fx['t4_points'], fx['t5_points']], 6, 3)
    assert np.array_equal(torch.cat(o.keys).cpu().numpy(), fx['b45_key'])
    assert np.array_equal(torch.cat(o.children).cpu().numpy(), fx['b45_child'])
    assert np.array_equal(torch.cat(o.neighs[1:]).cpu().numpy(), fx['b45_neigh'])


@pytest.mark.parametrize('depth,n,B', [(9, 4096, 1), (9, 4096, 5), (7, 30000, 3), (6, 777, 4),
                                       (9, 65536, 2), (9, 262144, 1)])   # configs[4] point-count sweep sizes
def test_lidar_clouds(depth, n, B):
    g = torch.Generator().manual_seed(11)
    _check_against_oracle([M.lidar_cloud(n, g) for _ in range(B)], depth)


def test_ragged_and_edge_clouds():
    rng = np.random.default_rng(3)
    clouds = [
        np.array([[1.0, 1.0, 1.0], [-1.0, -1.0, -1.0], [-0.0, 0.0, 0.0], [1.0, -1.0, 0.5]], np.float32),
        np.zeros((1, 3), np.float32),                                  # P = 1
        np.full((100, 3), 0.123, np.float32),                          # one leaf
        rng.uniform(-1, 1, size=(5000, 3)).astype(np.float32),         # uniform cube
        np.repeat(rng.uniform(-1, 1, size=(50, 3)).astype(np.float32), 7, 0),   # duplicates
    ]
    _check_against_oracle(clouds, 9)
    _check_against_oracle(clouds[::-1], 5)
    _check_against_oracle(clouds[3:4], 3, full_depth=2)


def test_point_leaf_index():
    g = torch.Generator().manual_seed(5)
    c = M.lidar_cloud(3000, g)
    from hotformerloc_b200.octree import Octree, Points
    o = Octree(7, 2)
    idx = o.build_octree(Points(torch.from_numpy(c))).cpu().numpy()
    r = R.RefOctree(7, 2)
    assert np.array_equal(idx, r.build_octree(c))


def test_many_submaps_and_sort_property():
    """size-independent property at BASELINE scale: 256 x 4096 points, depth 9."""
    g = torch.Generator().manual_seed(21)
    clouds = [M.lidar_cloud(4096, g) for _ in range(256)]
    o = _native(clouds, 9)
    for d in range(2, 10):
        k = o.key(d, nempty=True)
        assert bool((k[1:] > k[:-1]).all()), 'node keys strictly increasing'
        assert int(o.nnum[d]) == (8 * int(o.nnum_nempty[d - 1]) if d > 2 else 256 * 64)
        c = o.children[d]
        nz = c[c >= 0]
        assert torch.equal(nz, torch.arange(nz.numel(), device=nz.device, dtype=nz.dtype))
    # per-submap results are independent of the rest of the batch
    r = R.build_octree(clouds[200], 9)
    assert np.array_equal(o.batch_nnum_nempty[:, 200].numpy(), r.nnum_nempty)


def test_errors_are_reported_not_thrown_across_abi():
    from hotformerloc_b200 import native as N
    from hotformerloc_b200.octree import build_batch
    with pytest.raises(ValueError):
        build_batch([np.zeros((0, 3), np.float32)], 7)
    with pytest.raises(AssertionError):
        build_batch([np.zeros((4, 3), np.float32)], 3, full_depth=3)


def test_reference_call_flow_build_then_merge():
    """The reference's eval loop builds one octree per submap and merges them
    (eval/pnv_evaluate.py:173-175, :122-126): Octree(depth, full_depth=2).build_octree(Points(x)) ...
    merge_octrees(list).  The merged octree must equal the batched build bit for bit, also when
    more octrees are built before an earlier one is finalised than the pinned read-back pool has slots."""
    from hotformerloc_b200.octree import Octree, Points, build_batch, merge_octrees
    g = torch.Generator().manual_seed(31)
    clouds = [M.lidar_cloud(n, g) for n in (4096, 300, 4096, 1, 2500, 4096, 77, 4096, 1500, 4096, 900, 64)]
    singles = []
    for c in clouds:
        o = Octree(7, full_depth=2)
        idx = o.build_octree(Points(torch.from_numpy(c)))
        assert idx.shape == (len(c),)
        singles.append(o)
    # 12 builds in flight > 8 pinned slots: every per-submap octree must still report ITS OWN counts
    r1 = [R.build_batch([c], 7) for c in clouds]
    for o, r in zip(singles, r1):
        assert np.array_equal(o.finalize().nnum_nempty.numpy(), r.nnum_nempty)
        assert np.array_equal(o.nnum.numpy(), r.nnum)
    merged = merge_octrees(singles)
    ref = R.build_batch(clouds, 7)
    direct = build_batch(clouds, 7, 2, 'cuda')
    assert merged.batch_size == len(clouds)
    for d in range(8):
        assert np.array_equal(merged.keys[d].cpu().numpy(), ref.keys[d]), d
        assert np.array_equal(merged.children[d].cpu().numpy(), ref.children[d]), d
        assert torch.equal(merged.keys[d], direct.keys[d])
    assert np.array_equal(merged.batch_nnum_nempty.numpy(), ref.batch_nnum_nempty)
    # merging already-batched octrees keeps the submap order
    two = merge_octrees([build_batch(clouds[:5], 7, 2, 'cuda'), build_batch(clouds[5:], 7, 2, 'cuda')])
    for d in range(8):
        assert torch.equal(two.keys[d], direct.keys[d])
        assert torch.equal(two.children[d], direct.children[d])
