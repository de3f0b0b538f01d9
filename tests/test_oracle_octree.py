"""Pins oracle/octree_ref.py against the reference's own known-answer vectors
(libs/dwconv/test/data/octree/test_00{1..5}.npz, batch_45.npz; re-packed by
oracle/make_golden.py).  Bit-exact."""
import os

import numpy as np
import pytest

from oracle import octree_ref as R


@pytest.fixture(scope='module')
def fx(golden_dir):
    return np.load(os.path.join(golden_dir, 'octree_fixtures.npz'))


@pytest.mark.parametrize('i', [1, 2, 3, 4, 5])
def test_single_octree_matches_reference_fixture(fx, i):
    o = R.build_octree(fx[f't{i}_points'], int(fx[f't{i}_depth']), int(fx[f't{i}_full_depth']))
    assert np.array_equal(np.concatenate(o.keys), fx[f't{i}_key'])
    assert np.array_equal(np.concatenate(o.children), fx[f't{i}_child'])
    assert np.array_equal(o.nnum, fx[f't{i}_nnum'])
    assert np.array_equal(o.nnum_nempty, fx[f't{i}_nnum_nempty'])


def test_merge_and_neigh_match_reference_fixture(fx):
    octs = [R.build_octree(fx[f't{i}_points'], int(fx[f't{i}_depth']), int(fx[f't{i}_full_depth']))
            for i in (4, 5)]
    m = R.merge_octrees(octs)
    m.construct_all_neigh()
    assert np.array_equal(np.concatenate(m.keys), fx['b45_key'])
    assert np.array_equal(np.concatenate(m.children), fx['b45_child'])
    assert np.array_equal(m.nnum, fx['b45_nnum'])
    assert np.array_equal(m.nnum_nempty, fx['b45_nnum_nempty'])
    assert np.array_equal(np.concatenate(m.neighs[1:]), fx['b45_neigh'])


def test_key_roundtrip_and_wrap():
    rng = np.random.default_rng(0)
    xyz = rng.integers(0, 512, size=(1000, 3))
    b = rng.integers(0, 300, size=1000)
    key = R.xyz2key(xyz[:, 0], xyz[:, 1], xyz[:, 2], b, depth=9)
    x, y, z, bb = R.key2xyz(key, 9)
    assert np.array_equal(np.stack([x, y, z], 1), xyz) and np.array_equal(bb, b)
    # coordinate 2^depth wraps to 0: only the low `depth` bits are used
    assert R.xyz2key([512], [0], [0], depth=9)[0] == 0


def test_get_neigh_222_is_children_table():
    """'222'/stride-2/nempty table == children[d].view(-1,8) (used by the CUDA path)."""
    rng = np.random.default_rng(1)
    clouds = [rng.uniform(-1, 1, size=(3000, 3)).astype(np.float32) for _ in range(3)]
    o = R.build_batch(clouds, depth=6)
    for d in range(3, 7):
        assert np.array_equal(o.get_neigh(d, '222', 2, True), o.children[d].reshape(-1, 8))


def test_edge_clouds():
    # +-1.0 inclusive, duplicates, single point, all points in one leaf
    pts = np.array([[1.0, 1.0, 1.0], [-1.0, -1.0, -1.0], [-0.0, 0.0, 0.0], [1.0, -1.0, 0.5]],
                   dtype=np.float32)
    o = R.build_octree(pts, 5)
    assert o.nnum_nempty[5] == 3          # (1,1,1) wraps onto (-1,-1,-1)'s cell
    o = R.build_octree(np.zeros((1, 3), np.float32), 9)
    assert list(o.nnum_nempty[3:]) == [1] * 7
    o = R.build_octree(np.full((100, 3), 0.123, np.float32), 7)
    assert o.nnum_nempty[7] == 1 and np.allclose(o.points[7][0], (0.123 + 1) * 64)
