"""CPU tests of the host-side logic: C-ABI surface, config schema, state_dict layout,
loaders, evaluation bookkeeping, relay-token tables, and the multi-process sharding
(gloo, world_size 2)."""
import json
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

from oracle import model_ref as M
from oracle import octree_ref as R
from oracle.make_golden import CASES, recall_case
from tests.common import GOLDEN, ROOT, case_clouds


def test_library_exports_every_declared_symbol():
    from hotformerloc_b200 import native
    path = native.build()
    header = open(os.path.join(ROOT, 'include', 'hfl.h')).read()
    declared = set(re.findall(r'\b(hfl_[a-z0-9_]+)\s*\(', header))
    assert declared, 'no declarations found'
    import ctypes
    lib = ctypes.CDLL(path)
    for name in declared:
        assert hasattr(lib, name), f'{name} declared in include/hfl.h but not exported'
    assert declared == set(native.SIGNATURES), declared ^ set(native.SIGNATURES)
    assert lib.hfl_version() >= 100
    # the product path has no CPU fallback: building an octree without CUDA must fail loudly
    if not torch.cuda.is_available():
        from hotformerloc_b200.octree import build_batch
        with pytest.raises(Exception):
            build_batch([np.zeros((10, 3), np.float32)], 7, 2, 'cpu')


@pytest.mark.parametrize('cfg', ['oxford', 'cs-wild-places', 'wild-places', 'cs-campus3d'])
def test_config_schema_and_state_dict_layout(cfg, tmp_path):
    from hotformerloc_b200.config.presets import write_configs
    from hotformerloc_b200.misc.utils import TrainingParams
    from hotformerloc_b200.models.model_factory import model_factory
    paths = write_configs(str(tmp_path), cfg, dataset_folder=str(tmp_path))
    params = TrainingParams(paths['config'], paths['model_config'])
    assert params.load_octree and params.model_params.num_pyramid_levels == 3
    model = model_factory(params.model_params)
    ref = json.load(open(os.path.join(GOLDEN, f'state_shapes_{cfg}.json')))
    mine = {k: list(v.shape) for k, v in model.state_dict().items()}
    assert mine == ref                       # names AND shapes of the reference checkpoint layout
    model.load_state_dict(M.synthetic_state_dict(ref))


def test_loaders(tmp_path):
    from hotformerloc_b200.datasets.pointnetvlad.pnv_raw import PNVPointCloudLoader
    from hotformerloc_b200.datasets.CSWildPlaces.CSWildPlaces_raw import CSWildPlacesPointCloudLoader
    pts = np.random.default_rng(0).uniform(-1, 1, (100, 3))
    p = tmp_path / 'a.bin'
    pts.astype(np.float64).tofile(p)
    assert np.array_equal(PNVPointCloudLoader()(str(p)), pts.astype(np.float32))
    hdr = ('# .PCD v0.7\nVERSION 0.7\nFIELDS x y z intensity\nSIZE 4 4 4 4\nTYPE F F F F\n'
           'COUNT 1 1 1 1\nWIDTH 100\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS 100\n')
    rec = np.concatenate([pts.astype(np.float32), np.ones((100, 1), np.float32)], 1)
    with open(tmp_path / 'b.pcd', 'wb') as f:
        f.write((hdr + 'DATA binary\n').encode())
        f.write(rec.tobytes())
    with open(tmp_path / 'c.pcd', 'w') as f:
        f.write(hdr + 'DATA ascii\n')
        for r in rec:
            f.write(' '.join(repr(float(v)) for v in r) + '\n')
    xyz_hdr = ('VERSION 0.7\nFIELDS x y z\nSIZE 4 4 4\nTYPE F F F\nCOUNT 1 1 1\nWIDTH 100\nHEIGHT 1\n'
               'POINTS 100\nDATA binary\n')
    bad = pts.astype(np.float32).copy()
    bad[7, 1] = np.nan                                       # non-finite points are dropped (open3d behaviour)
    with open(tmp_path / 'd.pcd', 'wb') as f:                # xyz-only float32 records: the reshape fast path
        f.write(xyz_hdr.encode())
        f.write(bad.tobytes())
    for name in ('b.pcd', 'c.pcd'):
        assert np.array_equal(CSWildPlacesPointCloudLoader()(str(tmp_path / name)),
                              pts.astype(np.float32))
    assert np.array_equal(CSWildPlacesPointCloudLoader()(str(tmp_path / 'd.pcd')),
                          np.delete(pts.astype(np.float32), 7, axis=0))


def test_normalize_matches_reference_golden():
    """Normalize (datasets/augmentation.py:185-235) in every mode the eval configs can select:
    bit-exact against outputs of the reference's own class (oracle/make_golden.py:make_normalize_golden)."""
    from hotformerloc_b200.datasets.coordinate_utils import Normalize
    g = np.load(os.path.join(GOLDEN, 'normalize.npz'))
    variants = {'bbox': {}, 'scale30': dict(scale_factor=30.0), 'sphere': dict(unit_sphere_norm=True),
                'sphere_scale40': dict(unit_sphere_norm=True, scale_factor=40.0),
                'range2': dict(norm_range=2.0), 'bbox_nocenter': dict(zero_mean=False)}
    for name, kw in variants.items():
        out = Normalize(**kw)(torch.from_numpy(g['cloud']).clone()).numpy()
        assert np.array_equal(out, g[name]), name


def test_cylindrical_matches_reference_golden():
    from hotformerloc_b200.datasets.coordinate_utils import cylindrical_for_octree
    g = np.load(os.path.join(GOLDEN, 'cylindrical.npz'))
    assert np.array_equal(cylindrical_for_octree(g['cloud']), g['out'])


def test_recall_bookkeeping_matches_reference_get_recall():
    """tests/golden/recall.npz was produced by the reference's own get_recall."""
    from hotformerloc_b200.eval.pnv_evaluate import recall_from_neighbors
    sets, vecs = recall_case()
    gold = np.load(os.path.join(GOLDEN, 'recall.npz'))
    for m in range(3):
        for n in range(3):
            if m == n:
                continue
            d = ((vecs[n][:, None, :].astype(np.float64) - vecs[m][None]) ** 2).sum(-1)
            idx = np.argsort(d, axis=1, kind='stable')[:, :25]
            rec, opr, mrr = recall_from_neighbors(idx, sets[n], m, len(vecs[m]))
            assert np.allclose(rec, gold[f'recall_{m}_{n}'])
            assert np.isclose(opr, gold[f'opr_{m}_{n}']) and np.isclose(mrr, gold[f'mrr_{m}_{n}'])


@pytest.mark.parametrize('name', ['cswp_b6_stress', 'oxford_b4_init'])
def test_relay_token_tables_match_reference_octree_t(name):
    """The engine's host tables vs the reference's OctreeT.build_t (tests/golden/octree_t.npz),
    including submaps that own zero relay tokens at the coarsest level."""
    from hotformerloc_b200.models.hotformerloc import _Engine
    cfg, depth, spec, seed, mode = CASES[name]
    gold = np.load(os.path.join(GOLDEN, 'octree_t.npz'))
    K = 64 if cfg == 'cs-wild-places' else 48
    ref = R.build_batch(case_clouds(name), depth, neigh=False)
    B, d0 = ref.batch_size, depth - 2
    depths = [d0 - 1 - j for j in range(3)]

    class FakeOct:
        device = 'cpu'
        batch_nnum_nempty = torch.from_numpy(ref.batch_nnum_nempty)
    nl = [int(ref.nnum_nempty[d]) for d in depths]
    npad = [-(-v // (4 * K)) * 4 * K for v in nl]
    nwin = [v // K for v in npad]
    rows = [v * (K + 1) for v in nwin]
    Roff = np.concatenate([[0], np.cumsum(rows)])
    from hotformerloc_b200 import native
    class _Pool:                                     # CPU stand-in for the pinned staging pool
        def take(self, n): return torch.empty(n, dtype=torch.uint8)
        def mark(self): pass
    saved, native.pinned = native.pinned, _Pool()
    try:
        t = _Engine.__new__(_Engine)._host_tables(FakeOct, depths, nl, npad, nwin, Roff, K, B)
    finally:
        native.pinned = saved
    for j, d in enumerate(depths):
        assert npad[j] == int(gold[f'{name}_nnum_a_{d}'])
        assert np.array_equal(t['num_windows'][j], gold[f'{name}_num_windows_{d}'])
    tot = np.sum(t['num_windows'], 0)
    assert np.array_equal(tot, gold[f'{name}_rt_combined'])
    # attention-allowed pattern of the padded (B,N,N) reference mask vs the ragged ids
    shp = tuple(gold[f'{name}_rt_attn_shape'])
    allowed = np.unpackbits(gold[f'{name}_rt_attn_allowed'])[:np.prod(shp)].reshape(shp).astype(bool)
    cu, ids = t['cu'].numpy(), t['ids'].numpy()
    for b in range(B):
        i = ids[cu[b]:cu[b + 1]]
        assert np.array_equal(i[:, None] == i[None, :], allowed[b, :len(i), :len(i)])
    # rt rows: level-major window order inside each submap, hat-layout row of each relay token
    rt_rows = t['rt_rows'].numpy()
    for b in range(B):
        seg = rt_rows[cu[b]:cu[b + 1]]
        o = 0
        for j in range(3):
            nw = int(t['num_windows'][j][b])
            start = int(np.cumsum(t['num_windows'][j])[b] - nw)
            assert np.array_equal(seg[o:o + nw], Roff[j] + (start + np.arange(nw)) * (K + 1))
            o += nw


def test_sharding_two_ranks_gloo(tmp_path):
    """world_size-2 gloo run of the batch-granular sharding + descriptor gather."""
    script = tmp_path / 'w.py'
    script.write_text('''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, %r)
from hotformerloc_b200.eval.pnv_evaluate import shard_batches, gather_rows
dist.init_process_group('gloo')
rank, world = dist.get_rank(), dist.get_world_size()
n, bs = 1037, 128
spans = shard_batches(n, bs, rank, world)
assert [t for t, _, _ in spans] == list(range(rank, (n + bs - 1) // bs, world))
local = torch.cat([torch.arange(b, e, dtype=torch.float32)[:, None].repeat(1, 4) for _, b, e in spans])
full = gather_rows(local, spans, n, rank, world, bs)
assert torch.equal(full, torch.arange(n, dtype=torch.float32)[:, None].repeat(1, 4)), rank
dist.destroy_process_group()
print('ok', rank)
''' % ROOT)
    env = dict(os.environ, MASTER_ADDR='127.0.0.1', MASTER_PORT='29571')
    r = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1',
                        '--nproc-per-node', '2', '--master-addr', '127.0.0.1', '--master-port',
                        '29571', str(script)], env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count('ok') == 2


def test_evaluate_splits_bookkeeping(tmp_path, monkeypatch):
    """eval/pnv_evaluate_splits.py mirror: per-split stats keyed by the split directory of the query
    run, 'average' only for more than one pair, the doubly nested average over locations, the
    report files (reference pnv_evaluate_splits.py:81-133, :335-371).  The device stages are
    replaced by given descriptors and a brute-force numpy search; the recall numbers per pair
    are the reference's own get_recall golden."""
    import pickle
    from types import SimpleNamespace
    from hotformerloc_b200.eval import pnv_evaluate as E
    from hotformerloc_b200.eval import pnv_evaluate_splits as S
    sets, vecs = recall_case()
    gold = np.load(os.path.join(GOLDEN, 'recall.npz'))
    sets = [dict(s) for s in sets]
    for r, s in enumerate(sets):                       # run r lives under split 'split<r % 2>'
        for k in s:
            s[k] = dict(s[k], query=f'split{r % 2}/run{r}/{k}.bin')
    by_path = {s[0]['query']: r for r, s in enumerate(sets)}
    monkeypatch.setattr(S, 'get_latent_vectors', lambda model, data_set, device, params: vecs[by_path[data_set[0]['query']]])

    def fake_get_recall(m, n, dbv, qv, query_sets, database_sets, log=False, model_name='model'):
        d = ((qv[n][:, None, :].astype(np.float64) - dbv[m][None]) ** 2).sum(-1)
        idx = np.argsort(d, axis=1, kind='stable')[:, :25]
        return E.recall_from_neighbors(idx, query_sets[n], m, len(dbv[m]))
    monkeypatch.setattr(S, 'get_recall', fake_get_recall)
    model = SimpleNamespace(eval=lambda: None)
    params = SimpleNamespace(dataset_name='Oxford', skip_same_run=True, dataset_folder=str(tmp_path))
    st = S.evaluate_dataset(model, 'cpu', params, sets, sets)
    # pairs (i, j), i != j, are stored under the split of query run j; later pairs overwrite earlier
    assert set(st) == {'split0', 'split1', 'average'}
    assert np.allclose(st['split1']['ave_recall'], gold['recall_2_1'])       # last pair with j == 1
    assert np.allclose(st['split0']['ave_recall'], gold['recall_2_0'])       # last pair with j in {0, 2}
    mean_all = np.mean([gold[f'recall_{m}_{n}'] for m in range(3) for n in range(3) if m != n], axis=0)
    assert np.allclose(st['average']['ave_recall'], mean_all)
    one = S.evaluate_dataset(model, 'cpu', params, sets[:2], sets[:2])
    assert 'average' in one and len(one) == 3
    single = S.evaluate_dataset(model, 'cpu', SimpleNamespace(dataset_name='Oxford', skip_same_run=False,
                                                              dataset_folder='.'), sets[:1], sets[:1])
    assert set(single) == {'split0'}                                         # one pair: no 'average'
    for name in ('oxford', 'university', 'residential', 'business'):
        pickle.dump(sets, open(os.path.join(tmp_path, f'{name}_evaluation_database.pickle'), 'wb'))
        pickle.dump(sets, open(os.path.join(tmp_path, f'{name}_evaluation_query.pickle'), 'wb'))
    stats = S.evaluate(model, 'cpu', params)
    assert set(stats) == {'oxford', 'university', 'residential', 'business', 'average'}
    assert np.allclose(stats['average']['average']['ave_recall'], mean_all)
    out = os.path.join(tmp_path, 'res.txt')
    S.pnv_write_eval_stats(out, 'prefix', stats)
    txt = open(out).read()
    assert 'Split: [split0]' in txt and '[average]' in txt and 'AR@1%' in txt
    S.print_eval_stats(stats)


def test_config4_dataset_roundtrip(tmp_path):
    """tools/config4_eval.py writes the reference's on-disk evaluation format (binary .pcd submaps +
    per-run dicts with 'query' / 'northing' / 'easting' / true-neighbour lists); the
    native PCD reader and prepare_cloud (Normalize -> range mask -> cylindrical) consume it."""
    import importlib.util
    from hotformerloc_b200.config.presets import write_configs
    from hotformerloc_b200.datasets.CSWildPlaces.CSWildPlaces_raw import CSWildPlacesPointCloudLoader
    from hotformerloc_b200.datasets.coordinate_utils import CylindricalCoordinates, Normalize
    from hotformerloc_b200.eval import pnv_evaluate as E
    from hotformerloc_b200.misc.utils import TrainingParams
    spec = importlib.util.spec_from_file_location('config4_eval', os.path.join(ROOT, 'tools', 'config4_eval.py'))
    c4 = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(c4)
    sets = c4.make_dataset(str(tmp_path), runs=2, per_run=3, points=2000, seed=3)
    assert len(sets) == 2 and sorted(sets[0]) == [0, 1, 2]
    assert sets[0][1][1] == [1] and sets[0][1][0] == []           # true neighbour in the OTHER run only
    loader = CSWildPlacesPointCloudLoader()
    pts = loader(os.path.join(tmp_path, sets[1][2]['query']))
    assert pts.dtype == np.float32 and pts.shape[1] == 3 and 1500 < len(pts) <= 2000
    assert np.isfinite(pts).all() and np.abs(pts).max() <= 31.0   # metres
    paths = write_configs(os.path.join(tmp_path, 'cfg'), 'wild-places', dataset_folder=str(tmp_path))
    params = TrainingParams(paths['config'], paths['model_config'])
    out = E.prepare_cloud(pts, params, Normalize(scale_factor=params.scale_factor,
                                                 unit_sphere_norm=params.unit_sphere_norm),
                          CylindricalCoordinates(use_octree=True))
    assert out.dtype == np.float32 and out.shape[1] == 3 and len(out) > 1000
    assert np.abs(out).max() <= 1.0
    # batches of val_batch_size in dataset order, round-robin over the ranks
    spans = [E.shard_batches(300, 128, r, 2) for r in range(2)]
    assert spans[0] == [(0, 0, 128), (2, 256, 300)] and spans[1] == [(1, 128, 256)]


def test_prefetching_loader_keeps_order_and_batch_composition(tmp_path, monkeypatch):
    """get_latent_vectors reads / prepares the next batches on a thread pool while the current one is
    embedded; the descriptors must still come out in dataset order with the reference's batch
    composition (chunks of val_batch_size in dict order, eval/pnv_evaluate.py:151-185)."""
    from types import SimpleNamespace
    from hotformerloc_b200.eval import pnv_evaluate as E
    n, bs = 11, 3
    rng = np.random.default_rng(5)
    data_set = {}
    for i in range(n):
        pts = rng.uniform(-1, 1, (50 + i, 3))
        pts[:, 0] = np.clip(pts[:, 0] * 0.01 + i / 20.0, -1, 1)          # cloud i is identifiable by its mean x
        (tmp_path / f'{i}.bin').write_bytes(pts.astype(np.float64).tobytes())
        data_set[100 + i] = {'query': f'{i}.bin'}
    batches = []
    monkeypatch.setattr(E, 'collate_batch', lambda clouds, device, params: clouds)

    def fake_embed(model, clouds):
        batches.append([round(float(c[:, 0].mean()) * 20) for c in clouds])
        return torch.tensor([[c[:, 0].mean(), len(c), 0.0, 1.0] for c in clouds], dtype=torch.float32)
    monkeypatch.setattr(E, 'compute_embedding', fake_embed)
    params = SimpleNamespace(debug=False, dataset_name='Oxford', normalize_points=False, scale_factor=None,
                             unit_sphere_norm=False, load_octree=True, val_batch_size=bs,
                             dataset_folder=str(tmp_path),
                             model_params=SimpleNamespace(coordinates='cartesian', output_dim=4))
    model = SimpleNamespace(eval=lambda: None)
    outs = {}
    for threads in ('1', '4'):
        monkeypatch.setenv('HFL_LOADER_THREADS', threads)
        batches.clear()
        outs[threads] = E.get_latent_vectors(model, data_set, 'cpu', params)
        assert batches == [[0, 1, 2], [3, 4, 5], [6, 7, 8], [9, 10]]
    assert np.array_equal(outs['1'], outs['4'])
    assert np.allclose(outs['4'][:, 0] * 20, np.arange(n), atol=0.05)
    assert np.array_equal(outs['4'][:, 1], 50 + np.arange(n))


def test_get_latent_vectors_edge_cases(tmp_path, monkeypatch):
    """Empty run (-> None, eval/pnv_evaluate.py:151,187), a run smaller than one batch, and the reference's --debug short cut
    (eval/pnv_evaluate.py:131-133: random vectors of the right shape, no model call)."""
    from types import SimpleNamespace
    from hotformerloc_b200.eval import pnv_evaluate as E
    calls = []
    monkeypatch.setattr(E, 'collate_batch', lambda clouds, device, params: clouds)
    monkeypatch.setattr(E, 'compute_embedding',
                        lambda model, clouds: (calls.append(len(clouds)), torch.zeros(len(clouds), 4))[1])
    mk = lambda **kw: SimpleNamespace(debug=False, dataset_name='Oxford', normalize_points=False,
                                      scale_factor=None, unit_sphere_norm=False, load_octree=True,
                                      val_batch_size=8, dataset_folder=str(tmp_path),
                                      model_params=SimpleNamespace(coordinates='cartesian', output_dim=4), **kw)
    model = SimpleNamespace(eval=lambda: None)
    # an empty run: None, like the reference (its `embeddings` is never allocated); evaluate_dataset skips it
    assert E.get_latent_vectors(model, {}, 'cpu', mk()) is None and calls == []
    (tmp_path / 'a.bin').write_bytes(np.zeros((5, 3)).tobytes())
    out = E.get_latent_vectors(model, {7: {'query': 'a.bin'}}, 'cpu', mk())
    assert out.shape == (1, 4) and calls == [1]
    p = mk()
    p.debug = True
    out = E.get_latent_vectors(model, {i: {'query': 'missing.bin'} for i in range(3)}, 'cpu', p)
    assert out.shape == (3, 4) and calls == [1]
    with pytest.raises(ValueError):
        bad = mk()
        bad.dataset_name = 'NoSuchDataset'
        E.get_latent_vectors(model, {0: {'query': 'a.bin'}}, 'cpu', bad)


@pytest.mark.parametrize('cfg', ['wild-places', 'cs-wild-places', 'oxford'])
def test_prepare_batch_is_bit_identical(cfg, tmp_path):
    """The batched host prep (one vectorised pass per evaluation batch) returns exactly the arrays of the
    per-submap prepare_cloud() -- the mirror of eval/pnv_evaluate.py:158-171 -- for clouds of very different sizes,
    clouds that lose points to the range masks, and an empty cloud."""
    from hotformerloc_b200.config.presets import write_configs
    from hotformerloc_b200.eval import pnv_evaluate as E
    from hotformerloc_b200.misc.utils import TrainingParams
    paths = write_configs(str(tmp_path), cfg, dataset_folder=str(tmp_path))
    params = TrainingParams(paths['config'], paths['model_config'])
    normalize = None
    if params.normalize_points or params.scale_factor is not None:
        normalize = E.Normalize(scale_factor=params.scale_factor, unit_sphere_norm=params.unit_sphere_norm)
    cyl = E.CylindricalCoordinates(use_octree=True) \
        if params.load_octree and params.model_params.coordinates == 'cylindrical' else None
    rng = np.random.default_rng(3)
    scale = 30.0 if normalize is not None else 1.0
    raws = []
    for n in (30000, 7, 1, 4096, 60000, 513):
        pts = rng.uniform(-1, 1, (n, 3)) * scale * rng.uniform(0.5, 1.3)      # some clouds exceed the unit range
        pts[:, 2] *= 0.3
        raws.append(pts.astype(np.float64))
    got = E.prepare_batch(raws, params, normalize, cyl)
    assert len(got) == len(raws)
    for g, r in zip(got, raws):
        ref = E.prepare_cloud(r, params, normalize, cyl)
        assert g.dtype == ref.dtype and g.shape == ref.shape
        assert np.array_equal(g, ref)
    # float32 inputs (the .pcd reader's xyz fast path) too
    raws32 = [r.astype(np.float32) for r in raws[:3]]
    for g, r in zip(E.prepare_batch(raws32, params, normalize, cyl), raws32):
        ref = E.prepare_cloud(r, params, normalize, cyl)
        assert g.dtype == ref.dtype and np.array_equal(g, ref)
