"""Evaluation entry point -- same functions, arguments and result files as the
reference's ``eval/pnv_evaluate.py`` (evaluate :32, evaluate_dataset :76,
get_latent_vectors :129, get_recall :228, print_eval_stats :318,
pnv_write_eval_stats :327, CLI :351-407), with the hot loops on the GPU:

* the per-submap CPU ``build_octree`` + ``merge_octrees`` + ``construct_all_neigh`` loop
  (:155-176, :122-126) is one batched device build per ``val_batch_size`` submaps;
* the faiss / KDTree search (:200-225) is the exact fp32 L2 top-k kernel;
* with ``torch.distributed`` initialised (one process per GPU) evaluation batches are
  sharded across ranks at *batch granularity* (batch t -> rank t mod W, preserving the
  reference's batch composition and therefore its descriptors), descriptors are
  all-gathered over NCCL, and the database is sharded across ranks for the top-k search.
"""
from __future__ import annotations

import argparse
import os
import pickle
from typing import List, Optional

import numpy as np
import torch
import torch.distributed as dist

from ..datasets.coordinate_utils import CylindricalCoordinates, Normalize
from ..datasets.CSWildPlaces.CSWildPlaces_raw import CSWildPlacesPointCloudLoader
from ..datasets.pointnetvlad.pnv_raw import PNVPointCloudLoader
from ..misc.utils import TrainingParams, set_seed
from ..models.model_factory import model_factory
from ..octree import build_batch
from .utils import get_query_database_splits

NUM_NEIGHBORS = 25


def _dist_info():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def evaluate(model, device, params: TrainingParams, log: bool = False, model_name: str = 'model',
             show_progress: bool = False):
    eval_database_files, eval_query_files = get_query_database_splits(params)
    assert len(eval_database_files) == len(eval_query_files)
    stats, ave_recall, ave_opr, ave_mrr = {}, [], [], []
    for database_file, query_file in zip(eval_database_files, eval_query_files):
        pos = 1 if 'CSWildPlaces' in params.dataset_name else 0
        location_name = database_file.split('_')[pos]
        assert location_name == query_file.split('_')[pos], \
            f'Database location: {database_file} does not match query location: {query_file}'
        with open(os.path.join(params.dataset_folder, database_file), 'rb') as f:
            database_sets = pickle.load(f)
        with open(os.path.join(params.dataset_folder, query_file), 'rb') as f:
            query_sets = pickle.load(f)
        temp = evaluate_dataset(model, device, params, database_sets, query_sets, log=log,
                                model_name=model_name, show_progress=show_progress)
        stats[location_name] = temp
        ave_opr.append(temp['ave_one_percent_recall'])
        ave_recall.append(temp['ave_recall'])
        ave_mrr.append(temp['ave_mrr'])
    stats['average'] = {'ave_one_percent_recall': np.mean(ave_opr),
                        'ave_recall': np.mean(ave_recall, axis=0), 'ave_mrr': np.mean(ave_mrr)}
    return stats


def evaluate_dataset(model, device, params: TrainingParams, database_sets, query_sets,
                     log: bool = False, model_name: str = 'model', show_progress: bool = False):
    recall = np.zeros(NUM_NEIGHBORS)
    count, one_percent_recall, mrr = 0, [], []
    model.eval()
    database_embeddings = [get_latent_vectors(model, s, device, params) for s in database_sets]
    query_embeddings = [get_latent_vectors(model, s, device, params) for s in query_sets]
    for i in range(len(database_sets)):
        for j in range(len(query_sets)):
            if (i == j and params.skip_same_run) or database_embeddings[i] is None \
                    or query_embeddings[j] is None:
                continue
            if 'CSCampus3D' in params.dataset_name and i != 1:
                continue
            pair_recall, pair_opr, pair_mrr = get_recall(i, j, database_embeddings,
                                                         query_embeddings, query_sets,
                                                         database_sets, log=log,
                                                         model_name=model_name)
            recall += np.array(pair_recall)
            count += 1
            one_percent_recall.append(pair_opr)
            mrr.append(pair_mrr)
    return {'ave_one_percent_recall': np.mean(one_percent_recall), 'ave_recall': recall / count,
            'ave_mrr': np.mean(mrr)}


def prepare_cloud(data: np.ndarray, params: TrainingParams, normalize=None, cyl=None) -> np.ndarray:
    """Host-side input prep of one submap, operation for operation as
    eval/pnv_evaluate.py:158-171 (the octree exactness contract starts at its output)."""
    data = torch.tensor(data)
    if normalize is not None:
        data = normalize(data)
    data = data[torch.all(abs(data) <= 1.0, dim=1)]
    if cyl is not None:
        data = data[torch.linalg.norm(data[:, :2], dim=1) <= 1.0]
        data = cyl(data)
    return data.numpy()


def prepare_batch(raws: List[np.ndarray], params: TrainingParams, normalize=None, cyl=None) -> List[np.ndarray]:
    """prepare_cloud() for the submaps of one evaluation batch at once: the same torch operations, applied to the
    concatenated points with per-submap scalars broadcast through a segment index, so that the Python / dispatch
    cost is paid per batch instead of per submap (the per-submap form costs tens of ms of host time per cloud and
    bounds the file -> descriptor pipeline at a fraction of what the GPU embeds).  Every operation is elementwise
    or a per-row / per-submap min / max, so the values are bit-identical to the per-submap path
    (tests/test_host_cpu.py::test_prepare_batch_is_bit_identical); unit-sphere normalisation (a per-submap float
    mean / max of norms) keeps the per-submap path."""
    if not raws:
        return []
    if normalize is not None and normalize.unit_sphere_norm:
        return [prepare_cloud(r, params, normalize, cyl) for r in raws]
    B = len(raws)
    lens = torch.tensor([len(r) for r in raws], dtype=torch.int64)
    data = torch.tensor(np.concatenate(raws, axis=0))
    seg = torch.repeat_interleave(torch.arange(B), lens)
    if normalize is not None:
        lo = torch.segment_reduce(data, 'min', lengths=lens, axis=0)
        hi = torch.segment_reduce(data, 'max', lengths=lens, axis=0)
        if normalize.zero_mean:
            data = data - ((lo + hi) * 0.5)[seg]
        if normalize.scale_factor is not None:
            data = data / normalize.scale_factor
        else:
            box = (hi - lo).max(dim=1).values + 1.0e-6
            data = data * (2.0 * normalize.norm_range / box)[seg][:, None]
    keep = torch.all(abs(data) <= 1.0, dim=1)
    data, seg = data[keep], seg[keep]
    if cyl is not None:
        keep = torch.linalg.norm(data[:, :2], dim=1) <= 1.0
        data, seg = data[keep], seg[keep]
        data = cyl(data)
    counts = torch.bincount(seg, minlength=B).tolist()
    out, o = [], 0
    arr = data.numpy()
    for c in counts:
        out.append(arr[o:o + c])
        o += c
    return out


def device_prep_supported(normalize, cyl) -> bool:
    """The opt-in device-side prep (HFL_DEVICE_PREP=1, hfl_prepare_clouds) covers the bounding-box / fixed-scale
    Normalize and CylindricalCoordinates(use_octree=True) on fp32 clouds."""
    return normalize is None or not normalize.unit_sphere_norm


class _PinnedStage:
    """Reusable pinned staging buffers for the raw points of a batch (allocating + pinning tens of MB per batch costs
    more host time than the copy itself)."""

    def __init__(self):
        self.free = []

    def take(self, n_points: int) -> torch.Tensor:
        for i, t in enumerate(self.free):
            if t.shape[0] >= n_points:
                return self.free.pop(i)
        return torch.empty((max(n_points, 1) * 5 // 4, 3), dtype=torch.float32).pin_memory()

    def give(self, t: torch.Tensor) -> None:
        if len(self.free) < 4:
            self.free.append(t)


_PINNED = _PinnedStage()


def stage_raw_batch(raws: List[np.ndarray]):
    """Host half of the device-side prep (runs on a worker thread): the raw fp32 clouds of a batch packed back to
    back into a pinned staging buffer + their offsets."""
    lens = np.array([len(r) for r in raws], dtype=np.int64)
    off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    n = int(off[-1])
    host = _PINNED.take(n)
    if n:
        np.concatenate([np.asarray(r, dtype=np.float32) for r in raws], axis=0, out=host[:n].numpy())
    return host, n, off


def prepare_batch_device(raws, params: TrainingParams, normalize, cyl, device, staged=None):
    """prepare_batch() on the device: raw fp32 clouds -> one pinned H2D copy -> Normalize / masks / cylindrical /
    compaction kernels.  Returns (points, offsets, n_ticket): device tensors for build_batch_device and a pinned
    count + event the caller waits for right before the build (the point count after the masks is the one
    host-side number the octree build needs).  Values are bit-identical to prepare_cloud() except where the CPU's
    sqrt / atan2 kernels and CUDA's differ in the last bit (see csrc/prep.cu); hence opt-in."""
    from .. import ops
    host, n, off = staged if staged is not None else stage_raw_batch(raws)
    pts = host[:n].to(device, non_blocking=True)
    offd = torch.from_numpy(off).to(device, non_blocking=True)
    out, off_out, total = ops.prepare_clouds(
        pts, offd, norm=normalize is not None, zero_mean=normalize.zero_mean if normalize is not None else True,
        scale_factor=normalize.scale_factor if normalize is not None else None,
        norm_range=normalize.norm_range if normalize is not None else 1.0, cyl=cyl is not None)
    n_pin = torch.empty(1, dtype=torch.int32).pin_memory()
    n_pin.copy_(total, non_blocking=True)
    ev = torch.cuda.Event()
    ev.record()
    return out, off_out, (n_pin, ev, host)


def collate_batch(data: List[np.ndarray], device, params: TrainingParams):
    """One merged, neighbour-complete, device-resident octree for a list of prepared clouds."""
    return {'octree': build_batch(data, params.octree_depth, 2, device)}


def compute_embedding(model, batch) -> torch.Tensor:
    with torch.inference_mode():
        return model(batch)['global']


def shard_batches(n_items: int, bs: int, rank: int, world: int):
    """Batch t = items [t*bs, (t+1)*bs) goes to rank t mod world."""
    n_batches = (n_items + bs - 1) // bs
    return [(t, t * bs, min((t + 1) * bs, n_items)) for t in range(n_batches) if t % world == rank]


def gather_rows(local: torch.Tensor, spans, n_items: int, rank: int, world: int, bs: int = 0) -> torch.Tensor:
    """All-gather the rank-local descriptor rows into the (n_items, D) matrix in dataset order.
    `spans` = this rank's [(t, begin, end)] in the order `local` was filled.  One
    ``all_gather_into_tensor`` of equally sized (zero-padded to the largest shard) blocks: every rank
    can recompute every other rank's spans from (n_items, bs), so no size exchange is needed
    (payload ~ n_items * 1 KB: latency bound)."""
    D = local.shape[1]
    if world == 1:
        out = torch.empty((n_items, D), dtype=local.dtype, device=local.device)
        o = 0
        for _, b, e in spans:
            out[b:e] = local[o:o + (e - b)]
            o += e - b
        return out
    assert bs > 0, 'the batch size is needed to recompute the other ranks\' shards'
    all_spans = [shard_batches(n_items, bs, r, world) for r in range(world)]
    rows = [sum(e - b for _, b, e in sp) for sp in all_spans]
    assert rows[rank] == local.shape[0]
    cap = max(max(rows), 1)
    send = torch.zeros((cap, D), dtype=local.dtype, device=local.device)
    send[:local.shape[0]] = local
    recv = torch.empty((world, cap, D), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(recv.view(world * cap, D), send)
    out = torch.empty((n_items, D), dtype=local.dtype, device=local.device)
    for r, sp in enumerate(all_spans):
        o = 0
        for _, b, e in sp:
            out[b:e] = recv[r, o:o + (e - b)]
            o += e - b
    return out


def get_latent_vectors(model, data_set, device, params: TrainingParams):
    if params.debug:
        return np.random.rand(len(data_set), params.model_params.output_dim)
    if params.dataset_name in ['Oxford', 'CSCampus3D']:
        pc_loader = PNVPointCloudLoader()
    elif 'CSWildPlaces' in params.dataset_name or 'WildPlaces' in params.dataset_name:
        pc_loader = CSWildPlacesPointCloudLoader()
    else:
        raise ValueError('Invalid dataset_name')
    normalize = None
    if params.normalize_points or params.scale_factor is not None:
        normalize = Normalize(scale_factor=params.scale_factor,
                              unit_sphere_norm=params.unit_sphere_norm)
    cyl = CylindricalCoordinates(use_octree=True) \
        if params.load_octree and params.model_params.coordinates == 'cylindrical' else None
    model.eval()
    keys = list(data_set)
    if len(keys) == 0:
        # the reference never allocates `embeddings` for an empty set and returns None
        # (eval/pnv_evaluate.py:151, 187); evaluate_dataset() skips such pairs.  Every rank sees the
        # same empty set, so no collective is skipped one-sidedly.
        return None
    rank, world = _dist_info()
    spans = shard_batches(len(keys), params.val_batch_size, rank, world)
    chunks = []

    def read(k):
        return pc_loader(os.path.join(params.dataset_folder, data_set[k]['query']))

    # The reference reads and prepares every submap serially in the main process, in line with the
    # GPU work (eval/pnv_evaluate.py:155-176).  Here the files of the NEXT batches are read and prepared by a
    # thread pool while the GPU embeds the current one, in groups of `grp` submaps per task: one vectorised
    # prepare_batch() pass per group (bit-identical values) keeps the data cache-resident and pays the Python /
    # dispatch cost -- which is what the pool threads queue for, holding the GIL -- once per group instead of
    # once per submap.  Results are consumed strictly in dataset order, so batch composition is unchanged.
    workers = int(os.environ.get('HFL_LOADER_THREADS', min(16, os.cpu_count() or 1)))
    grp = int(os.environ.get('HFL_LOADER_GROUP', 8))
    if os.environ.get('HFL_DEVICE_PREP', '0') == '1' and device_prep_supported(normalize, cyl):
        # opt-in: the pool threads only read files; Normalize / masks / cylindrical run on the GPU for the whole
        # batch (hfl_prepare_clouds), the prepared points never visit the host
        from concurrent.futures import ThreadPoolExecutor
        from ..octree import build_batch_device
        with ThreadPoolExecutor(max_workers=max(workers, 1)) as pool, ThreadPoolExecutor(max_workers=2) as stager:
            def stage_job(b, e):
                files = [pool.submit(read, k) for k in keys[b:e]]
                return stager.submit(lambda: stage_raw_batch([f.result() for f in files]))
            ahead = 3                                            # batches in flight on the host side
            pending = [stage_job(b, e) for _, b, e in spans[:ahead]]
            staged = []

            def embed_oldest():
                pts, off, (n_pin, ev, host) = staged.pop(0)
                ev.synchronize()                                 # H2D + prep kernels of that batch are done
                _PINNED.give(host)
                o = build_batch_device(pts[:int(n_pin[0])], off, params.octree_depth, 2)
                chunks.append(compute_embedding(model, {'octree': o}).float())
            for t in range(len(spans)):
                st = pending.pop(0).result()
                if t + ahead < len(spans):
                    _, b, e = spans[t + ahead]
                    pending.append(stage_job(b, e))
                staged.append(prepare_batch_device(None, params, normalize, cyl, device, staged=st))
                if len(staged) > 1:                              # embed batch t - 1 while batch t's prep is in flight
                    embed_oldest()
            while staged:
                embed_oldest()
        workers = -1                                             # done

    def load_group(ks):
        return prepare_batch([read(k) for k in ks], params, normalize, cyl)

    if workers == -1:
        pass
    elif workers <= 1 or not spans:
        for _, b, e in spans:
            clouds = load_group(keys[b:e])
            chunks.append(compute_embedding(model, collate_batch(clouds, device, params)).float())
    else:
        from concurrent.futures import ThreadPoolExecutor
        # one intra-op thread per torch CPU op while the pool runs: the group tensors are small (a few hundred
        # K points) and N pool threads each fanning out to an OpenMP team oversubscribe the cores
        intra = torch.get_num_threads()
        torch.set_num_threads(1)
        try:
            with ThreadPoolExecutor(max_workers=workers) as pool:
                def batch_job(b, e):
                    return [pool.submit(load_group, keys[i:min(i + grp, e)]) for i in range(b, e, grp)]
                ahead = 2                                        # batches in flight on the host side
                pending = [batch_job(b, e) for _, b, e in spans[:ahead]]
                for t in range(len(spans)):
                    clouds = [c for f in pending.pop(0) for c in f.result()]
                    if t + ahead < len(spans):
                        _, b, e = spans[t + ahead]
                        pending.append(batch_job(b, e))
                    chunks.append(compute_embedding(model, collate_batch(clouds, device, params)).float())
        finally:
            torch.set_num_threads(intra)
    dim = params.model_params.output_dim
    local = torch.cat(chunks) if chunks else torch.zeros((0, dim), device=device)
    return gather_rows(local, spans, len(keys), rank, world, params.val_batch_size).cpu().numpy()


def knn_search(database_output: np.ndarray, queries_output: np.ndarray, k: int = NUM_NEIGHBORS):
    """Exact L2 top-k.  The database is sharded over the ranks, partial lists are
    all-gathered and merged on the device; every rank returns the global result."""
    from .. import ops
    rank, world = _dist_info()
    dev = torch.device('cuda', torch.cuda.current_device())
    q = torch.from_numpy(np.ascontiguousarray(queries_output, dtype=np.float32)).to(dev)
    n_db = len(database_output)
    per = (n_db + world - 1) // world
    lo, hi = min(rank * per, n_db), min((rank + 1) * per, n_db)
    shard = torch.from_numpy(np.ascontiguousarray(database_output[lo:hi], dtype=np.float32)).to(dev)
    if hi > lo:
        d, i = ops.knn_topk(q, shard, k, idx_offset=lo)
    else:
        d = torch.full((len(q), k), float('inf'), device=dev)
        i = torch.full((len(q), k), -1, dtype=torch.int32, device=dev)
    if world > 1:
        ds = torch.empty((world,) + tuple(d.shape), dtype=d.dtype, device=dev)
        is_ = torch.empty((world,) + tuple(i.shape), dtype=i.dtype, device=dev)
        dist.all_gather_into_tensor(ds, d)
        dist.all_gather_into_tensor(is_, i)
        d, i = ops.topk_merge(ds, is_)
    return torch.sqrt(d.clamp(min=0)).cpu().numpy(), i.cpu().numpy().astype(np.int64)


def recall_from_neighbors(indices: np.ndarray, query_set: dict, m: int, n_db: int):
    """Recall bookkeeping of eval/pnv_evaluate.py:236-315 given the retrieved indices."""
    k = indices.shape[1]
    recall = np.zeros(k)
    recall_idx, one_percent_retrieved, num_evaluated = [], 0, 0
    threshold = max(int(round(n_db / 100.0)), 1)
    for i in range(len(indices)):
        true_neighbors = query_set[i][m]
        if len(true_neighbors) == 0:
            continue
        num_evaluated += 1
        truth = set(true_neighbors)
        hits = [j for j in range(k) if indices[i][j] in truth]
        if hits:
            recall[hits[0]] += 1
            recall_idx.append(hits[0] + 1)
        if truth.intersection(indices[i][:threshold].tolist()):
            one_percent_retrieved += 1
    one_percent_recall = (one_percent_retrieved / float(num_evaluated)) * 100
    recall = (np.cumsum(recall) / float(num_evaluated)) * 100
    mrr = np.mean(1 / np.array(recall_idx)) * 100
    return recall, one_percent_recall, mrr


def get_recall(m, n, database_vectors, query_vectors, query_sets, database_sets, log=False,
               model_name: str = 'model'):
    database_output, queries_output = database_vectors[m], query_vectors[n]
    distances, indices = knn_search(database_output, queries_output, NUM_NEIGHBORS)
    if log and _dist_info()[0] == 0:
        _log_search_results(m, n, distances, indices, query_sets, database_sets, model_name)
    return recall_from_neighbors(indices, query_sets[n], m, len(database_output))


def _log_search_results(m, n, distances, indices, query_sets, database_sets, model_name):
    with open(f'{model_name}_log_fp.txt', 'a') as fp, \
            open(f'{model_name}_log_search_results.txt', 'a') as fs:
        for i in range(len(indices)):
            q = query_sets[n][i]
            truth = q[m]
            if len(truth) == 0:
                continue
            wd = lambda e: np.sqrt((q['northing'] - e['northing']) ** 2 + (q['easting'] - e['easting']) ** 2)
            if indices[i][0] not in truth:
                fpe = database_sets[m][indices[i][0]]
                s = '{}, {}, {:0.2f}, {:0.2f}'.format(q['query'], fpe['query'], distances[i, 0], wd(fpe))
                tp = next((k for k in range(indices.shape[1]) if indices[i][k] in truth), None)
                if tp is None:
                    s += ', 0, 0, 0\n'
                else:
                    e = database_sets[m][indices[i][tp]]
                    s += ', {}, {:0.2f}, {:0.2f}\n'.format(e['query'], distances[i][tp], wd(e))
                fp.write(s)
            s = f"{q['query']}, {q['northing']}, {q['easting']}"
            for k in range(min(indices.shape[1], 5)):
                e = database_sets[m][indices[i][k]]
                s += f", {e['query']}, {distances[i][k]:0.2f}, , {wd(e):0.2f}, " \
                     f"{1 if indices[i][k] in truth else 0}, "
            fs.write(s + '\n')


def print_eval_stats(stats):
    for database_name in stats:
        print('Dataset: {}'.format(database_name))
        t = 'Avg. top 1% recall: {:.2f}   Avg. MRR: {:.2f}   Avg. recall @N:'
        print(t.format(stats[database_name]['ave_one_percent_recall'], stats[database_name]['ave_mrr']))
        print(stats[database_name]['ave_recall'])


def pnv_write_eval_stats(file_name, prefix, stats):
    s = prefix
    with open(file_name, 'a') as f:
        for ds in stats:
            s += f'\n[{ds}]\n'
            s += 'AR@1%: {:0.2f}, AR@1: {:0.2f}, MRR: {:0.2f}, AR@N:\n'.format(
                stats[ds]['ave_one_percent_recall'], stats[ds]['ave_recall'][0], stats[ds]['ave_mrr'])
            s += str(stats[ds]['ave_recall'])
        s += '\n------------------------------------------------------------------------\n\n'
        f.write(s)


def main(argv=None, evaluate_fn=None, print_fn=None, write_fn=None, results_suffix='results'):
    parser = argparse.ArgumentParser(description='Evaluate model on PointNetVLAD-protocol test sets')
    parser.add_argument('--config', type=str, required=True, help='Path to configuration file')
    parser.add_argument('--model_config', type=str, required=True,
                        help='Path to the model-specific configuration file')
    parser.add_argument('--weights', type=str, required=False, help='Trained model weights')
    parser.add_argument('--debug', dest='debug', action='store_true')
    parser.add_argument('--visualize', dest='visualize', action='store_true')
    parser.add_argument('--log', dest='log', action='store_true',
                        help='Log false positives and top-5 retrievals')
    args = parser.parse_args(argv)
    if 'RANK' in os.environ and int(os.environ.get('WORLD_SIZE', '1')) > 1:
        torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', 0)))
        dist.init_process_group('nccl')
    rank, _ = _dist_info()
    if rank == 0:
        print('Config path: {}'.format(args.config))
        print('Model config path: {}'.format(args.model_config))
        print('Weights: {}'.format(args.weights or 'RANDOM WEIGHTS'))
    set_seed()
    params = TrainingParams(args.config, args.model_config, debug=args.debug)
    if not torch.cuda.is_available():
        raise RuntimeError('hotformerloc_b200 evaluates on a CUDA device only (no CPU fallback)')
    device = torch.device('cuda', torch.cuda.current_device())
    model = model_factory(params.model_params)
    if args.weights is not None:
        assert os.path.exists(args.weights), 'Cannot open network weights: {}'.format(args.weights)
        state = torch.load(args.weights, map_location='cpu')
        if os.path.splitext(args.weights)[1] == '.ckpt':
            state = state['model_state_dict']
        model.load_state_dict(state)
    model.to(device)
    model_name = os.path.split(args.weights)[1] if args.weights else 'random_init'
    prefix = 'Model Params: {}, Config: {}, Model: {}'.format(
        os.path.split(params.model_params.model_params_path)[1], os.path.split(params.params_path)[1],
        model_name)
    stats = (evaluate_fn or evaluate)(model, device, params, args.log, model_name, show_progress=True)
    if rank == 0:
        (print_fn or print_eval_stats)(stats)
        (write_fn or pnv_write_eval_stats)(f'pnv_{params.dataset_name}_{results_suffix}.txt', prefix, stats)
    if dist.is_available() and dist.is_initialized():
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
