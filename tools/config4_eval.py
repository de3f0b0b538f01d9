"""BASELINE.json configs[3]: Wild-Places cfg, database + query embedding sharded across the GPUs
of one box, NCCL descriptor all-gather, sharded exact top-25, recall@1 / @1 %.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        tools/config4_eval.py [--runs 4 --per-run 256 --points 30000 --out gpurun_out/config4.json]
    (also runs as a plain single process)

Synthetic dataset in the reference's on-disk format: `runs` traversals of
the same trajectory of `per-run` places; every submap is a binary `.pcd` file (read by the native
PCD reader), every run an evaluation-pickle dict {idx: {'query': relpath, 'northing', 'easting',
<db run>: [true neighbour ids]}}.  A place is one lidar-ish cloud; a traversal re-observes it with
point jitter, a small yaw and 10 % of the points dropped.  Through the public entry points
(eval.pnv_evaluate.get_latent_vectors / get_recall) this exercises: file loaders -> Normalize ->
range mask -> cylindrical coordinates -> batched device octree build -> forward -> descriptor
all-gather -> database-sharded top-k -> all-gather + merge of the partial lists -> recall.

Rank 0 additionally re-embeds run 0 alone (all batches on one GPU, same batch composition) and checks
that the sharded descriptors are bitwise identical, and repeats one (db, query) search with the
unsharded top-k kernel and checks the indices are identical.  One JSON line + --out.
"""
import argparse
import json
import os
import pickle
import shutil
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch
import torch.distributed as dist


from hotformerloc_b200.datasets.synthetic import make_eval_dataset as make_dataset  # noqa: E402


def main(argv=None, emit=True):
    ap = argparse.ArgumentParser()
    ap.add_argument('--runs', type=int, default=4)
    ap.add_argument('--per-run', type=int, default=256)
    ap.add_argument('--points', type=int, default=30000)
    ap.add_argument('--root', default='/tmp/hfl_config4')
    ap.add_argument('--out', default=os.path.join(ROOT, 'gpurun_out', 'config4.json'))
    args = ap.parse_args(argv)
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (('RANK', 0), ('WORLD_SIZE', 1), ('LOCAL_RANK', 0)))
    torch.cuda.set_device(local)
    own_pg = world > 1 and not dist.is_initialized()
    if own_pg:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    from hotformerloc_b200 import ops
    from hotformerloc_b200.config.presets import write_configs
    from hotformerloc_b200.eval import pnv_evaluate as E
    from hotformerloc_b200.misc.utils import TrainingParams
    from hotformerloc_b200.models.model_factory import model_factory

    t0 = time.time()
    if rank == 0:
        shutil.rmtree(args.root, ignore_errors=True)
        sets = make_dataset(args.root, args.runs, args.per_run, args.points)
        pickle.dump(sets, open(os.path.join(args.root, 'sets.pickle'), 'wb'))
    if world > 1:
        dist.barrier()
    sets = pickle.load(open(os.path.join(args.root, 'sets.pickle'), 'rb'))
    t_gen = time.time() - t0
    paths = write_configs(os.path.join(args.root, f'cfg{rank}'), 'wild-places', dataset_folder=args.root)
    params = TrainingParams(paths['config'], paths['model_config'])
    torch.manual_seed(0)
    model = model_factory(params.model_params).cuda().eval()

    E.get_latent_vectors(model, {0: sets[0][0], 1: sets[0][1]}, 'cuda', params)       # warm-up
    torch.cuda.synchronize()
    t0 = time.time()
    emb = [E.get_latent_vectors(model, s, 'cuda', params) for s in sets]
    torch.cuda.synchronize()
    t_embed = time.time() - t0
    t0 = time.time()
    recalls, oprs, mrrs = [], [], []
    for m in range(args.runs):
        for n in range(args.runs):
            if m == n and params.skip_same_run:
                continue
            r, opr, mrr = E.get_recall(m, n, emb, emb, sets, sets)
            recalls.append(r)
            oprs.append(opr)
            mrrs.append(mrr)
    torch.cuda.synchronize()
    t_search = time.time() - t0
    out = None
    checks = {}
    if rank == 0:
        # (1) sharded descriptors == one-GPU descriptors with the reference's batch composition
        loader = E.CSWildPlacesPointCloudLoader()
        norm = E.Normalize(scale_factor=params.scale_factor, unit_sphere_norm=params.unit_sphere_norm) \
            if (params.normalize_points or params.scale_factor is not None) else None
        cyl = E.CylindricalCoordinates(use_octree=True) if params.model_params.coordinates == 'cylindrical' else None
        keys = list(sets[0])
        chunks = []
        for _, b, e in E.shard_batches(len(keys), params.val_batch_size, 0, 1):
            clouds = [E.prepare_cloud(loader(os.path.join(args.root, sets[0][k]['query'])), params, norm, cyl)
                      for k in keys[b:e]]
            chunks.append(E.compute_embedding(model, E.collate_batch(clouds, 'cuda', params)).float())
        solo = torch.cat(chunks).cpu().numpy()
        checks['sharded_descriptors_bitwise_equal_single_gpu'] = bool(np.array_equal(solo, emb[0]))
        checks['descriptor_max_abs_diff'] = float(np.abs(solo - emb[0]).max())
        checks['unit_norm_max_dev'] = float(np.abs(np.linalg.norm(emb[0], axis=1) - 1).max())
    # (2) database-sharded search == unsharded search (collective inside: every rank takes part)
    _, idx = E.knn_search(emb[0], emb[1], 25)
    if rank == 0:
        _, idx1 = ops.knn_topk(torch.from_numpy(emb[1]).cuda(), torch.from_numpy(emb[0]).cuda(), 25)
        checks['sharded_topk_equal_unsharded'] = bool(np.array_equal(idx, idx1.cpu().numpy()))
        n_sub = args.runs * args.per_run
        out = {'config': 'wild-places cfg (cylindrical, K=48, no ADaPE), synthetic dataset '
                         f'{args.runs} runs x {args.per_run} places x ~{int(0.9 * args.points)} points (binary .pcd)',
               'n_gpus': world, 'submaps': n_sub,
               'recall_at_1': float(np.mean([r[0] for r in recalls])),
               'recall_at_5': float(np.mean([r[4] for r in recalls])),
               'recall_at_1pct': float(np.mean(oprs)), 'mrr': float(np.mean(mrrs)),
               'pairs': len(recalls),
               'embed_seconds_incl_file_io_and_host_prep': round(t_embed, 3),
               'embed_submaps_per_s_incl_file_io': round(n_sub / t_embed, 1),
               'search_seconds': round(t_search, 3), 'dataset_write_seconds': round(t_gen, 1),
               'checks': checks}
        if emit:
            print(json.dumps(out))
        os.makedirs(os.path.dirname(args.out), exist_ok=True)
        json.dump(out, open(args.out, 'w'), indent=1)
    if world > 1:
        dist.barrier()
        if own_pg:
            dist.destroy_process_group()
    return out if rank == 0 else None


if __name__ == '__main__':
    main()
