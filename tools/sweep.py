"""BASELINE.json configs[4]: point-count sweep (4 K - 256 K points / submap, depth 9) and window-size
sweep (patch_size K) of the hot path on one B200, per kernel family against both roofs.

    python tools/sweep.py [--out gpurun_out/sweep.json]

Every line: one (points, batch, K) cell -- octree build (+ neighbour tables) time, forward time,
submaps/s, points/s and the time / algorithmic HBM + tensor rates of the window-attention, RTSA
(varlen) attention, CPE and GEMM families (CUDA events on the launching stream, bench.kernel_profile).
Batches keep ~1 M points per step so the cells are comparable.  K = 96 (97 keys with the relay token) runs on
the mma.sync window-attention kernel (13 key tiles); K + relay token <= 64 on the fused tcgen05 kernel.

Under torchrun (one rank per GPU) every rank runs the same cells on its own seeded clouds (weak scaling, no
data-path collective); the per-cell times are the MAX over ranks and submaps/s the whole-job aggregate.
"""
import argparse
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch

import bench


def make_model(K, device):
    from hotformerloc_b200.config import presets
    from hotformerloc_b200.misc.utils import ModelParams
    from hotformerloc_b200.models.model_factory import model_factory
    d = tempfile.mkdtemp(prefix='hfl_sweep_')
    name = 'oxford'
    saved = presets.MODEL_PRESETS[name]
    presets.MODEL_PRESETS[name] = dict(saved, patch_size=K)
    try:
        paths = presets.write_configs(d, name, dataset_folder=d)
    finally:
        presets.MODEL_PRESETS[name] = saved
    torch.manual_seed(0)
    return model_factory(ModelParams(paths['model_config'])).to(device).eval()


def cell(model, P, B, depth, steps, device, peaks, rank=0, world=1):
    from hotformerloc_b200.octree import build_batch_device
    clouds = bench.synthetic_batches(2, B, P, seed0=7000 + P % 977 + 131 * rank)
    devb = []
    for cl in clouds:
        pts = torch.from_numpy(np.concatenate(cl)).to(device)
        off = torch.tensor(np.concatenate([[0], np.cumsum([len(c) for c in cl])]), dtype=torch.int32,
                           device=device)
        devb.append((pts, off))
    ev = lambda: torch.cuda.Event(enable_timing=True)
    for i in range(2):
        o = build_batch_device(*devb[i % 2], depth, 2)
        model({'octree': o})
    torch.cuda.synchronize()
    tb = tf = 0.0
    for i in range(steps):
        a, b, c = ev(), ev(), ev()
        a.record()
        o = build_batch_device(*devb[i % 2], depth, 2)
        o.finalize()
        b.record()
        model({'octree': o})
        c.record()
        torch.cuda.synchronize()
        tb += a.elapsed_time(b)
        tf += b.elapsed_time(c)
    tb, tf = tb / steps, tf / steps
    if world > 1:
        import torch.distributed as dist
        t = torch.tensor([tb, tf], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        tb, tf = float(t[0]), float(t[1])
    fam = bench.kernel_profile(model, lambda: build_batch_device(*devb[0], depth, 2))
    n = [o.n(d) for d in range(depth + 1)]
    out = {'points': P, 'batch': B, 'octree_depth': depth, 'nodes_leaf': n[depth], 'tokens_d0': n[depth - 2],
           'build_ms': round(tb, 3), 'forward_ms': round(tf, 3),
           'n_gpus': world, 'submaps_per_s': round(world * B / ((tb + tf) / 1e3), 1),
           'points_per_s': round(world * B * P / ((tb + tf) / 1e3)),
           'build_points_per_s': round(world * B * P / (tb / 1e3))}
    for k in ('qkv_attn', 'window_attn', 'varlen_attn', 'cpe_ln', 'gather_gemm', 'proj_mlp_fused', 'mlp_fused'):
        if k in fam:
            f = fam[k]
            r = {'ms': round(f['ms'], 3)}
            if f.get('byte'):
                r['hbm_frac'] = round(f['byte'] / (f['ms'] / 1e3) / 1e9 / peaks['hbm'], 3)
            if f.get('flop'):
                r['tensor_frac'] = round(f['flop'] / (f['ms'] / 1e3) / 1e12 / peaks['tf'], 3)
            out[k] = r
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--out', default=os.path.join(ROOT, 'gpurun_out', 'sweep.json'))
    ap.add_argument('--steps', type=int, default=3)
    args = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (('RANK', 0), ('WORLD_SIZE', 1), ('LOCAL_RANK', 0)))
    device = torch.device('cuda', local)
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault('MASTER_ADDR', '127.0.0.1')
        dist.init_process_group('nccl', device_id=device)
    try:
        pk = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pk = {}
    peaks = {'hbm': pk.get('hbm_gbs', 6650.0), 'tf': pk.get('bf16_tflops_sustained', 1400.0)}
    rows = []
    model = make_model(48, device)
    for P, B in ((4096, 256), (16384, 64), (65536, 16), (262144, 4)):
        r = cell(model, P, B, 9, args.steps, device, peaks, rank, world)
        r['patch_size'] = 48
        rows.append(r)
        if rank == 0:
            print(json.dumps(r), flush=True)
    for K in (32, 48, 64, 96):
        try:
            m = make_model(K, device)
            r = cell(m, 4096, 256, 9, args.steps, device, peaks, rank, world)
            r['patch_size'] = K
        except Exception as e:
            r = {'points': 4096, 'batch': 256, 'patch_size': K, 'unsupported': str(e)[:200]}
        rows.append(r)
        if rank == 0:
            print(json.dumps(r), flush=True)
    if rank == 0:
        os.makedirs(os.path.dirname(args.out), exist_ok=True)
        json.dump({'peaks': peaks, 'n_gpus': world, 'cells': rows}, open(args.out, 'w'), indent=1)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
