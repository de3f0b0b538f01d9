"""Multi-GPU correctness check (run under torchrun, one rank per GPU):
sharded evaluation == single-rank evaluation, sharded top-k == global top-k."""
import os, sys, pickle, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist

rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
torch.cuda.set_device(local)
dist.init_process_group('nccl', device_id=torch.device('cuda', local))
from oracle.model_ref import lidar_cloud
from hotformerloc_b200.config.presets import write_configs
from hotformerloc_b200.eval import pnv_evaluate as E
from hotformerloc_b200.misc.utils import TrainingParams
from hotformerloc_b200.models.model_factory import model_factory
from hotformerloc_b200 import ops

root = '/tmp/hfl_mgpu'
n = 37
if rank == 0:
    os.makedirs(root + '/run0', exist_ok=True)
    g = torch.Generator().manual_seed(5)
    for i in range(n):
        lidar_cloud(4096, g).astype(np.float64).tofile(f'{root}/run0/{i}.bin')
dist.barrier()
paths = write_configs(root + f'/cfg{rank}', 'oxford', dataset_folder=root)
cfg = open(paths['config']).read().replace('val_batch_size=256', 'val_batch_size=8')
open(paths['config'], 'w').write(cfg)
params = TrainingParams(paths['config'], paths['model_config'])
torch.manual_seed(0)
model = model_factory(params.model_params).cuda().eval()
data_set = {i: {'query': f'run0/{i}.bin'} for i in range(n)}
emb = E.get_latent_vectors(model, data_set, 'cuda', params)           # sharded over the ranks
# single-rank reference: all batches on this rank, same composition
loader = E.PNVPointCloudLoader()
chunks = []
for _, b, e in E.shard_batches(n, 8, 0, 1):
    clouds = [E.prepare_cloud(loader(f'{root}/run0/{i}.bin'), params) for i in range(b, e)]
    chunks.append(E.compute_embedding(model, E.collate_batch(clouds, 'cuda', params)).float())
ref = torch.cat(chunks).cpu().numpy()
assert np.array_equal(emb, ref), np.abs(emb - ref).max()
rng = np.random.default_rng(1)
db = rng.normal(size=(1001, 256)).astype(np.float32)
q = rng.normal(size=(333, 256)).astype(np.float32)
d, i = E.knn_search(db, q, 25)
d1, i1 = ops.knn_topk(torch.from_numpy(q).cuda(), torch.from_numpy(db).cuda(), 25)
assert np.array_equal(i, i1.cpu().numpy())
print(f'rank {rank}/{world}: sharded evaluation and top-k identical to single-rank', flush=True)
dist.destroy_process_group()
