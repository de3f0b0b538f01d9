"""k_knn at the BASELINE.json configs[3] size: 8192 queries x 8192 database rows x 256 dims, k = 25."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from hotformerloc_b200 import ops
nq = ndb = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
torch.manual_seed(0)
db = F.normalize(torch.randn(ndb, 256, device='cuda'), dim=1)
q = F.normalize(torch.randn(nq, 256, device='cuda'), dim=1)
for _ in range(3): ops.knn_topk(q, db, 25)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(10): ops.knn_topk(q, db, 25)
e.record(); torch.cuda.synchronize()
ms = s.elapsed_time(e) / 10
# CUDA-core fp32: one subtract + one FMA per (query, row, dim) = 3 flop; B200 fp32 peak = 148 SM x 128 lanes x 2 x ~1.9 GHz
flop = 3.0 * nq * ndb * 256
print(json.dumps({'kernel': 'k_knn + k_topk_merge', 'nq': nq, 'ndb': ndb, 'dim': 256, 'k': 25, 'ms': ms,
                  'fp32_tflops': flop / ms / 1e9, 'fp32_peak_tflops_nominal': 148 * 128 * 2 * 1.9e9 / 1e12,
                  'frac_of_fp32_issue': (2.0 * nq * ndb * 256 / 32) / (ms * 1e-3 * 1.9e9 * 148 * 4)}))
