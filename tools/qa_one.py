"""one launch of the fused qkv + attention kernel at the headline level-0 shape (profiling aid)"""
import math, sys, torch
sys.path.insert(0, '.')
from hotformerloc_b200 import ops
DEV = 'cuda'
K, H, C, hat = 48, 16, 256, True
n_win = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
rows = n_win * (K + 1)
torch.manual_seed(0)
xyz = torch.randint(0, 64, (n_win * K, 3), dtype=torch.int16)
bid = torch.sort(torch.randint(0, 256, (n_win * K,))).values.to(torch.int16)
tok = torch.cat([xyz, bid[:, None]], 1).contiguous().to(DEV)
y = torch.randn(rows, C, device=DEV).bfloat16()
W = (torch.randn(3 * C, C, device=DEV) / math.sqrt(C)).bfloat16()
b = torch.randn(3 * C, device=DEV) * 0.1
bnd = 38
rpe = torch.randn(3 * (2 * bnd + 1), H, device=DEV) * 0.5
Wg, bg = ops.regroup_qkv(W, b)
out = torch.zeros(rows, C, device=DEV, dtype=torch.bfloat16)
import os
CODES = ops.qkv_attn_codes(tok, n_win, K, 1, hat, bnd, True) if os.environ.get('QA_CODES', '1') == '1' else None
for _ in range(3):
    ops.qkv_attn(y, Wg, bg, out, tok, rpe, n_win, H, C, K, 1, hat, bnd, 0.25, codes=CODES)
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(5):
    ops.qkv_attn(y, Wg, bg, out, tok, rpe, n_win, H, C, K, 1, hat, bnd, 0.25, codes=CODES)
e.record(); torch.cuda.synchronize()
print('qkv_attn n_win=%d: %.3f ms' % (n_win, s.elapsed_time(e) / 5))
