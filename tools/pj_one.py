"""one launch of the fused proj + LN + MLP kernel at the headline level-0 shape (profiling aid)"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hotformerloc_b200 import ops
M, C = int(sys.argv[1]) if len(sys.argv) > 1 else 980_000, 256
g = lambda *s: torch.randn(*s, device='cuda')
O = g(M, C).bfloat16()
Wp, W1, W2 = (g(C, C) / 16).bfloat16(), (g(4 * C, C) / 16).bfloat16(), (g(C, 4 * C) / 32).bfloat16()
bp, b1, b2, lg, lb = g(C), g(4 * C), g(C), 1 + 0.1 * g(C), 0.1 * g(C)
x = g(M, C); xb = torch.empty(M, C, device='cuda', dtype=torch.bfloat16)
run = lambda: ops.proj_mlp_fused(O, Wp, bp, lg, lb, W1, b1, W2, b2, res=x, out_f32=x, out_bf16=xb)
for _ in range(3): run()
torch.cuda.synchronize()
s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
s.record()
for _ in range(5): run()
e.record(); torch.cuda.synchronize()
ms = s.elapsed_time(e) / 5
print('proj_mlp_fused M=%d: %.3f ms  %.1f TFLOP/s' % (M, ms, 18 * C * C * M / ms / 1e9))
