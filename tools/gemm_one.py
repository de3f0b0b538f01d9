import math, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hotformerloc_b200 import ops
M, N, K = 1_050_000, 768, 256
A = torch.randn(M, K, device='cuda').to(torch.bfloat16)
W = (torch.randn(N, K, device='cuda') / 16).to(torch.bfloat16)
bias = torch.randn(N, device='cuda')
out = torch.empty(M, N, device='cuda', dtype=torch.bfloat16)
for _ in range(3):
    ops.gather_gemm(A, W, bias=bias, out_v_bf16=out)
torch.cuda.synchronize()
