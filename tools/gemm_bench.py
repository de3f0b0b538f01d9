"""Micro-benchmark of the gather-GEMM over the shapes of the Oxford B=256 forward."""
import math, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hotformerloc_b200 import ops

dev = 'cuda'
torch.manual_seed(0)

def bench(name, M, N, K, KD=1, **kw):
    Cin = K // KD
    rows = M if KD == 1 else M
    A = torch.randn(rows, Cin, device=dev).to(torch.bfloat16)
    W = (torch.randn(N, K, device=dev) / math.sqrt(K)).to(torch.bfloat16)
    idx = None
    if KD > 1:
        idx = torch.randint(-1, rows, (M, KD), device=dev, dtype=torch.int32)
        # realistic locality: neighbours close to the row
        base = torch.arange(M, device=dev)[:, None] + torch.randint(-50, 50, (M, KD), device=dev)
        idx = torch.where(torch.rand(M, KD, device=dev) < 0.4, base.clamp(0, rows - 1), torch.full_like(base, -1)).to(torch.int32)
    bias = torch.randn(N, device=dev)
    args = dict(bias=bias)
    if kw.get('bf16out', True):
        args['out_v_bf16'] = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    if kw.get('res'):
        x = torch.randn(M, N, device=dev)
        args.update(res=x, out_v_f32=x)
    if kw.get('ln'):
        args['ln'] = (torch.ones(N, device=dev), torch.zeros(N, device=dev))
        args['out_y_bf16'] = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    if kw.get('act'):
        args['act'] = 1
    f = lambda: ops.gather_gemm(A, W, idx=idx, KD=KD, **args)
    for _ in range(3): f()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 10
    s.record()
    for _ in range(n): f()
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / n
    tf = 2.0 * M * N * K / ms / 1e9
    byt = M * K * 2 + M * N * (2 if 'out_v_bf16' in args else 0) + (M * N * 8 if kw.get('res') else 0) + (M * N * 2 if kw.get('ln') else 0)
    print(f'{name:28s} M={M:8d} N={N:5d} K={K:5d} KD={KD:2d}  {ms:8.3f} ms  {tf:7.1f} TFLOP/s  {byt/ms/1e6:7.0f} GB/s')

M1 = 1_050_000   # hat rows level 0 (B=256)
bench('qkv C256', M1, 768, 256)
bench('proj C256 +res+ln', M1, 256, 256, res=True, ln=True, bf16out=False)
bench('proj C256 plain', M1, 256, 256)
bench('proj C256 +res f32out', M1, 256, 256, res=True, bf16out=False)
bench('proj C256 +ln bf16 only', M1, 256, 256, ln=True, bf16out=False)
bench('fc1 C256 gelu', M1, 1024, 256, act=True)
bench('fc1 C256 nogelu', M1, 1024, 256)
bench('fc2 C256 +res', M1, 256, 1024, res=True)
bench('fc2 C256 plain', M1, 256, 1024)
M0 = 1_032_000
bench('qkv C128', M0, 384, 128)
bench('proj C128 +res+ln', M0, 128, 128, res=True, ln=True, bf16out=False)
bench('fc1 C128 gelu', M0, 512, 128, act=True)
bench('fc2 C128 +res', M0, 128, 512, res=True)
bench('conv 27x128->128', M0, 128, 3456, KD=27, ln=True, bf16out=False)
bench('conv 27x64->64', 1_047_000, 64, 1728, KD=27, ln=True, bf16out=False)
bench('down 8x128->256', 955_000, 256, 1024, KD=8, ln=True, bf16out=False)
bench('down 8x256->256', 706_000, 256, 2048, KD=8, ln=True, bf16out=False)
bench('small M qkv', 43_000, 768, 256)

# --- epilogue decomposition experiments (QKV shape) ---
def raw(name, M, N, K, **args):
    A = torch.randn(M, K, device=dev).to(torch.bfloat16)
    W = (torch.randn(N, K, device=dev) / math.sqrt(K)).to(torch.bfloat16)
    f = lambda: ops.gather_gemm(A, W, **args)
    for _ in range(3): f()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(10): f()
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 10
    print(f'{name:28s} {ms:8.3f} ms  {2.0*M*N*K/ms/1e9:7.1f} TFLOP/s')

raw('qkv: no outputs at all', M1, 768, 256)
raw('qkv: bf16 out, no bias', M1, 768, 256, out_v_bf16=torch.empty(M1, 768, device=dev, dtype=torch.bfloat16))
raw('qkv: f32 out, no bias', M1, 768, 256, out_v_f32=torch.empty(M1, 768, device=dev))
raw('N=256 K=256 no outputs', M1, 256, 256)
raw('N=256 K=1024 no outputs', M1, 256, 1024)

# --- fused MLP ---
def mlp(name, M, C):
    y = torch.randn(M, C, device=dev).to(torch.bfloat16)
    W1 = (torch.randn(4 * C, C, device=dev) / math.sqrt(C)).to(torch.bfloat16)
    W2 = (torch.randn(C, 4 * C, device=dev) / math.sqrt(4 * C)).to(torch.bfloat16)
    b1, b2 = torch.randn(4 * C, device=dev), torch.randn(C, device=dev)
    x = torch.randn(M, C, device=dev)
    xb = torch.empty(M, C, device=dev, dtype=torch.bfloat16)
    f = lambda: ops.mlp_fused(y, W1, b1, W2, b2, res=x, out_f32=x, out_bf16=xb)
    for _ in range(3): f()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(10): f()
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / 10
    print(f'{name:28s} M={M:8d} C={C}  {ms:8.3f} ms  {16.0*M*C*C/ms/1e9:7.1f} TFLOP/s')

mlp('fused MLP C256', M1, 256)
mlp('fused MLP C128', M0, 128)
