"""ncu target: the token GEMM at one shape: rows K N [residual]."""
import os, sys, torch
sys.path.insert(0, os.getcwd())
from i2v_adapter_unofficial_b200 import ops
rows, K, N = (int(a) for a in sys.argv[1:4])
res = len(sys.argv) > 4
x = torch.randn(rows, K, device="cuda").to(torch.bfloat16)
w = (torch.randn(N, K, device="cuda") * K ** -0.5).to(torch.bfloat16)
r = torch.randn(rows, N, device="cuda").to(torch.bfloat16) if res else None
for _ in range(4):
    ops.linear(x, w, None, r, out=r)
torch.cuda.synchronize()
