"""ncu target: the fused feed-forward GEMM at the C2 level-0 shape (131072 x 320 -> 1280)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from i2v_adapter_unofficial_b200 import ops
torch.manual_seed(0)
rows, C = (131072, 320) if len(sys.argv) < 3 else (int(sys.argv[1]), int(sys.argv[2]))
x = torch.randn(rows, C, device="cuda", dtype=torch.bfloat16)
w = (torch.randn(8 * C, C, device="cuda") * C ** -0.5).to(torch.bfloat16)
b = (torch.randn(8 * C, device="cuda") * 0.1).to(torch.bfloat16)
for _ in range(4):
    ops.ff_geglu(x, w, b, ones_column=True)
torch.cuda.synchronize()
