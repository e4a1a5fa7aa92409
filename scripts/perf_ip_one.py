import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from i2v_adapter_unofficial_b200 import ops
B, S, d, H, g, nt, ni = 32, 4096, 40, 8, 16, 77, 4
q = torch.randn(B, S, H, d, device="cuda", dtype=torch.bfloat16)
kv = torch.randn(B // g, nt + ni, 2, H, d, device="cuda", dtype=torch.bfloat16)
for _ in range(4):
    ops.ip_xattn(q, kv[:, :, 0], kv[:, :, 1], nt, 1.0, g, None, ops.MODE_FAST)
torch.cuda.synchronize()
