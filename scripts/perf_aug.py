import os, sys, torch
sys.path.insert(0, os.getcwd())
from i2v_adapter_unofficial_b200 import ops
torch.manual_seed(3)
Bv, Fr, H, S, d = 2, 16, 8, 4096, 40
mk = lambda b: torch.randn(b, S, H, d, device="cuda", dtype=torch.bfloat16)
qa, ka, va = ops.augment_qkv(mk(Bv*Fr), mk(Bv*Fr), mk(Bv*Fr))
qxa, kxa, vxa = ops.augment_qkv(mk(Bv*Fr), mk(Bv), mk(Bv))
for _ in range(4):
    ops.fused_self_xframe_aug(qa, ka, va, qxa, kxa, vxa, Fr)
torch.cuda.synchronize()
