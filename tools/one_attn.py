#!/usr/bin/env python
"""Run one fused-attention launch shape a few times (for ncu):  python tools/one_attn.py B N heads d [mode]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from cycle_diffusion_b200.engine import Engine  # noqa: E402

B, N, heads, d = [int(v) for v in sys.argv[1:5]]
eng = Engine(0)
eng.set_mma_mode(int(sys.argv[5]) if len(sys.argv) > 5 else 1)
C = heads * d
q, k, v = (torch.randn(B, N, C, device='cuda') for _ in range(3))
for _ in range(3):
    eng.op_attention(q, k, v, heads, d ** -0.5)
torch.cuda.synchronize()
eng.profile(True)
for _ in range(5):
    eng.op_attention(q, k, v, heads, d ** -0.5)
print({k_: round(v_['ms'] / 5, 3) for k_, v_ in eng.profile_read().items()})
