#!/usr/bin/env python
"""Run the HBM-bound normalisation / softmax kernels at the 64x64-level shapes a few times (for ncu):
    python tools/one_norm.py gn|ln|cross [batch]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from cycle_diffusion_b200.engine import Engine  # noqa: E402

kind = sys.argv[1]
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
eng = Engine(0)
if kind == 'gn':
    x = torch.randn(B, 64, 64, 320, device='cuda')
    g, b = torch.randn(320, device='cuda'), torch.randn(320, device='cuda')
    for _ in range(4):
        eng.op_groupnorm(x, g, b, 1e-5, True)
elif kind == 'ln':
    x = torch.randn(B * 4096, 320, device='cuda')
    g, b = torch.randn(320, device='cuda'), torch.randn(320, device='cuda')
    for _ in range(4):
        eng.op_layernorm(x, g, b)
else:   # cross-attention: 4096 queries x 77 keys, 8 heads x 40
    q = torch.randn(B, 4096, 320, device='cuda')
    k, v = torch.randn(B, 77, 320, device='cuda'), torch.randn(B, 77, 320, device='cuda')
    for _ in range(4):
        eng.op_attention(q, k, v, 8, 40 ** -0.5)
torch.cuda.synchronize()
print('done')
