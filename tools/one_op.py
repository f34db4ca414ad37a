#!/usr/bin/env python
"""Run ONE conv3x3 / linear launch shape a few times (for ncu captures):  python tools/one_op.py conv 640 640 32 8 [mode]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from cycle_diffusion_b200.engine import Engine  # noqa: E402

kind = sys.argv[1]
a = [int(v) for v in sys.argv[2:]]
eng = Engine(0)
eng.set_mma_mode(a[4] if len(a) > 4 else 1)
if kind == 'conv':
    cin, cout, h, b = a[:4]
    x = torch.randn(b, h, h, cin, device='cuda')
    w = torch.randn(cout, cin, 3, 3, device='cuda') * 0.02
    bias = torch.randn(cout, device='cuda')
    for _ in range(3):
        eng.op_conv3x3(x, w, bias)
else:
    m, k, n = a[:3]
    x = torch.randn(m, k, device='cuda')
    w = torch.randn(n, k, device='cuda') * 0.02
    bias = torch.randn(n, device='cuda')
    for _ in range(3):
        eng.op_linear(x, w, bias)
torch.cuda.synchronize()
print('done')
