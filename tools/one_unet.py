#!/usr/bin/env python
"""Run a few SD U-Net calls at the CFG launch shape (batch 8) -- for `ncu --metrics gpu__time_duration.sum` launch lists."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from cycle_diffusion_b200 import specs  # noqa: E402
from cycle_diffusion_b200.engine import Engine, UNet  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2
eng = Engine(0)
cfg = specs.sd_unet_config(768)
unet = UNet(eng, cfg, 'openai').load_state_dict(specs.synth_state_dict(specs.openai_unet_params(cfg), 1234))
x = torch.randn(B, 4, 64, 64, device='cuda')
t = torch.full((B,), 501., device='cuda')
ctx = torch.randn(B, 77, 768, device='cuda')
for _ in range(n):
    unet(x, t, ctx)
torch.cuda.synchronize()
print('launches', eng.launches)
