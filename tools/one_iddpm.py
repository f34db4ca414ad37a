#!/usr/bin/env python
"""Per-family timing of one improved-DDPM U-Net call (BASELINE cfg5 shape: 256x256, batch 8 per GPU):
    CDX_PROF_DUMP=1 python tools/one_iddpm.py [image_size] [batch]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from cycle_diffusion_b200 import specs  # noqa: E402
from cycle_diffusion_b200.engine import Engine, UNet  # noqa: E402

R = int(sys.argv[1]) if len(sys.argv) > 1 else 256
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
eng = Engine(0)
cfg = specs.iddpm_config(R)
unet = UNet(eng, cfg, 'iddpm').load_state_dict(specs.synth_state_dict(specs.iddpm_unet_params(cfg), 31))
x = torch.randn(B, 3, R, R, device='cuda')
t = torch.full((B,), 500., device='cuda')
for _ in range(2):
    unet(x, t, None)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3):
    unet(x, t, None)
e1.record()
torch.cuda.synchronize()
print('ms per call', e0.elapsed_time(e1) / 3, 'workspace GB', eng.workspace_bytes / 2 ** 30)
eng.profile(True)
unet(x, t, None)
fam = eng.profile_read()
eng.profile(False)
for k, v in fam.items():
    tf = v['flops'] / (v['ms'] * 1e-3) / 1e12 if v['flops'] else 0
    print(f'{k:14s} {v["ms"]:8.3f} ms {v["launches"]:4d} launches {tf:7.1f} TF/s')
