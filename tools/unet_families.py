#!/usr/bin/env python
"""Per-family CUDA-event time of one SD v1-4 U-Net call (in-engine profiler): python tools/unet_families.py [batch] [reps]."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from cycle_diffusion_b200 import specs  # noqa: E402
from cycle_diffusion_b200.engine import Engine, UNet  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
eng = Engine(0)
cfg = specs.sd_unet_config(768)
unet = UNet(eng, cfg, 'openai').load_state_dict(specs.synth_state_dict(specs.openai_unet_params(cfg), 1234))
x = torch.randn(B, 4, 64, 64, device='cuda')
t = torch.full((B,), 501., device='cuda')
ctx = torch.randn(B, 77, 768, device='cuda')
for _ in range(2):
    unet(x, t, ctx)
torch.cuda.synchronize()
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
for _ in range(reps):
    unet(x, t, ctx)
ev1.record()
torch.cuda.synchronize()
print(f'unet B{B}: {ev0.elapsed_time(ev1) / reps:.3f} ms per call (unprofiled)')
eng.profile(True)
for _ in range(reps):
    unet(x, t, ctx)
fam = eng.profile_read()
eng.profile(False)
tot = 0.0
for k, v in sorted(fam.items(), key=lambda kv: -kv[1]['ms']):
    ms = v['ms'] / reps
    tot += ms
    tf = v['flops'] / (v['ms'] * 1e-3) / 1e12 if v['flops'] > 0 and v['ms'] > 0 else 0.0
    print(f'{k:24s} {ms:8.3f} ms  {v.get("launches", 0) // reps if "launches" in v else "":>5}  {tf:7.1f} TFLOP/s')
print(f'sum of families {tot:.3f} ms')
