#!/usr/bin/env python
"""Micro-benchmark of the dense-contraction kernels on the SD v1-4 layer shapes (batch 8 = the CFG launch shape).

Uses the in-engine CUDA-event profiler, so weight repacking etc. is excluded.  Run on the GPU box:
    python tools/bench_ops.py [--mode 0|1] [--batch 8]
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from cycle_diffusion_b200.engine import Engine  # noqa: E402

CONVS = [(320, 320, 64), (640, 640, 32), (1280, 1280, 16), (1280, 1280, 8), (2560, 1280, 8), (2560, 1280, 16), (1920, 1280, 16),
         (1920, 640, 32), (1280, 640, 32), (960, 640, 32), (960, 320, 64), (640, 320, 64), (320, 640, 32), (640, 1280, 16)]
LINEARS = [(4096, 320, 320), (4096, 320, 960), (4096, 320, 2560), (4096, 1280, 320), (1024, 640, 640), (1024, 640, 5120), (1024, 2560, 640),
           (256, 1280, 1280), (256, 1280, 10240), (256, 5120, 1280), (64, 1280, 1280)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mode", type=int, default=1, help="1 fp16-split (default), 3 3xTF32, 4 fast path, 0 FFMA")
    ap.add_argument('--batch', type=int, default=8)
    ap.add_argument('--reps', type=int, default=3)
    a = ap.parse_args()
    eng = Engine(0)
    eng.set_mma_mode(a.mode)
    rows = []
    for cin, cout, h in CONVS:
        x = torch.randn(a.batch, h, h, cin, device='cuda')
        w = torch.randn(cout, cin, 3, 3, device='cuda') * 0.02
        b = torch.randn(cout, device='cuda')
        eng.op_conv3x3(x, w, b)
        eng.profile(True)
        for _ in range(a.reps):
            eng.op_conv3x3(x, w, b)
        fam = eng.profile_read()
        eng.profile(False)
        for k, v in fam.items():
            if v['flops'] > 0:
                rows.append(dict(op=f'conv3x3 {cin}->{cout} @{h}^2 B{a.batch}', kernel=k, ms=round(v['ms'] / a.reps, 3),
                                 tflops=round(v['flops'] / (v['ms'] * 1e-3) / 1e12, 1)))
    for m, k, n in LINEARS:
        M = m * a.batch
        x = torch.randn(M, k, device='cuda')
        w = torch.randn(n, k, device='cuda') * 0.02
        b = torch.randn(n, device='cuda')
        eng.op_linear(x, w, b)
        eng.profile(True)
        for _ in range(a.reps):
            eng.op_linear(x, w, b)
        fam = eng.profile_read()
        eng.profile(False)
        for kk, v in fam.items():
            if v['flops'] > 0:
                rows.append(dict(op=f'linear M{M} K{k} N{n}', kernel=kk, ms=round(v['ms'] / a.reps, 3),
                                 tflops=round(v['flops'] / (v['ms'] * 1e-3) / 1e12, 1)))
    for r in rows:
        print(f"{r['op']:40s} {r['kernel']:14s} {r['ms']:9.3f} ms {r['tflops']:8.1f} TFLOP/s")
    print(json.dumps(rows))


if __name__ == '__main__':
    main()
