"""Shared helpers for the test-suite: golden fixture loading and the small test configurations."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')

# small configurations used by tests/golden/make_golden.py
NARROW = dict(in_channels=4, out_channels=4, model_channels=32, attention_resolutions=(4, 2, 1), num_res_blocks=2,
              channel_mult=(1, 2, 4, 4), num_heads=2, context_dim=48)
WIDE = dict(in_channels=4, out_channels=4, model_channels=320, attention_resolutions=(1, 2), num_res_blocks=1,
            channel_mult=(1, 2), num_heads=8, context_dim=768)
VAE_SMALL = dict(ch=32, ch_mult=(1, 2, 4, 4), num_res_blocks=2, in_channels=3, out_ch=3, z_channels=4, embed_dim=4)


def golden(name):
    d = np.load(os.path.join(GOLDEN, name + '.npz'))
    return {k: torch.from_numpy(d[k]) if d[k].dtype.kind in 'fi' and d[k].ndim > 0 else d[k] for k in d.files}


def wsum(sd):
    s = a = 0.0
    for v in sd.values():
        s += float(v.double().sum())
        a += float(v.double().abs().sum())
    return np.asarray([s, a])


def maxdiff(a, b):
    return float((a.double() - b.double()).abs().max())
