"""Host-side schedule tables of the product against the reference-generated fixture (bit-exact) -- CPU only."""
import ctypes

import numpy as np
import torch

from cycle_diffusion_b200.schedule import DDIMSchedule, PixelSchedule, ldm_alphas_cumprod
from tests.common import golden


def test_ddim_tables_bit_exact():
    g = golden('schedule_ldm')
    assert torch.equal(ldm_alphas_cumprod(), g['alphas_cumprod'])
    for S in (10, 50, 99, 100):
        sch = DDIMSchedule(S, 0.1)
        assert np.array_equal(sch.timesteps, g[f'ts_{S}'].numpy())
        a, ap, sg, s1 = g[f'a_{S}'], g[f'aprev_{S}'], g[f'sigma_{S}'], g[f'sqrt1ma_{S}']
        for i, c in enumerate(sch.coef):
            idx = S - 1 - i
            f = lambda v: ctypes.c_float(v).value
            assert f(c.sqrt_at) == a[idx].sqrt().item()
            assert f(c.sqrt_1m_at) == (1 - a[idx]).sqrt().item()
            assert f(c.sqrt_1m_at_tab) == s1[idx].item()
            assert f(c.sqrt_aprev) == ap[idx].sqrt().item()
            assert f(c.sigma) == sg[idx].item()
            assert f(c.dir_coef) == (1. - ap[idx] - sg[idx] ** 2).sqrt().item()
        assert sch.t_loop == [float(t) for t in np.flip(sch.timesteps)]


def test_ddim_skip_steps_geometry():
    sch = DDIMSchedule(10, 0.1, skip_steps=3)
    assert sch.refine_steps == 7 and len(sch.coef) == 7
    assert sch.t_loop == [601., 501., 401., 301., 201., 101., 1.]
    full = DDIMSchedule(10, 0.1)
    assert ctypes.c_float(sch.coef[0].sigma).value == ctypes.c_float(full.coef[3].sigma).value


def test_pixel_schedule_geometry():
    s = PixelSchedule('ddim', 10, 10, eta=0.1)
    assert [p[0] for p in s.pairs] == [900, 800, 700, 600, 500, 400, 300, 200, 100, 0]
    assert s.pairs[-1] == (0, -1)
    # last step: at_next = 1 -> c1 = c2 = 0 (DW:196-199, 217-218)
    assert s.coef[-1].c1 == 0.0 and s.coef[-1].c2 == 0.0 and s.coef[-1].sqrt_at_next == 1.0
    d = PixelSchedule('ddpm', 20, 6)
    assert d.coef[-1].mask == 0.0 and d.coef[0].mask == 1.0
