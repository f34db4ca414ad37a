"""Per-op parity on the GPU: every kernel family against a plain PyTorch fp32 CPU reference of the same op.

Tolerances are fp32 round-off budgets (different summation order), stated per test; the per-step scheduler
kernels are required to be BIT-EXACT against the reference formulas evaluated op-by-op with torch on the CPU.
"""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def eng():
    from cycle_diffusion_b200.engine import Engine
    return Engine(0)


def rel(a, b):
    return float((a.double() - b.double()).abs().max() / max(1e-30, float(b.double().abs().max())))


def nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


def nchw(x):
    return x.permute(0, 3, 1, 2).contiguous()


CONV_CASES = [
    # B, Cin, Cout, H, stride, pad_lo, up
    (2, 4, 32, 16, 1, 1, 1),      # SD conv_in shape class (K = 36)
    (1, 3, 32, 16, 1, 1, 1),      # pixel / VAE conv_in: Cin = 3 -> scalar gather path
    (2, 32, 4, 16, 1, 1, 1),      # out conv: N = 4
    (1, 64, 6, 8, 1, 1, 1),       # i-DDPM out conv: N = 6
    (2, 64, 64, 16, 2, 1, 1),     # OAI Downsample (stride 2, pad 1)
    (2, 32, 32, 16, 2, 0, 1),     # VAE Downsample: pad (0,1,0,1), stride 2
    (2, 64, 32, 8, 1, 1, 2),      # Upsample: nearest x2 folded into the gather
    (1, 320, 320, 32, 1, 1, 1),   # 128x128 tile path (M = 1024 ... few CTAs) 
    (4, 128, 256, 64, 1, 1, 1),   # big enough for the 128x128 tiles (M = 16384)
    (1, 20, 36, 7, 1, 1, 1),      # ragged everything
]


@pytest.mark.parametrize('B,Cin,Cout,H,stride,pad,up', CONV_CASES)
def test_conv3x3(eng, B, Cin, Cout, H, stride, pad, up):
    g = torch.Generator().manual_seed(Cin * 1000 + Cout)
    x = torch.randn(B, Cin, H, H, generator=g)
    w = torch.randn(Cout, Cin, 3, 3, generator=g) / math.sqrt(9 * Cin)
    b = torch.randn(Cout, generator=g)
    xin = F.interpolate(x, scale_factor=2, mode='nearest') if up == 2 else x
    if stride == 2 and pad == 0:
        ref = F.conv2d(F.pad(xin, (0, 1, 0, 1)), w, b, stride=2, padding=0)
    else:
        ref = F.conv2d(xin, w, b, stride=stride, padding=pad)
    y = nchw(eng.op_conv3x3(nhwc(x).cuda(), w.cuda(), b.cuda(), stride, pad, up).cpu())
    assert y.shape == ref.shape
    assert rel(y, ref) < 2e-5


@pytest.mark.parametrize('M,K,N', [(4, 1280, 1280), (308, 768, 640), (4096, 320, 2560), (1000, 77, 50), (64, 36, 4), (20000, 640, 640)])
def test_linear(eng, M, K, N):
    g = torch.Generator().manual_seed(M + K + N)
    x = torch.randn(M, K, generator=g)
    w = torch.randn(N, K, generator=g) / math.sqrt(K)
    b = torch.randn(N, generator=g)
    y = eng.op_linear(x.cuda(), w.cuda(), b.cuda()).cpu()
    assert rel(y, F.linear(x, w, b)) < 2e-5


@pytest.mark.parametrize('B,HW,C,eps,silu', [(2, 256, 320, 1e-5, True), (1, 64, 32, 1e-6, False), (3, 1024, 64, 1e-5, True),
                                              (2, 16, 1280, 1e-5, True), (1, 4096, 128, 1e-6, True)])
def test_groupnorm(eng, B, HW, C, eps, silu):
    g = torch.Generator().manual_seed(C)
    x = torch.randn(B, C, HW, generator=g) * 3 + 1.5
    gamma, beta = torch.randn(C, generator=g), torch.randn(C, generator=g)
    ref = F.group_norm(x, 32, gamma, beta, eps)
    if silu:
        ref = F.silu(ref)
    y = eng.op_groupnorm(x.permute(0, 2, 1).reshape(B, HW, 1, C).contiguous().cuda(), gamma.cuda(), beta.cuda(), eps, silu).cpu()
    y = y.reshape(B, HW, C).permute(0, 2, 1)
    assert float((y - ref).abs().max()) < 2e-5


@pytest.mark.parametrize('M,C', [(512, 320), (100, 1280), (7, 32)])
def test_layernorm(eng, M, C):
    g = torch.Generator().manual_seed(M)
    x = torch.randn(M, C, generator=g) * 2 + 0.3
    gamma, beta = torch.randn(C, generator=g), torch.randn(C, generator=g)
    y = eng.op_layernorm(x.cuda(), gamma.cuda(), beta.cuda()).cpu()
    assert float((y - F.layer_norm(x, (C,), gamma, beta)).abs().max()) < 2e-5


@pytest.mark.parametrize('B,Nq,Nk,heads,d', [(2, 256, 256, 8, 40), (1, 64, 77, 8, 160), (1, 200, 200, 1, 512), (2, 1024, 77, 8, 80),
                                               (1, 1024, 1024, 8, 80), (2, 64, 64, 6, 64), (1, 16, 16, 2, 16)])
def test_attention(eng, B, Nq, Nk, heads, d):
    g = torch.Generator().manual_seed(Nq + Nk)
    C = heads * d
    q, k, v = (torch.randn(B, n, C, generator=g) for n in (Nq, Nk, Nk))
    scale = d ** -0.5
    sp = lambda t: t.reshape(B, -1, heads, d).permute(0, 2, 1, 3)
    attn = (torch.einsum('bhid,bhjd->bhij', sp(q), sp(k)) * scale).softmax(-1)
    ref = torch.einsum('bhij,bhjd->bhid', attn, sp(v)).permute(0, 2, 1, 3).reshape(B, Nq, C)
    y = eng.op_attention(q.cuda(), k.cuda(), v.cuda(), heads, scale).cpu()
    assert float((y - ref).abs().max()) < 2e-5


def test_layout_roundtrip(eng):
    x = torch.randn(3, 37, 9, 11)
    y = eng.op_nchw_to_nhwc(x.cuda())
    assert torch.equal(y.cpu(), nhwc(x))
    assert torch.equal(eng.op_nhwc_to_nchw(y).cpu(), x)


# ------------------------------------------------------------------ scheduler kernels: bit-exact
def test_ddim_step_kernels_bit_exact(eng):
    from cycle_diffusion_b200.schedule import DDIMSchedule
    sch = DDIMSchedule(50, 0.1)
    g = torch.Generator().manual_seed(3)
    shape = (2, 4, 16, 16)
    x0, xt, nz, ec, eu = (torch.randn(shape, generator=g) for _ in range(5))
    for i in (0, 17, 49):
        c = sch.coef[i]
        f = lambda v: torch.full((1,), v)      # fp32 scalars, broadcast like the reference's [B,1,1,1] tensors
        sa, s1, s1t, sp, dc, sg = f(c.sqrt_at), f(c.sqrt_1m_at), f(c.sqrt_1m_at_tab), f(c.sqrt_aprev), f(c.dir_coef), f(c.sigma)
        # sample_xt_next, ddim.py:597-600
        e_t = (xt - sa * x0) / s1
        ref_next = sp * x0 + dc * e_t + sg * nz
        got = eng.ddim_posterior_sample(x0, xt, nz, c).cpu()
        assert torch.equal(got, ref_next)
        for scale, uc in ((1.0, None), (7.5, eu)):
            e = ec if uc is None else uc + scale * (ec - uc)
            pred_x0 = (xt - s1t * e) / sa
            ref_eps = (ref_next - sp * pred_x0 - dc * e) / sg / 1.0
            got = eng.ddim_compute_eps(xt, ref_next, ec, uc, scale, c).cpu()
            assert torch.equal(got, ref_eps)
            ref_prev = sp * pred_x0 + dc * e + sg * nz * 1.0
            got = eng.ddim_step_with_eps(xt, ec, uc, scale, nz, c).cpu()
            assert torch.equal(got, ref_prev)


def test_misc_elementwise_bit_exact(eng):
    g = torch.Generator().manual_seed(4)
    x = torch.rand(2, 3, 8, 8, generator=g)
    assert torch.equal(eng.shift_scale(x, -0.5, 2.0).cpu(), (x - 0.5) * 2.0)
    assert torch.equal(eng.shift_scale(x, 1.0, 0.5).cpu(), (x + 1.0) / 2.0)
    assert torch.equal(eng.affine(x, 1. / 0.18215, 0.0).cpu(), 1. / 0.18215 * x)
    nz = torch.randn(2, 3, 8, 8, generator=g)
    assert torch.equal(eng.q_sample(x, nz, 0.3, 0.9).cpu(), torch.full((1,), 0.3) * x + torch.full((1,), 0.9) * nz)
    mom = torch.randn(2, 8, 4, 4, generator=g) * 3
    n2 = torch.randn(2, 4, 4, 4, generator=g)
    mean, logvar = torch.chunk(mom, 2, dim=1)
    ref = 0.18215 * (mean + torch.exp(0.5 * torch.clamp(logvar, -30.0, 20.0)) * n2)
    assert float((eng.vae_posterior(mom, n2, 0.18215).cpu() - ref).abs().max()) < 1e-6   # expf vs torch.exp: 1 ulp
    assert torch.equal(eng.vae_posterior(mom, None, 0.18215).cpu(), 0.18215 * mean)


def test_pixel_step_kernels_bit_exact(eng):
    from cycle_diffusion_b200.schedule import PixelSchedule
    g = torch.Generator().manual_seed(5)
    shape = (2, 3, 8, 8)
    x0, xt, nz = (torch.randn(shape, generator=g) for _ in range(3))
    et6 = torch.randn(2, 6, 8, 8, generator=g)
    et = et6[:, :3]
    f = lambda v: torch.full((1,), v)
    sch = PixelSchedule('ddim', 10, 10, eta=0.1)
    for i in (0, 5):
        c = sch.coef[i]
        e0 = (xt - f(c.sqrt_at) * x0) / f(c.sqrt_1m_at)
        ref_next = f(c.sqrt_at_next) * x0 + f(c.c2) * e0 + f(c.c1) * nz
        assert torch.equal(eng.pixel_posterior_sample(x0, xt, nz, c).cpu(), ref_next)
        x0_t = (xt - et * f(c.sqrt_1m_at)) / f(c.sqrt_at)
        ref_eps = (ref_next - f(c.sqrt_at_next) * x0_t - f(c.c2) * et) / f(c.c1)
        assert torch.equal(eng.pixel_compute_eps(xt, ref_next, et6, c).cpu(), ref_eps)
        ref_step = f(c.sqrt_at_next) * x0_t + f(c.c2) * et + f(c.c1) * nz
        assert torch.equal(eng.pixel_step_with_eps(xt, et6, nz, c).cpu(), ref_step)
    sch = PixelSchedule('ddpm', 20, 6)
    c = sch.coef[1]
    mean = f(c.w0) * x0 + f(c.wt) * xt
    assert torch.equal(eng.pixel_posterior_sample(x0, xt, nz, c).cpu(), mean + f(c.post_std) * nz)
    m2 = f(c.inv_sqrt_1m_bt) * (xt - f(c.weight) * et)
    assert torch.equal(eng.pixel_compute_eps(xt, nz, et6, c).cpu(), (nz - m2) / f(c.std_model))
    assert torch.equal(eng.pixel_step_with_eps(xt, et6, nz, c).cpu(), m2 + f(c.mask) * f(c.std_model) * nz)
