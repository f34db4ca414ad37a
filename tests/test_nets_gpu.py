"""Network-level parity on the GPU: U-Net / VAE forwards of the engine against (a) the committed fixtures produced
by the real reference modules and (b) the CPU oracle on the same seeded inputs.

Tolerance: fp32 round-off through ~100 layers with a different summation order -- 2e-4 relative to the output's
max magnitude (measured values are printed; they are typically a few 1e-6)."""
import pytest
import torch

from cycle_diffusion_b200 import specs
from tests.common import NARROW, VAE_SMALL, WIDE, golden

pytestmark = pytest.mark.gpu
TOL = 2e-4


@pytest.fixture(scope='module')
def eng():
    from cycle_diffusion_b200.engine import Engine
    return Engine(0)


def relmax(a, b):
    return float((a.double() - b.double()).abs().max() / max(1.0, float(b.double().abs().max())))


@pytest.mark.parametrize('name,cfg', [('unet_sd_narrow', NARROW), ('unet_sd_wide', WIDE)])
def test_openai_unet_vs_reference_fixture(eng, name, cfg):
    from cycle_diffusion_b200.engine import UNet
    g = golden(name)
    sd = specs.synth_state_dict(specs.openai_unet_params(cfg), int(g['seed']))
    net = UNet(eng, cfg, 'openai').load_state_dict(sd)
    y = net(g['x'], g['t'], g['ctx']).cpu()
    r = relmax(y, g['y'])
    print(f'{name}: rel max err vs reference fixture {r:.3e}')
    assert r < TOL


def test_openai_unet_cfg_batch_and_prefix_loading(eng):
    """2B batch with distinct contexts (the CFG launch shape) and checkpoint-style key prefixes."""
    from cycle_diffusion_b200.engine import UNet
    from oracle import unet_openai
    sd = specs.synth_state_dict(specs.openai_unet_params(NARROW), 11)
    net = UNet(eng, NARROW, 'openai').load_state_dict({'model.diffusion_model.' + k: v for k, v in sd.items()},
                                                     prefix='model.diffusion_model.')
    g = torch.Generator().manual_seed(9)
    x = torch.randn(3, 4, 32, 32, generator=g)
    x = torch.cat([x, x])
    ctx = torch.randn(6, 77, 48, generator=g)
    t = torch.full((6,), 501.)
    y = net(x, t, ctx).cpu()
    with torch.no_grad():
        ref = unet_openai.unet_forward(sd, NARROW, x, t.long(), ctx)
    assert relmax(y, ref) < TOL


def test_iddpm_unet_vs_reference_fixture(eng):
    from cycle_diffusion_b200.engine import UNet
    g = golden('unet_iddpm64')
    cfg = specs.iddpm_config(64)
    sd = specs.synth_state_dict(specs.iddpm_unet_params(cfg), int(g['seed']))
    net = UNet(eng, cfg, 'iddpm').load_state_dict(sd)
    y = net(g['x'], g['t']).cpu()
    r = relmax(y, g['y'])
    print(f'iddpm64: rel max err vs reference fixture {r:.3e}')
    assert r < TOL


def test_vae_vs_reference_fixture(eng):
    from cycle_diffusion_b200.engine import VAE
    g = golden('vae_small')
    sd = specs.synth_state_dict(specs.kl_vae_params(VAE_SMALL), int(g['seed']))
    vae = VAE(eng, VAE_SMALL).load_state_dict(sd)
    m = vae.encode_moments(g['img']).cpu()
    r = vae.decode(g['z']).cpu()
    print(f'vae: moments {relmax(m, g["moments"]):.3e} rec {relmax(r, g["rec"]):.3e}')
    assert relmax(m, g['moments']) < TOL
    assert relmax(r, g['rec']) < TOL


def test_blob_adoption_roundtrip(eng):
    """The multi-GPU weight path on one device: copy rank-0's packed blob into a second net and adopt it."""
    from cycle_diffusion_b200.engine import UNet
    sd = specs.synth_state_dict(specs.openai_unet_params(NARROW), 11)
    a = UNet(eng, NARROW, 'openai').load_state_dict(sd)
    b = UNet(eng, NARROW, 'openai')
    b.blob_tensor().copy_(a.blob_tensor())
    b.adopt_blob()
    g = golden('unet_sd_narrow')
    assert torch.equal(a(g['x'], g['t'], g['ctx']), b(g['x'], g['t'], g['ctx']))


def test_error_paths(eng):
    from cycle_diffusion_b200.engine import UNet
    net = UNet(eng, NARROW, 'openai')
    with pytest.raises(AssertionError):          # forward before load/finalize
        net(torch.zeros(1, 4, 16, 16), torch.zeros(1), torch.zeros(1, 77, 48))
    sd = specs.synth_state_dict(specs.openai_unet_params(NARROW), 11)
    bad = dict(sd)
    bad.pop('out.2.bias')
    with pytest.raises(AssertionError):
        UNet(eng, NARROW, 'openai').load_state_dict(bad)
    net.load_state_dict(sd)
    with pytest.raises(AssertionError):          # spatial size not divisible by the U-Net's downsampling
        net(torch.zeros(1, 4, 12, 12), torch.zeros(1), torch.zeros(1, 77, 48))
