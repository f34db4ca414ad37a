"""Ho-et-al DDPM pixel U-Net (CDX_UNET_DDPM) against the reference class's own output, and the pixel wrapper on it."""
import pytest
import torch

from cycle_diffusion_b200 import specs
from tests.common import golden, maxdiff

pytestmark = pytest.mark.gpu

DDPM_SMALL = dict(image_size=32, in_channels=3, out_channels=3, model_channels=32, num_res_blocks=2, channel_mult=(1, 2, 2), attention_resolutions=(2,))


@pytest.fixture(scope='module')
def eng():
    from cycle_diffusion_b200.engine import Engine
    return Engine(0)


@pytest.mark.parametrize('mode', [1, 0], ids=['tcgen05', 'ffma'])
def test_forward_vs_reference_fixture(eng, mode):
    from cycle_diffusion_b200.engine import UNet
    g = golden('unet_ddpm')
    eng.set_mma_mode(mode)
    try:
        net = UNet(eng, DDPM_SMALL, 'ddpm').load_state_dict(specs.synth_state_dict(specs.ddpm_unet_params(DDPM_SMALL), 61))
        y = net(g['x'], g['t']).cpu()
    finally:
        eng.set_mma_mode(1)
    r = maxdiff(y, g['y']) / float(g['y'].abs().max())
    print(f'unet_ddpm[mode {mode}]: rel {r:.2e}')
    assert r < 2e-4


def test_pixel_wrapper_on_the_ddpm_family_vs_oracle(eng):
    """DDPMDDIMWrapper with a CelebA-HQ / LSUN style model (DW:360-369): encode -> forward against the CPU oracle loops on oracle.unet_ddpm."""
    from cycle_diffusion_b200.engine import UNet
    from cycle_diffusion_b200.wrappers import DDPMDDIMWrapper
    from oracle import dpm_encoder, unet_ddpm
    sd = specs.synth_state_dict(specs.ddpm_unet_params(DDPM_SMALL), 61)
    net = UNet(eng, DDPM_SMALL, 'ddpm').load_state_dict(sd)
    w = DDPMDDIMWrapper('celeba_hq_32', 'ddim', custom_steps=10, es_steps=10, eta=0.1, unet=net, image_size=32)
    assert w.model_family == 'ddpm'
    img = torch.rand(1, 3, 32, 32, generator=torch.Generator().manual_seed(2))
    torch.manual_seed(11)
    z = w.encode(img)
    out = w(z).cpu()
    fn = lambda x, t: unet_ddpm.unet_forward(sd, DDPM_SMALL, x, t)
    ora = dpm_encoder.PixelCycle(fn, sample_type='ddim', custom_steps=10, es_steps=10, eta=0.1, resolution=32)
    torch.manual_seed(11)
    with torch.no_grad():
        z_ref = ora.encode(img)
        out_ref = ora.forward(z_ref)
    rz = maxdiff(z.cpu(), z_ref) / float(z_ref.abs().max())
    print(f'ddpm-family pixel wrapper: rel|dz| {rz:.2e}  |d img| {maxdiff(out, out_ref):.2e}  cycle |img - rec| {maxdiff(out, img):.2e}')
    assert rz < 2e-4 and maxdiff(out, out_ref) < 1e-3
