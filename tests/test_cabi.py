"""CPU-only checks of the C ABI: the library loads, exports every symbol include/cdx.h declares, fails loudly
without a GPU, and its parameter inventories agree with cycle_diffusion_b200.specs (which make_golden.py pins
against the reference module trees with load_state_dict(strict=True))."""
import ctypes as C
import os
import re

import pytest
import torch

from cycle_diffusion_b200 import _cabi, specs
from cycle_diffusion_b200.engine import UNet, VAE
from tests.common import NARROW, VAE_SMALL, WIDE

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    txt = open(os.path.join(ROOT, 'include', 'cdx.h')).read()
    txt = re.sub(r'/\*.*?\*/', '', txt, flags=re.S)
    return sorted(set(re.findall(r'\b(cdx_[a-z0-9_]+)\s*\(', txt)))


def test_every_declared_symbol_is_exported_and_bound():
    syms = header_symbols()
    assert len(syms) >= 40
    for s in syms:
        assert hasattr(_cabi.lib, s), f'{s} declared in cdx.h but not exported by libcdx.so'
        assert s in _cabi.SIGNATURES, f'{s} has no ctypes signature'
    assert sorted(_cabi.SIGNATURES) == syms
    assert _cabi.lib.cdx_abi_version() == 2


@pytest.mark.skipif(torch.cuda.is_available(), reason='needs a box without CUDA')
def test_engine_creation_fails_loudly_without_gpu():
    h = C.c_void_p()
    rc = _cabi.lib.cdx_engine_create(0, C.byref(h))
    assert rc == -2 and b'no usable CUDA device' in _cabi.lib.cdx_last_error()
    from cycle_diffusion_b200.engine import Engine
    with pytest.raises(RuntimeError):
        Engine(0)


@pytest.mark.parametrize('cfg,kind,ref', [
    (specs.sd_unet_config(768), 'openai', specs.openai_unet_params), (specs.sd_unet_config(1280), 'openai', specs.openai_unet_params),
    (NARROW, 'openai', specs.openai_unet_params), (WIDE, 'openai', specs.openai_unet_params),
    (specs.iddpm_config(256), 'iddpm', specs.iddpm_unet_params), (specs.iddpm_config(64), 'iddpm', specs.iddpm_unet_params)])
def test_unet_inventory_matches_specs(cfg, kind, ref):
    net = UNet(None, cfg, kind)       # inventory-only (no engine, no GPU)
    assert net.inventory() == [(n, tuple(s)) for n, s, _ in ref(cfg)]


@pytest.mark.parametrize('cfg', [specs.kl_f8_config(), VAE_SMALL])
def test_vae_inventory_matches_specs(cfg):
    net = VAE(None, cfg)
    assert net.inventory() == [(n, tuple(s)) for n, s, _ in specs.kl_vae_params(cfg)]


def test_inventory_only_net_rejects_compute():
    net = UNet(None, NARROW, 'openai')
    with pytest.raises(AssertionError):
        net.finalize()


def test_clip_rank_tower_inventories_match_specs():
    """the CLIP image tower / projected text tower inventories of csrc/nets.cu equal specs.py (CPU, no GPU work)."""
    from cycle_diffusion_b200 import specs
    from cycle_diffusion_b200.engine import ClipVision, TextEncoder
    vc = specs.clip_b32_vision_config()
    assert [(a, tuple(b)) for a, b in ClipVision(None, vc).inventory()] == [(a, tuple(b)) for a, b, _ in specs.clip_vision_params(vc)]
    tc = specs.clip_b32_text_config()
    inv = TextEncoder(None, tc).inventory()
    assert [(a, tuple(b)) for a, b in inv[:-1]] == [(a, tuple(b)) for a, b, _ in specs.clip_text_params(tc)]
    assert inv[-1] == ('text_projection.weight', (512, 512))


def test_uncond_ldm_inventories_match_specs():
    """the context-free OPENAI U-Net (AttentionBlock) and the VQ-f4 first stage inventories equal specs.py."""
    from cycle_diffusion_b200 import specs
    from cycle_diffusion_b200.engine import UNet, VAE
    for cfg in (specs.ldm_uncond_unet_config(), dict(in_channels=3, out_channels=3, model_channels=32, attention_resolutions=(2, 4), num_res_blocks=1,
                                                     channel_mult=(1, 2, 2), num_head_channels=16, context_dim=0)):
        assert [(a, tuple(b)) for a, b in UNet(None, cfg, 'openai').inventory()] == [(a, tuple(b)) for a, b, _ in specs.openai_unet_params(cfg)]
    vc = specs.vq_f4_config()
    assert [(a, tuple(b)) for a, b in VAE(None, vc).inventory()] == [(a, tuple(b)) for a, b, _ in specs.kl_vae_params(vc)]


def test_ddpm_unet_inventory_matches_specs():
    """Ho-et-al DDPM U-Net inventory (csrc/nets.cu build_ddpm_inventory) equals specs.ddpm_unet_params."""
    from cycle_diffusion_b200 import specs
    from cycle_diffusion_b200.engine import UNet
    for cfg in (specs.ddpm_config(256), dict(image_size=32, in_channels=3, out_channels=3, model_channels=32, num_res_blocks=2, channel_mult=(1, 2, 2),
                                             attention_resolutions=(2,))):
        assert [(a, tuple(b)) for a, b in UNet(None, cfg, 'ddpm').inventory()] == [(a, tuple(b)) for a, b, _ in specs.ddpm_unet_params(cfg)]
