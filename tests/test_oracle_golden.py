"""The oracle restatement against fixtures produced by the real reference modules (tests/golden/make_golden.py).

CPU only.  Tolerances: the oracle calls the same ATen CPU kernels as the reference modules, in the same
order, so single forwards agree to fp32 round-off (<= 2e-5 abs on O(1) outputs); recovered noise ``z`` is
amplified by 1/sigma_t, so it is compared relative to its own magnitude.
"""
import numpy as np
import pytest
import torch

from cycle_diffusion_b200 import specs
from oracle import dpm_encoder, schedules, unet_iddpm, unet_openai, vae_kl
from tests.common import NARROW, VAE_SMALL, WIDE, golden, maxdiff, wsum

torch.set_num_threads(8)


def test_schedule_tables_bit_exact():
    g = golden('schedule_ldm')
    ac = schedules.ldm_alphas_cumprod()
    assert torch.equal(ac, g['alphas_cumprod'])
    for S in (10, 50, 99, 100):
        tab = schedules.DDIMTables(S, 0.1)
        assert np.array_equal(tab.timesteps, g[f'ts_{S}'].numpy())
        f32 = lambda t: torch.stack([torch.full((1,), t[i]) for i in range(S)]).flatten()
        assert torch.equal(f32(tab.alphas), g[f'a_{S}'])
        assert torch.equal(f32(tab.alphas_prev), g[f'aprev_{S}'])
        assert torch.equal(f32(tab.sigmas), g[f'sigma_{S}'])
        assert torch.equal(f32(tab.sqrt_one_minus_alphas), g[f'sqrt1ma_{S}'])
    assert list(schedules.ddim_timesteps(10)) == [1, 101, 201, 301, 401, 501, 601, 701, 801, 901]


@pytest.mark.parametrize('name,cfg', [('unet_sd_narrow', NARROW), ('unet_sd_wide', WIDE)])
def test_openai_unet(name, cfg):
    g = golden(name)
    sd = specs.synth_state_dict(specs.openai_unet_params(cfg), int(g['seed']))
    assert np.allclose(wsum(sd), g['wsum'], rtol=1e-12), 'synthetic weight generator drifted'
    with torch.no_grad():
        y = unet_openai.unet_forward(sd, cfg, g['x'], g['t'], g['ctx'])
    assert maxdiff(y, g['y']) <= 2e-5 * max(1.0, float(g['y'].abs().max()))


def test_vae():
    g = golden('vae_small')
    sd = specs.synth_state_dict(specs.kl_vae_params(VAE_SMALL), int(g['seed']))
    assert np.allclose(wsum(sd), g['wsum'], rtol=1e-12)
    with torch.no_grad():
        m = vae_kl.encode_moments(sd, VAE_SMALL, g['img'])
        r = vae_kl.decode(sd, VAE_SMALL, g['z'])
    assert maxdiff(m, g['moments']) <= 2e-5 * max(1.0, float(g['moments'].abs().max()))
    assert maxdiff(r, g['rec']) <= 2e-5 * max(1.0, float(g['rec'].abs().max()))


def test_iddpm_unet():
    g = golden('unet_iddpm64')
    cfg = specs.iddpm_config(64)
    sd = specs.synth_state_dict(specs.iddpm_unet_params(cfg), int(g['seed']))
    assert np.allclose(wsum(sd), g['wsum'], rtol=1e-12)
    with torch.no_grad():
        y = unet_iddpm.unet_forward(sd, cfg, g['x'], g['t'])
    assert maxdiff(y, g['y']) <= 2e-5 * max(1.0, float(g['y'].abs().max()))


@pytest.mark.parametrize('tag', ['a', 'b'])
def test_ddim_cycle(tag):
    g = golden('ddim_cycle_narrow')
    S, skip, wb, enc_scale, dec_scale, seed = [float(v) for v in g[f'cfg_{tag}']]
    S, skip, wb, seed = int(S), int(skip), int(wb), int(seed)
    sd = specs.synth_state_dict(specs.openai_unet_params(NARROW), 11)
    unet = lambda x, t, c: unet_openai.unet_forward(sd, NARROW, x, t, c)
    B = g['x0'].shape[0]
    torch.manual_seed(seed)
    with torch.no_grad():
        z_list = dpm_encoder.latent_encode(unet, g['x0'], g['c_src'], g['uc'], S, 0.1, skip, wb, enc_scale)
        z = torch.stack(z_list, dim=1).view(B, -1)
        zref = g[f'z_{tag}']
        assert z.shape == zref.shape
        assert maxdiff(z, zref) <= 1e-4 * float(zref.abs().max())
        eps_list = zref.view(B, wb - skip, 4, 16, 16)
        same = dpm_encoder.latent_decode(unet, eps_list[:, 0], eps_list[:, 1:], g['c_src'], g['uc'], S, 0.1, skip, enc_scale)
        tgt = dpm_encoder.latent_decode(unet, eps_list[:, 0], eps_list[:, 1:], g['c_tgt'], g['uc'], S, 0.1, skip, dec_scale)
    assert maxdiff(same, g[f'same_{tag}']) <= 1e-4
    assert maxdiff(tgt, g[f'tgt_{tag}']) <= 1e-4
    # built-in known-answer property: same-condition cycle reconstructs x0
    assert maxdiff(same, g['x0']) <= 1e-4


@pytest.mark.parametrize('tag,kw', [('ddim', dict(sample_type='ddim', eta=0.1, custom_steps=10, es_steps=10)),
                                    ('ddpm', dict(sample_type='ddpm', eta=None, custom_steps=20, es_steps=6))])
def test_pixel_cycle(tag, kw):
    g = golden('pixel_cycle_iddpm64')
    cfg = specs.iddpm_config(64)
    sd = specs.synth_state_dict(specs.iddpm_unet_params(cfg), 31)
    model = lambda x, t: unet_iddpm.unet_forward(sd, cfg, x, t)
    cyc = dpm_encoder.PixelCycle(model, resolution=64, **kw)
    torch.manual_seed(2000)
    with torch.no_grad():
        z = cyc.encode(g['image'])
        zref = g[f'z_{tag}']
        assert maxdiff(z, zref) <= 2e-4 * float(zref.abs().max())
        img = cyc.forward(zref)
    assert maxdiff(img, g[f'img_{tag}']) <= 1e-3


@pytest.mark.parametrize('tag', ['small', 'wide'])
def test_clip_text_oracle_vs_transformers_fixture(tag):
    """oracle/clip_text.py against the installed transformers CLIPTextModel (the third-party model behind FrozenCLIPEmbedder)."""
    from oracle import clip_text
    g = golden('clip_text')
    cfg = dict(zip(('vocab_size', 'width', 'layers', 'heads', 'max_len', 'mlp_width'), (int(v) for v in g[f'cfg_{tag}'])))
    sd = specs.synth_state_dict(specs.clip_text_params(cfg), 77 + cfg['width'], gain=2.0)
    with torch.no_grad():
        y = clip_text.text_forward(sd, cfg, g[f'ids_{tag}'])
        ys = clip_text.text_forward(sd, cfg, g[f'ids_short_{tag}'])
    assert maxdiff(y, g[f'out_{tag}']) <= 2e-5
    assert maxdiff(ys, g[f'out_short_{tag}']) <= 2e-5
    # causality: a token's output must not depend on later tokens
    ids2 = g[f'ids_{tag}'].clone()
    ids2[:, 40:] = (ids2[:, 40:] + 1) % cfg['vocab_size']
    with torch.no_grad():
        y2 = clip_text.text_forward(sd, cfg, ids2)
    assert torch.equal(y2[:, :40], y[:, :40]) and not torch.equal(y2[:, 40:], y[:, 40:])


@pytest.mark.parametrize('tag', ['small', 'wide'])
def test_bert_text_oracle_vs_reference_fixture(tag):
    """oracle/bert_text.py against the reference's own x_transformer TransformerWrapper (the LDM BERTEmbedder)."""
    from oracle import bert_text
    g = golden('bert_text')
    keys = ('vocab_size', 'width', 'layers', 'heads', 'dim_head', 'max_len', 'mlp_width')
    cfg = dict(zip(keys, (int(v) for v in g[f'cfg_{tag}'])), kind='xtransformer')
    sd = specs.synth_state_dict(specs.bert_text_params(cfg), 11 + cfg['width'], gain=2.0)
    with torch.no_grad():
        y = bert_text.text_forward(sd, cfg, g[f'tok_{tag}'])
    assert maxdiff(y, g[f'out_{tag}']) <= 2e-5


def test_clip_rank_oracle_vs_third_party_and_reference_metrics():
    """oracle.clip_rank against the installed transformers CLIPModel (features, scores) and against the reference's own
    evaluation/utils.py (PSNR / SSIM / L2), fixture tests/golden/clip_rank.npz."""
    from cycle_diffusion_b200 import specs
    from oracle import clip_rank
    g = golden('clip_rank')
    vc = dict(kind='clip_vision', width=64, layers=2, heads=4, mlp_width=256, patch=8, image_size=32, proj_dim=48)
    tc = dict(kind='clip', vocab_size=600, width=96, layers=2, heads=4, max_len=77, mlp_width=384, proj_dim=48)
    sd = dict(specs.synth_state_dict(specs.clip_vision_params(vc), 31, gain=2.0))
    sd.update(specs.synth_state_dict(specs.clip_text_params(tc) + [('text_projection.weight', (48, tc['width']), 'w')], 32, gain=2.0))
    with torch.no_grad():
        pre = clip_rank.preprocess(g['img'], 32)
        assert maxdiff(pre, g['pre_img']) < 1e-5
        assert maxdiff(clip_rank.image_features(sd, vc, pre), g['f_img']) < 2e-5 * float(g['f_img'].abs().max()) + 1e-6
        assert maxdiff(clip_rank.text_features(sd, tc, g['ids_e'].long()), g['f_enc']) < 2e-5 * float(g['f_enc'].abs().max()) + 1e-6
        clip, dclip = clip_rank.directional_clip(sd, vc, tc, g['img'], g['orig'], g['ids_e'].long(), g['ids_d'].long())
    assert maxdiff(clip, g['clip']) < 1e-5 and maxdiff(dclip, g['dclip']) < 1e-4
    for i in range(2):
        m = clip_rank.metrics(g['met_a'][i], g['met_b'][i])
        ref = g['met'][i].tolist()
        assert abs(m[0] - ref[0]) < 1e-4 and abs(m[1] - ref[1]) < 1e-9 and abs(m[2] - ref[2]) < 1e-4, (m, ref)


UNCOND_SMALL = dict(in_channels=3, out_channels=3, model_channels=32, attention_resolutions=(2, 4), num_res_blocks=1,
                    channel_mult=(1, 2, 2), num_head_channels=16, context_dim=0)
VQ_SMALL = dict(ch=32, ch_mult=(1, 2, 4), num_res_blocks=1, in_channels=3, out_ch=3, z_channels=3, embed_dim=3, vq=True, n_embed=256)


def test_ldm_uncond_oracle_vs_reference_fixture():
    """unconditional LDM pieces (AttentionBlock U-Net, VQ first stage, encode -> decode -> eta-1 refine) of the oracle
    against the fixture produced by the reference's own UNetModel / Encoder / Decoder / DDIMSampler."""
    g = golden('ldm_uncond')
    usd = specs.synth_state_dict(specs.openai_unet_params(UNCOND_SMALL), 41)
    vsd = specs.synth_state_dict(specs.kl_vae_params(VQ_SMALL), 42)
    assert abs(wsum(usd)[1] - float(g['wsum'][1])) < 1e-6 * float(g['wsum'][1])
    fn = lambda x, t, c: unet_openai.unet_forward(usd, UNCOND_SMALL, x, t, None)
    with torch.no_grad():
        assert maxdiff(fn(g['x'], g['t'].long(), None), g['y']) < 2e-5 * float(g['y'].abs().max())
        assert maxdiff(vae_kl.encode_moments(vsd, VQ_SMALL, g['img']), g['h']) < 2e-5 * float(g['h'].abs().max())
        assert maxdiff(vae_kl.decode(vsd, VQ_SMALL, g['zz']), g['rec']) < 2e-5 * float(g['rec'].abs().max())
        S, wb, r, seed = [int(v) for v in g['cyc']]
        torch.manual_seed(seed)
        z = torch.stack(dpm_encoder.latent_encode(fn, g['x0'], None, None, S, 0.1, 0, wb, 1.0), dim=1)
        dec = dpm_encoder.latent_decode(fn, z[:, 0], z[:, 1:], None, None, S, 0.1, 0, 1.0)
        ref = dpm_encoder.latent_refine(fn, dec, None, None, S, r)
    assert maxdiff(z.reshape(2, -1), g['z']) < 2e-4 * float(g['z'].abs().max())
    assert maxdiff(dec, g['dec']) < 1e-4 and maxdiff(ref, g['refined']) < 1e-4


DDPM_SMALL = dict(image_size=32, in_channels=3, out_channels=3, model_channels=32, num_res_blocks=2, channel_mult=(1, 2, 2), attention_resolutions=(2,))


def test_unet_ddpm_oracle_vs_reference_fixture():
    """oracle.unet_ddpm against the reference's own Ho-et-al DDPM class (tests/golden/unet_ddpm.npz)."""
    from oracle import unet_ddpm
    g = golden('unet_ddpm')
    sd = specs.synth_state_dict(specs.ddpm_unet_params(DDPM_SMALL), 61)
    with torch.no_grad():
        y = unet_ddpm.unet_forward(sd, DDPM_SMALL, g['x'], g['t'])
    assert maxdiff(y, g['y']) < 2e-5 * float(g['y'].abs().max())
