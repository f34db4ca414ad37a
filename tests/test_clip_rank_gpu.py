"""Directional-CLIP ranking and the text-task metrics in the engine (cdx_clip_preprocess / cdx_clip_image_features /
cdx_text_features / cdx_dclip_scores / cdx_image_metrics) against tests/golden/clip_rank.npz -- features and scores from the installed
transformers CLIPModel, metrics from the reference's own evaluation/utils.py."""
import pytest
import torch

from cycle_diffusion_b200 import specs
from tests.common import golden, maxdiff

pytestmark = pytest.mark.gpu

VC = dict(kind='clip_vision', width=64, layers=2, heads=4, mlp_width=256, patch=8, image_size=32, proj_dim=48)
TC = dict(kind='clip', vocab_size=600, width=96, layers=2, heads=4, max_len=77, mlp_width=384, proj_dim=48)


def _sd():
    sd = dict(specs.synth_state_dict(specs.clip_vision_params(VC), 31, gain=2.0))
    sd.update(specs.synth_state_dict(specs.clip_text_params(TC) + [('text_projection.weight', (48, TC['width']), 'w')], 32, gain=2.0))
    return sd


@pytest.fixture(scope='module')
def eng():
    from cycle_diffusion_b200.engine import Engine
    return Engine(0)


@pytest.fixture(scope='module')
def dclip(eng):
    from cycle_diffusion_b200.clip_rank import DirectionalCLIP
    g = golden('clip_rank')
    table = {'enc': g['ids_e'].long(), 'dec': g['ids_d'].long()}
    tok = lambda texts: torch.stack([table[t.split(':')[0]][int(t.split(':')[1])] for t in texts])      # 'enc:1' -> row 1 of the fixture ids
    return DirectionalCLIP(eng, _sd(), tok, vision_cfg=VC, text_cfg=TC)


def test_preprocess_features_and_scores_vs_fixture(eng, dclip):
    g = golden('clip_rank')
    pre = eng.clip_preprocess(g['img'], 32).cpu()
    f_img = dclip.encode_image(g['img']).cpu()
    f_enc = dclip.encode_text([f'enc:{i}' for i in range(3)]).cpu()
    clip, dc = dclip(g['img'], g['orig'], [f'enc:{i}' for i in range(3)], [f'dec:{i}' for i in range(3)])
    r = lambda a, b: maxdiff(a, b) / float(b.abs().max())
    print(f'clip rank: preprocess |d| {maxdiff(pre, g["pre_img"]):.2e}  image feat rel {r(f_img, g["f_img"]):.2e}  text feat rel {r(f_enc, g["f_enc"]):.2e}'
          f'  clip |d| {maxdiff(clip.cpu(), g["clip"]):.2e}  dclip |d| {maxdiff(dc.cpu(), g["dclip"]):.2e}')
    assert maxdiff(pre, g['pre_img']) < 2e-5
    assert r(f_img, g['f_img']) < 5e-5 and r(f_enc, g['f_enc']) < 5e-5
    assert maxdiff(clip.cpu(), g['clip']) < 1e-4 and maxdiff(dc.cpu(), g['dclip']) < 1e-3


def test_metrics_vs_reference_functions(eng):
    from cycle_diffusion_b200.clip_rank import translate_text_metrics
    g = golden('clip_rank')
    m = translate_text_metrics(eng, g['met_a'], g['met_b'])
    ref = torch.as_tensor(g['met'])
    print(f'metrics: psnr {m["psnr"].tolist()} ssim {m["ssim"].tolist()} l2 {m["l2"].tolist()}  (reference {ref.tolist()})')
    assert maxdiff(m['psnr'].cpu().double(), ref[:, 0]) < 1e-3
    assert maxdiff(m['ssim'].cpu().double(), ref[:, 1]) < 1e-6
    assert maxdiff(m['l2'].cpu().double(), ref[:, 2]) < 1e-3
    same = translate_text_metrics(eng, g['met_a'], g['met_a'])
    assert same['psnr'].tolist() == [100.0, 100.0] and maxdiff(same['ssim'].cpu(), torch.ones(2)) < 1e-6


def test_openai_state_dict_layout_and_device_ranking(eng, dclip):
    """The OpenAI-clip key layout (fused in_proj, x @ proj) loads to the same towers; rank() picks per sample on the device."""
    from cycle_diffusion_b200.clip_rank import DirectionalCLIP
    sd = _sd()
    oa = {}
    for tower, src, dst in (('vision', 'vision_model.encoder.layers', 'visual.transformer.resblocks'), ('text', 'text_model.encoder.layers', 'transformer.resblocks')):
        for l in range(2):
            p = f'{src}.{l}'
            oa[f'{dst}.{l}.attn.in_proj_weight'] = torch.cat([sd[f'{p}.self_attn.{n}.weight'] for n in ('q_proj', 'k_proj', 'v_proj')])
            oa[f'{dst}.{l}.attn.in_proj_bias'] = torch.cat([sd[f'{p}.self_attn.{n}.bias'] for n in ('q_proj', 'k_proj', 'v_proj')])
            for a, b in (('out_proj', 'attn.out_proj'), ('layer_norm1', 'ln_1'), ('layer_norm2', 'ln_2'), ('mlp.fc1', 'mlp.c_fc'), ('mlp.fc2', 'mlp.c_proj')):
                src_a = f'{p}.self_attn.{a}' if a == 'out_proj' else f'{p}.{a}'
                oa[f'{dst}.{l}.{b}.weight'], oa[f'{dst}.{l}.{b}.bias'] = sd[src_a + '.weight'], sd[src_a + '.bias']
    oa.update({'visual.class_embedding': sd['vision_model.embeddings.class_embedding'], 'visual.conv1.weight': sd['vision_model.embeddings.patch_embedding.weight'],
               'visual.positional_embedding': sd['vision_model.embeddings.position_embedding.weight'],
               'visual.ln_pre.weight': sd['vision_model.pre_layrnorm.weight'], 'visual.ln_pre.bias': sd['vision_model.pre_layrnorm.bias'],
               'visual.ln_post.weight': sd['vision_model.post_layernorm.weight'], 'visual.ln_post.bias': sd['vision_model.post_layernorm.bias'],
               'visual.proj': sd['visual_projection.weight'].t().contiguous(), 'text_projection': sd['text_projection.weight'].t().contiguous(),
               'token_embedding.weight': sd['text_model.embeddings.token_embedding.weight'], 'positional_embedding': sd['text_model.embeddings.position_embedding.weight'],
               'ln_final.weight': sd['text_model.final_layer_norm.weight'], 'ln_final.bias': sd['text_model.final_layer_norm.bias']})
    d2 = DirectionalCLIP(eng, oa, dclip.tokenizer, vision_cfg=VC, text_cfg=TC)
    g = golden('clip_rank')
    e_t, d_t = [f'enc:{i}' for i in range(3)], [f'dec:{i}' for i in range(3)]
    a = dclip(g['img'], g['orig'], e_t, d_t)
    b = d2(g['img'], g['orig'], e_t, d_t)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])
    cands = [g['img'].to(eng.device), g['orig'].to(eng.device).flip(0), (0.5 * g['img'] + 0.5 * g['orig']).to(eng.device)]
    best_img, best, scores = dclip.rank(cands, g['orig'], e_t, d_t)
    assert best_img.is_cuda and scores.shape == (3, 3)
    ref_scores = torch.stack([dclip(c, g['orig'], e_t, d_t)[1] for c in cands], dim=1)
    assert torch.equal(best, ref_scores.argmax(1))
    for bi in range(3):
        assert torch.equal(best_img[bi], cands[best[bi].item()][bi])
