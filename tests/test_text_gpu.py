"""CLIP text tower through the C ABI against the transformers-generated fixture and the oracle."""
import pytest
import torch

from cycle_diffusion_b200 import specs
from tests.common import golden, maxdiff

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def eng():
    from cycle_diffusion_b200.engine import Engine
    return Engine(0)


def _cfg(g, tag):
    return dict(zip(('vocab_size', 'width', 'layers', 'heads', 'max_len', 'mlp_width'), (int(v) for v in g[f'cfg_{tag}'])))


@pytest.mark.parametrize('mode', [0, 1])
@pytest.mark.parametrize('tag', ['small', 'wide'])
def test_text_encoder_vs_transformers_fixture(eng, tag, mode):
    from cycle_diffusion_b200.engine import TextEncoder
    g = golden('clip_text')
    cfg = _cfg(g, tag)
    eng.set_mma_mode(mode)
    try:
        sd = specs.synth_state_dict(specs.clip_text_params(cfg), 77 + cfg['width'], gain=2.0)
        enc = TextEncoder(eng, cfg)
        assert [n for n, _ in enc.inventory()] == [n for n, _, _ in specs.clip_text_params(cfg)]
        enc.load_state_dict(sd)
        y = enc(g[f'ids_{tag}']).cpu()
        ys = enc(g[f'ids_short_{tag}']).cpu()
        print(f'text[{tag}, mma {mode}]: |dy| {maxdiff(y, g[f"out_{tag}"]):.2e}  short {maxdiff(ys, g[f"out_short_{tag}"]):.2e}')
        assert maxdiff(y, g[f'out_{tag}']) < 5e-5
        assert maxdiff(ys, g[f'out_short_{tag}']) < 5e-5
        with pytest.raises(AssertionError):
            enc(torch.zeros(1, 78, dtype=torch.long))           # longer than the position table
    finally:
        eng.set_mma_mode(1)


def test_text_encoder_full_size_vs_oracle(eng):
    """The real ViT-L/14 text-tower shape (12 x 768, 12 heads, vocab 49408) against the CPU oracle on synthetic weights."""
    from cycle_diffusion_b200.engine import TextEncoder
    from oracle import clip_text
    cfg = specs.clip_text_config()
    sd = specs.synth_state_dict(specs.clip_text_params(cfg), 4242, gain=2.0)
    enc = TextEncoder(eng, cfg).load_state_dict(sd)
    g = torch.Generator().manual_seed(3)
    ids = torch.randint(0, cfg['vocab_size'], (4, 77), generator=g)
    ids[:, 0] = 49406
    y = enc(ids).cpu()
    with torch.no_grad():
        ref = clip_text.text_forward(sd, cfg, ids)
    print(f'text[ViT-L/14 shape]: |dy| {maxdiff(y, ref):.2e}  |y|max {float(ref.abs().max()):.2f}')
    assert maxdiff(y, ref) < 1e-4


@pytest.mark.parametrize('tag', ['small', 'wide'])
def test_bert_text_encoder_vs_reference_fixture(eng, tag):
    """LDM BERTEmbedder transformer (x_transformer) through the C ABI against the fixture from the reference's own module."""
    from cycle_diffusion_b200.engine import TextEncoder
    g = golden('bert_text')
    keys = ('vocab_size', 'width', 'layers', 'heads', 'dim_head', 'max_len', 'mlp_width')
    cfg = dict(zip(keys, (int(v) for v in g[f'cfg_{tag}'])), kind='xtransformer')
    sd = specs.synth_state_dict(specs.bert_text_params(cfg), 11 + cfg['width'], gain=2.0)
    enc = TextEncoder(eng, cfg)
    assert [n for n, _ in enc.inventory()] == [n for n, _, _ in specs.bert_text_params(cfg)]
    enc.load_state_dict(sd)
    y = enc(g[f'tok_{tag}']).cpu()
    print(f'bert_text[{tag}]: |dy| {maxdiff(y, g[f"out_{tag}"]):.2e}')
    assert maxdiff(y, g[f'out_{tag}']) < 5e-5
