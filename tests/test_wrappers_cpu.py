"""CPU-only checks of the drop-in wrapper surface: default checkpoint locations (the same ``[gan]`` INI section must
resolve to the same files as the reference) and the conditioning-model policy (no silent noise tokens with real weights)."""
import os

import pytest

from cycle_diffusion_b200 import wrappers


def test_default_checkpoint_paths_match_reference():
    # stable_diffusion_stochastic_text_wrapper.py:21-23
    assert wrappers.SDStochasticTextWrapper.default_checkpoint('sd-v1-4.ckpt') == os.path.join('ckpts', 'stable_diffusion', 'sd-v1-4.ckpt')
    # latentdiff_stochastic_text_wrapper.py:20-23
    assert wrappers.LatentDiffStochasticTextWrapper.default_checkpoint('text2img-large') == \
        os.path.join('ckpts', 'ldm_models', 'text2img-large', 'model.ckpt')


def test_cond_prefixes():
    assert wrappers.SDStochasticTextWrapper.COND_PREFIX == 'cond_stage_model.transformer.'
    assert wrappers.LatentDiffStochasticTextWrapper.COND_PREFIX == 'cond_stage_model.'
    assert wrappers.SDStochasticTextWrapper.COND_CLASS is wrappers.ClipTextCondStage
    assert wrappers.LatentDiffStochasticTextWrapper.COND_CLASS is wrappers.BertTextCondStage


def test_get_gan_wrapper_rejects_unknown_type():
    with pytest.raises(ValueError):
        wrappers.get_gan_wrapper(dict(gan_type='StyleGAN2', source_model_type='x'))
