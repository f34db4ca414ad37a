"""Diffusers-style surface: ``CycleDiffusionPipeline.__call__`` delegating to the same C-ABI loop drivers.

The Diffusers pipeline is NOT part of /root/reference and diffusers is not installed here, so parity at this surface is
*unpinned* (SURVEY.md 8b); the argument list follows diffusers <= 0.2x from the survey.  Semantics are mapped onto the
reference's sampler: ``strength`` -> ``skip_steps = S - int(S * strength)`` (ddim.py:470), DDIMScheduler ``steps_offset=1``
== the ``+1`` of util.py:58, ``posterior_sample`` == sample_xt_next (ddim.py:582-601), ``compute_noise`` == compute_eps
(ddim.py:575-579).  The loop is the engine's lock-step driver (``cdx_cycle_lockstep``): the source chain (source prompt,
source_guidance_scale) and the target chain (prompt, guidance_scale) advance together, one U-Net call per step on the batch
[source segments | target segments] and one fused elementwise kernel that recovers the step's noise and consumes it at once --
no ``z`` buffer exists.  ``two_phase=True`` runs the reference wrapper's encode -> z -> decode instead (same result per sample
up to split-K summation order; tests/test_cycle_gpu.py compares the two).
"""
from dataclasses import dataclass

import torch

from .schedule import DDIMSchedule


@dataclass
class CycleDiffusionPipelineOutput:
    images: object
    nsfw_content_detected: object = None


class CycleDiffusionPipeline:
    def __init__(self, generator):
        """generator: wrappers._LatentGenerator (engine + U-Net + VAE + text encoder callable)."""
        self.g = generator
        self.engine = generator.engine

    @classmethod
    def from_wrapper(cls, wrapper):
        return cls(wrapper.generator)

    @torch.no_grad()
    def __call__(self, prompt, source_prompt, image=None, strength=0.8, num_inference_steps=50, guidance_scale=7.5,
                 source_guidance_scale=1, num_images_per_prompt=1, eta=0.1, generator=None, prompt_embeds=None, output_type='pt',
                 return_dict=True, callback=None, callback_steps=1, cross_attention_kwargs=None, clip_skip=None, two_phase=False):
        if strength < 0 or strength > 1:
            raise ValueError(f'The value of strength should in [0.0, 1.0] but is {strength}')
        if not isinstance(callback_steps, int) or callback_steps <= 0:
            raise ValueError('`callback_steps` has to be a positive integer')
        assert eta > 0, 'CycleDiffusion needs a stochastic sampler (eta > 0), ddim.py:268'
        g, e = self.g, self.engine
        prompts = [prompt] if isinstance(prompt, str) else list(prompt)
        sources = [source_prompt] if isinstance(source_prompt, str) else list(source_prompt)
        assert torch.is_tensor(image) and image.dim() == 4, 'image: float tensor [B,3,H,W] in [0,1] (PIL preprocessing is host glue)'
        B = image.shape[0] * num_images_per_prompt
        if num_images_per_prompt > 1:
            image = image.repeat_interleave(num_images_per_prompt, dim=0)
            prompts = [p for p in prompts for _ in range(num_images_per_prompt)]
            sources = [p for p in sources for _ in range(num_images_per_prompt)]
        if len(prompts) == 1 and B > 1:
            prompts, sources = prompts * B, sources * B
        rnd = lambda shape: torch.randn(shape, generator=generator)
        c_tgt = prompt_embeds if prompt_embeds is not None else g.get_learned_conditioning(prompts)
        c_src = g.get_learned_conditioning(sources)
        uc = g.get_learned_conditioning(B * [''])
        S = num_inference_steps
        skip = S - min(int(S * strength), S)
        sched = DDIMSchedule(S, eta, skip, g.alphas_cumprod)
        x = e.shift_scale(image, -0.5, 2.0)
        moments = g.encode_first_stage(x)
        lat_shape = (B, moments.shape[1] // 2, moments.shape[2], moments.shape[3])
        x0 = e.vae_posterior(moments, rnd(lat_shape) if g.sample_posterior else None, g.scale_factor)
        n_rec = sched.refine_steps
        noise = torch.zeros((n_rec + 1,) + lat_shape)
        noise[0] = rnd(lat_shape)
        for i in range(n_rec):
            if sched.refine_steps - 1 - i != 0:
                noise[1 + i] = rnd(lat_shape)
        if two_phase:
            z = g.unet.latent_encode(x0, c_src, uc, source_guidance_scale, sched, n_rec, noise)
            latents = g.unet.latent_decode(z, c_tgt, uc, guidance_scale, sched)
        else:
            latents = g.unet.cycle_lockstep(x0, c_src, c_tgt, uc, source_guidance_scale, guidance_scale, sched, noise)
        if callback is not None:
            callback(n_rec - 1, sched.t_loop[-1], latents)
        img = e.shift_scale(g.decode_first_stage(latents), 1.0, 0.5).clamp(0, 1)
        if output_type == 'np':
            img = img.permute(0, 2, 3, 1).float().cpu().numpy()
        elif output_type == 'pil':
            from PIL import Image
            arr = (img.permute(0, 2, 3, 1).float().cpu().numpy() * 255).round().astype('uint8')
            img = [Image.fromarray(a) for a in arr]
        if not return_dict:
            return (img, None)
        return CycleDiffusionPipelineOutput(images=img)
