"""Oracle: DPM-Encoder inversion + decode-with-recovered-noise loops (test infrastructure only).

Restates, in reference evaluation order (every tensor op is a separate fp32 torch op, no fusing):

Latent models
  * _ddpm_ddim_encoding            ref ldm/models/diffusion/ddim.py:450-501
  * sample_xt_next                 ref ldm/models/diffusion/ddim.py:582-601
  * compute_eps                    ref ldm/models/diffusion/ddim.py:545-580
  * ddim_sampling_with_eps         ref ldm/models/diffusion/ddim.py:395-448
  * p_sample_ddim_with_eps         ref ldm/models/diffusion/ddim.py:603-646
  * wrapper glue (ensemble loops)  ref model/gan_wrapper/stable_diffusion_stochastic_text_wrapper.py:142-206
Pixel models
  * encode / generate              ref model/gan_wrapper/ddpm_ddim_wrapper.py:392-523
  * sample_xt / sample_xt_next / compute_eps / denoising_step_with_eps   ref ddpm_ddim_wrapper.py:114-314
  * denoising_step / extract       ref model/lib/ddpm_ddim/utils/diffusion_utils.py:12-136

RNG contract: every draw is a ``torch.randn(shape)`` from the global CPU generator, in the
reference's order, so ``torch.manual_seed(s)`` before a call reproduces the reference CPU run.
"""
import numpy as np
import torch

from .schedules import DDIMTables, pixel_betas, pixel_logvar, pixel_seq


# --------------------------------------------------------------------------------------
# latent models (DDIMSampler)
# --------------------------------------------------------------------------------------

def _guided_eps(unet_fn, x, t, c, uc, scale):
    """CFG batching of ddim.py:550-559 / 608-617 (uncond first)."""
    if uc is None or scale == 1.0:
        return unet_fn(x, t, c)
    if scale == 0:
        return unet_fn(x, t, uc)
    x_in = torch.cat([x] * 2)
    t_in = torch.cat([t] * 2)
    c_in = torch.cat([uc, c])
    e_uc, e_c = unet_fn(x_in, t_in, c_in).chunk(2)
    return e_uc + scale * (e_c - e_uc)


def _coeffs(tab, index, b):
    a_t = torch.full((b, 1, 1, 1), tab.alphas[index])
    a_prev = torch.full((b, 1, 1, 1), tab.alphas_prev[index])
    sigma_t = torch.full((b, 1, 1, 1), tab.sigmas[index])
    sqrt_1m_at = torch.full((b, 1, 1, 1), tab.sqrt_one_minus_alphas[index])
    return a_t, a_prev, sigma_t, sqrt_1m_at


def latent_sample_xt_next(tab, x0, xt, index):
    if index == 0:
        return x0
    b = x0.shape[0]
    a_t, a_prev, sigma_t, _ = _coeffs(tab, index, b)
    e_t = (xt - a_t.sqrt() * x0) / (1 - a_t).sqrt()
    dir_xt = (1. - a_prev - sigma_t ** 2).sqrt() * e_t
    noise = sigma_t * torch.randn(x0.shape)
    return a_prev.sqrt() * x0 + dir_xt + noise


def latent_compute_eps(tab, unet_fn, xt, xt_next, c, uc, t, index, scale, temperature=1.):
    b = xt.shape[0]
    e_t = _guided_eps(unet_fn, xt, t, c, uc, scale)
    a_t, a_prev, sigma_t, sqrt_1m_at = _coeffs(tab, index, b)
    pred_x0 = (xt - sqrt_1m_at * e_t) / a_t.sqrt()
    dir_xt = (1. - a_prev - sigma_t ** 2).sqrt() * e_t
    return (xt_next - a_prev.sqrt() * pred_x0 - dir_xt) / sigma_t / temperature


def latent_encode(unet_fn, x0, c, uc, S, eta, skip_steps, white_box_steps, scale, alphas_cumprod=None):
    """-> z_list = [x_T, eps_first, ..., eps_last]; mirrors DDIMSampler.ddpm_ddim_encoding."""
    assert eta > 0
    tab = DDIMTables(S, eta, alphas_cumprod)
    b = x0.shape[0]
    timesteps = tab.timesteps
    time_range = np.flip(timesteps)
    refine_steps = timesteps.shape[0] - skip_steps
    refine_time_range = time_range[-refine_steps:]
    at = tab.alphas[refine_steps - 1]
    xt = at.sqrt() * x0 + (1 - at).sqrt() * torch.randn(x0.shape)
    z_list = [xt]
    for i, step in enumerate(refine_time_range):
        index = refine_steps - i - 1
        ts = torch.full((b,), int(step), dtype=torch.long)
        if i < white_box_steps - skip_steps - 1:
            xt_next = latent_sample_xt_next(tab, x0, xt, index)
            eps = latent_compute_eps(tab, unet_fn, xt, xt_next, c, uc, ts, index, scale)
            xt = xt_next
            z_list.append(eps)
        else:
            break
    return z_list


def latent_decode(unet_fn, x_T, eps_list, c, uc, S, eta, skip_steps, scale, alphas_cumprod=None, temperature=1.):
    """eps_list: [B, n, C, h, w]; mirrors DDIMSampler.sample_with_eps / ddim_sampling_with_eps."""
    tab = DDIMTables(S, eta, alphas_cumprod)
    b = x_T.shape[0]
    timesteps = tab.timesteps
    time_range = np.flip(timesteps)
    refine_steps = timesteps.shape[0] - skip_steps
    refine_time_range = time_range[-refine_steps:]
    img = x_T
    for i, step in enumerate(refine_time_range):
        index = refine_steps - i - 1
        ts = torch.full((b,), int(step), dtype=torch.long)
        e_t = _guided_eps(unet_fn, img, ts, c, uc, scale)
        a_t, a_prev, sigma_t, sqrt_1m_at = _coeffs(tab, index, b)
        pred_x0 = (img - sqrt_1m_at * e_t) / a_t.sqrt()
        dir_xt = (1. - a_prev - sigma_t ** 2).sqrt() * e_t
        if i < eps_list.shape[1]:
            noise = sigma_t * eps_list[:, i] * temperature
        else:
            noise = sigma_t * torch.randn(img.shape) * temperature
        img = a_prev.sqrt() * pred_x0 + dir_xt + noise
    return img


def latent_refine(unet_fn, x0, c, uc, S, refine_steps, scale=1.0, alphas_cumprod=None):
    """DDIMSampler.refine -> _refine (ddim.py:114-168, 339-393) as latentdiff_stochastic_wrapper.py:68-77 calls it (eta = 1): x_t at
    ddim_alphas[refine_steps - 1] from a fresh draw, then p_sample_ddim over the last `refine_steps` timesteps, fresh noise per step."""
    tab = DDIMTables(S, 1.0, alphas_cumprod)
    assert refine_steps < tab.timesteps.shape[0]
    at = tab.alphas[refine_steps - 1]
    img = at.sqrt() * x0 + (1 - at).sqrt() * torch.randn(x0.shape)
    time_range = np.flip(tab.timesteps)[-refine_steps:]
    b = x0.shape[0]
    for i, step in enumerate(time_range):
        index = refine_steps - i - 1
        ts = torch.full((b,), int(step), dtype=torch.long)
        e_t = _guided_eps(unet_fn, img, ts, c, uc, scale)
        a_t, a_prev, sigma_t, sqrt_1m_at = _coeffs(tab, index, b)
        pred_x0 = (img - sqrt_1m_at * e_t) / a_t.sqrt()
        dir_xt = (1. - a_prev - sigma_t ** 2).sqrt() * e_t
        noise = sigma_t * torch.randn(img.shape) * 1.0
        img = a_prev.sqrt() * pred_x0 + dir_xt + noise
    return img


class LatentCycle:
    """SDStochasticTextWrapper / LatentDiffStochasticTextWrapper restated (SDW:100-253, LDW:102-252).

    unet_fn(x, t, ctx), vae_moments_fn(img in [-1,1]) -> [B,8,h,w], vae_decode_fn(z) -> img,
    cond_fn(list[str]) -> [B,77,D].  ``sample_posterior`` True for SD (ddpm.py:536-543), False for
    the latentdiff copy (latentdiff/.../ddpm.py:537-538 uses the mean).
    """

    def __init__(self, unet_fn, vae_moments_fn, vae_decode_fn, cond_fn, *, custom_steps, eta, white_box_steps,
                 skip_steps, encoder_unconditional_guidance_scales, decoder_unconditional_guidance_scales, n_trials,
                 channels=4, latent_size=64, resolution=512, scale_factor=0.18215, sample_posterior=True,
                 alphas_cumprod=None):
        self.unet_fn, self.vae_moments_fn, self.vae_decode_fn, self.cond_fn = unet_fn, vae_moments_fn, vae_decode_fn, cond_fn
        self.custom_steps, self.eta, self.white_box_steps, self.skip_steps = custom_steps, eta, white_box_steps, skip_steps
        self.enc_scales, self.dec_scales, self.n_trials = encoder_unconditional_guidance_scales, decoder_unconditional_guidance_scales, n_trials
        self.channels, self.latent_size, self.resolution = channels, latent_size, resolution
        self.scale_factor, self.sample_posterior, self.alphas_cumprod = scale_factor, sample_posterior, alphas_cumprod

    def first_stage_encode(self, image01):
        image = (image01 - 0.5) * 2.0
        assert image.shape[2] == image.shape[3] == self.resolution
        moments = self.vae_moments_fn(image)
        mean, logvar = torch.chunk(moments, 2, dim=1)
        logvar = torch.clamp(logvar, -30.0, 20.0)
        std = torch.exp(0.5 * logvar)
        if self.sample_posterior:
            z = mean + std * torch.randn(mean.shape)
        else:
            z = mean
        return self.scale_factor * z

    def encode(self, image01, encode_text):
        x0 = self.first_stage_encode(image01)
        bsz = image01.shape[0]
        z_ensemble = []
        for _ in range(self.n_trials):
            for enc_scale in self.enc_scales:
                for skip in self.skip_steps:
                    uc = self.cond_fn(bsz * [""])
                    c = self.cond_fn(encode_text)
                    z_list = latent_encode(self.unet_fn, x0, c, uc, self.custom_steps, self.eta, skip,
                                           self.white_box_steps, enc_scale, self.alphas_cumprod)
                    z_ensemble.append(torch.stack(z_list, dim=1).view(bsz, -1))
        return z_ensemble

    def generate(self, z_ensemble, decode_text):
        imgs = []
        for i, z in enumerate(z_ensemble):
            skip = self.skip_steps[i % len(self.skip_steps)]
            bsz = z.shape[0]
            eps_list = z.view(bsz, self.white_box_steps - skip, self.channels, self.latent_size, self.latent_size)
            x_T, eps_list = eps_list[:, 0], eps_list[:, 1:]
            for dec_scale in self.dec_scales:
                uc = self.cond_fn(bsz * [""])
                c = self.cond_fn(decode_text)
                sample = latent_decode(self.unet_fn, x_T, eps_list, c, uc, self.custom_steps, self.eta, skip,
                                       dec_scale, self.alphas_cumprod)
                imgs.append(self.vae_decode_fn(1. / self.scale_factor * sample))
        return imgs

    def forward_all(self, z_ensemble, decode_text):
        """post-processed ensemble, (x+1)/2, no ranking (D-CLIP ranking is out of scope)."""
        return [(im + 1.0) / 2.0 for im in self.generate(z_ensemble, decode_text)]


# --------------------------------------------------------------------------------------
# pixel models (DDPMDDIMWrapper)
# --------------------------------------------------------------------------------------

def _extract(a, t, x_shape):
    bs, = t.shape
    out = torch.gather(torch.as_tensor(a, dtype=torch.float), 0, t.long())
    return out.reshape((bs,) + (1,) * (len(x_shape) - 1))


def _eps_model(model_fn, xt, t):
    et = model_fn(xt, t)
    if et.shape != xt.shape:
        et, _ = torch.split(et, et.shape[1] // 2, dim=1)
    return et


def pixel_sample_xt(x0, t, b):
    at = _extract((1.0 - b).cumprod(dim=0), t, x0.shape)
    return at.sqrt() * x0 + (1 - at).sqrt() * torch.randn(x0.shape)


def pixel_sample_xt_next(x0, xt, t, t_next, sampling_type, b, eta):
    bt = _extract(b, t, xt.shape)
    at = _extract((1.0 - b).cumprod(dim=0), t, xt.shape)
    at_next = _extract((1.0 - b).cumprod(dim=0), t_next, xt.shape)
    if sampling_type == 'ddpm':
        w0 = at_next.sqrt() * bt / (1 - at)
        wt = (1 - bt).sqrt() * (1 - at_next) / (1 - at)
        mean = w0 * x0 + wt * xt
        var = bt * (1 - at_next) / (1 - at)
        return mean + var.sqrt() * torch.randn(x0.shape)
    et = (xt - at.sqrt() * x0) / (1 - at).sqrt()
    c1 = eta * ((1 - at / at_next) * (1 - at_next) / (1 - at)).sqrt()
    c2 = ((1 - at_next) - c1 ** 2).sqrt()
    return at_next.sqrt() * x0 + c2 * et + c1 * torch.randn(x0.shape)


def pixel_compute_eps(xt, xt_next, t, t_next, model_fn, sampling_type, b, logvars, eta):
    et = _eps_model(model_fn, xt, t)
    logvar = _extract(logvars, t, xt.shape)
    bt = _extract(b, t, xt.shape)
    at = _extract((1.0 - b).cumprod(dim=0), t, xt.shape)
    at_next = _extract((1.0 - b).cumprod(dim=0), t_next, xt.shape)
    if sampling_type == 'ddpm':
        weight = bt / torch.sqrt(1 - at)
        mean = 1 / torch.sqrt(1.0 - bt) * (xt - weight * et)
        return (xt_next - mean) / torch.exp(0.5 * logvar)
    x0_t = (xt - et * (1 - at).sqrt()) / at.sqrt()
    c1 = eta * ((1 - at / at_next) * (1 - at_next) / (1 - at)).sqrt()
    c2 = ((1 - at_next) - c1 ** 2).sqrt()
    return (xt_next - at_next.sqrt() * x0_t - c2 * et) / c1


def pixel_denoise(xt, eps, t, t_next, model_fn, sampling_type, b, logvars, eta):
    """denoising_step_with_eps (eps given) / denoising_step (eps None -> fresh randn)."""
    et = _eps_model(model_fn, xt, t)
    logvar = _extract(logvars, t, xt.shape)
    bt = _extract(b, t, xt.shape)
    at = _extract((1.0 - b).cumprod(dim=0), t, xt.shape)
    if t_next.sum() == -t_next.shape[0]:
        at_next = torch.ones_like(at)
    else:
        at_next = _extract((1.0 - b).cumprod(dim=0), t_next, xt.shape)
    if sampling_type == 'ddpm':
        weight = bt / torch.sqrt(1 - at)
        mean = 1 / torch.sqrt(1.0 - bt) * (xt - weight * et)
        noise = eps if eps is not None else torch.randn(xt.shape)
        mask = 1 - (t == 0).float()
        mask = mask.reshape((xt.shape[0],) + (1,) * (len(xt.shape) - 1))
        return (mean + mask * torch.exp(0.5 * logvar) * noise).float()
    x0_t = (xt - et * (1 - at).sqrt()) / at.sqrt()
    if eta == 0:
        return at_next.sqrt() * x0_t + (1 - at_next).sqrt() * et
    c1 = eta * ((1 - at / at_next) * (1 - at_next) / (1 - at)).sqrt()
    c2 = ((1 - at_next) - c1 ** 2).sqrt()
    noise = eps if eps is not None else torch.randn(xt.shape)
    return at_next.sqrt() * x0_t + c2 * et + c1 * noise


class PixelCycle:
    """DDPMDDIMWrapper restated (ddpm_ddim_wrapper.py:317-538), refine_steps == 0 path plus refine.

    Works per sample for B>1 exactly like the reference would if its ``at > at_next`` tensor-bool
    check (ddpm_ddim_wrapper.py:216) did not raise for B>1: all samples share t, so it is elementwise.
    """

    def __init__(self, model_fn, *, sample_type, custom_steps, es_steps, eta=None, t_0=None, refine_steps=0,
                 refine_iterations=1, resolution=256, channels=3, beta_start=1e-4, beta_end=2e-2, T=1000):
        self.model_fn = model_fn
        self.sample_type, self.custom_steps, self.es_steps = sample_type, custom_steps, es_steps
        self.eta, self.t_0 = eta, (t_0 if t_0 is not None else 999)
        self.refine_steps, self.refine_iterations = refine_steps, refine_iterations
        if sample_type == 'ddim':
            assert eta > 0
        else:
            assert eta is None
        betas64 = pixel_betas(beta_start, beta_end, T)
        self.betas = torch.from_numpy(betas64).float()
        self.logvar = pixel_logvar(betas64)
        self.resolution, self.channels = resolution, channels
        self.latent_dim = resolution ** 2 * channels * es_steps

    def encode(self, image01):
        seq_inv, seq_inv_next = pixel_seq(self.custom_steps, self.es_steps, self.t_0)
        x0 = (image01 - 0.5) * 2.0
        assert x0.shape[2] == x0.shape[3] == self.resolution
        bsz = x0.shape[0]
        T = torch.ones(bsz) * (self.es_steps - 1)
        xT = pixel_sample_xt(x0, T, self.betas)
        z_list = [xT]
        xt = xT
        for it, (i, j) in enumerate(zip(reversed(seq_inv), reversed(seq_inv_next))):
            t = torch.ones(bsz) * i
            t_next = torch.ones(bsz) * j
            if it < self.es_steps - 1:
                xt_next = pixel_sample_xt_next(x0, xt, t, t_next, self.sample_type, self.betas, self.eta)
                eps = pixel_compute_eps(xt, xt_next, t, t_next, self.model_fn, self.sample_type, self.betas,
                                        self.logvar, self.eta)
                xt = xt_next
                z_list.append(eps)
            else:
                break
        z = torch.stack(z_list, dim=1).view(bsz, -1)
        assert z.shape[1] == self.latent_dim
        return z

    def generate(self, z):
        seq_inv, seq_inv_next = pixel_seq(self.custom_steps, self.es_steps, self.t_0)
        bsz = z.shape[0]
        eps_list = z.view(bsz, self.es_steps, self.channels, self.resolution, self.resolution)
        x, eps_list = eps_list[:, 0], eps_list[:, 1:]
        for it, (i, j) in enumerate(zip(reversed(seq_inv), reversed(seq_inv_next))):
            t = torch.ones(bsz) * i
            t_next = torch.ones(bsz) * j
            eps = eps_list[:, it] if it < self.es_steps - 1 else None
            x = pixel_denoise(x, eps, t, t_next, self.model_fn, self.sample_type, self.betas, self.logvar, self.eta)
        if self.refine_steps != 0:
            for _ in range(self.refine_iterations):
                t = torch.ones(bsz) * self.refine_steps - 1
                x = pixel_sample_xt(x, t, self.betas)
                assert self.refine_steps < self.custom_steps
                for i, j in zip(reversed(seq_inv[:self.refine_steps]), reversed(seq_inv_next[:self.refine_steps])):
                    t = torch.ones(bsz) * i
                    t_next = torch.ones(bsz) * j
                    x = pixel_denoise(x, None, t, t_next, self.model_fn, self.sample_type, self.betas, self.logvar, 1)
        return x

    def forward(self, z):
        return (self.generate(z) + 1.0) / 2.0
