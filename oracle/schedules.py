"""Oracle: diffusion schedules, restated from the reference (test infrastructure only).

Latent models (SD / LDM):
  * betas / alphas_cumprod     ref ldm/modules/diffusionmodules/util.py:21-43 (make_beta_schedule, "linear"),
                               ref ldm/models/diffusion/ddpm.py:117-138 (register_schedule: fp64 numpy cumprod -> fp32 buffers)
  * ddim timesteps             ref ldm/modules/diffusionmodules/util.py:46-61 (make_ddim_timesteps, "uniform", +1 offset)
  * ddim alphas/sigmas         ref ldm/modules/diffusionmodules/util.py:64-75 (make_ddim_sampling_parameters, fp32 torch arithmetic)
                               ref ldm/models/diffusion/ddim.py:25-55 (make_schedule)
Pixel models (DDPM / i-DDPM):
  * betas                      ref model/lib/ddpm_ddim/utils/diffusion_utils.py:5-9 (fp64 linspace), model/gan_wrapper/ddpm_ddim_wrapper.py:345-352 (-> .float())
  * alpha-bar                  ref ddpm_ddim_wrapper.py:194-199 ((1-b).cumprod(0) in fp32, re-done every step)
  * logvar                     ref ddpm_ddim_wrapper.py:354-373 (fp64 numpy posterior variance, log(max(.,1e-20)))
"""
import numpy as np
import torch


def ldm_alphas_cumprod(n_timestep=1000, linear_start=0.00085, linear_end=0.012):
    """fp32 alphas_cumprod buffer exactly as LatentDiffusion.register_schedule builds it."""
    betas = (torch.linspace(linear_start ** 0.5, linear_end ** 0.5, n_timestep, dtype=torch.float64) ** 2).numpy()
    alphas = 1.0 - betas
    alphas_cumprod = np.cumprod(alphas, axis=0)
    return torch.tensor(alphas_cumprod, dtype=torch.float32)


def ddim_timesteps(num_ddim_timesteps, num_ddpm_timesteps=1000):
    c = num_ddpm_timesteps // num_ddim_timesteps
    ts = np.asarray(list(range(0, num_ddpm_timesteps, c)))[:num_ddim_timesteps]
    return ts + 1


class DDIMTables:
    """Per-step fp32 coefficient tables of DDIMSampler.make_schedule (ddim.py:25-55)."""

    def __init__(self, S, eta, alphas_cumprod=None, num_ddpm_timesteps=1000):
        if alphas_cumprod is None:
            alphas_cumprod = ldm_alphas_cumprod(num_ddpm_timesteps)
        self.timesteps = ddim_timesteps(S, num_ddpm_timesteps)
        ac = alphas_cumprod.to(torch.float32)
        alphas = ac[self.timesteps]
        alphas_prev = np.asarray([ac[0]] + ac[self.timesteps[:-1]].tolist())
        # dtype flow of the reference, reproduced object-for-object: ``alphas`` is an fp32 torch tensor,
        # ``alphas_prev`` a float64 *numpy* array (np.asarray of python floats), and the mixed
        # numpy/torch expression below yields a float64 torch tensor.  torch.full(..., table[i]) later
        # rounds each entry to fp32 -- that fp32 value is what the step arithmetic consumes.
        sigmas = eta * np.sqrt((1 - alphas_prev) / (1 - alphas) * (1 - alphas / alphas_prev))
        self.alphas = alphas                                   # fp32
        self.alphas_prev = alphas_prev                         # fp64 numpy array of fp32-representable values
        self.sigmas = sigmas                                   # fp64
        self.sqrt_one_minus_alphas = np.sqrt(1.0 - alphas)     # fp32


def pixel_betas(beta_start=1e-4, beta_end=2e-2, T=1000):
    return np.linspace(beta_start, beta_end, T, dtype=np.float64)


def pixel_logvar(betas64):
    alphas = 1.0 - betas64
    alphas_cumprod = np.cumprod(alphas, axis=0)
    alphas_cumprod_prev = np.append(1.0, alphas_cumprod[:-1])
    posterior_variance = betas64 * (1.0 - alphas_cumprod_prev) / (1.0 - alphas_cumprod)
    return np.log(np.maximum(posterior_variance, 1e-20))


def pixel_seq(custom_steps, es_steps, t_0=999):
    """seq_inv / seq_inv_next of DDPMDDIMWrapper.encode/generate (ddpm_ddim_wrapper.py:393-400)."""
    if (t_0 + 1) % custom_steps == 0:
        seq_inv = range(0, t_0 + 1, (t_0 + 1) // custom_steps)
        assert len(seq_inv) == custom_steps
    else:
        seq_inv = np.linspace(0, 1, custom_steps) * t_0
    seq_inv = [int(s) for s in list(seq_inv)][:es_steps]
    seq_inv_next = ([-1] + list(seq_inv[:-1]))[:es_steps]
    return seq_inv, seq_inv_next
