"""CPU oracle for the CycleDiffusion hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``cycle_diffusion_b200/`` may import,
call or link this package.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs use it, and there
only as the checker / the CPU baseline, never as the thing shipped.

The oracle is a plain PyTorch fp32 *functional* restatement (``torch.nn.functional``
calls on a flat ``state_dict`` keyed with the reference's own parameter names) of

* the SD / LDM U-Net            -> ``oracle.unet_openai``   (ref: ldm/modules/diffusionmodules/openaimodel.py, ldm/modules/attention.py)
* the i-DDPM pixel U-Net        -> ``oracle.unet_iddpm``    (ref: model/lib/ddpm_ddim/models/improved_ddpm/unet.py)
* the KL-f8 VAE encoder/decoder -> ``oracle.vae_kl``        (ref: ldm/modules/diffusionmodules/model.py, ldm/models/autoencoder.py)
* DDIM / DDPM schedules         -> ``oracle.schedules``     (ref: ldm/modules/diffusionmodules/util.py, ldm/models/diffusion/ddim.py, ddpm_ddim/utils/diffusion_utils.py)
* DPM-Encoder + decode loops    -> ``oracle.dpm_encoder``   (ref: ldm/models/diffusion/ddim.py, model/gan_wrapper/*.py)

Pinning: ``tests/golden/make_golden.py`` imports the real reference modules from
``/root/reference`` (build container only), loads the *same* synthetic state_dict
into them, and commits their outputs under ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` checks this restatement against those fixtures.
The reference ships no tests / golden vectors of its own.
"""
