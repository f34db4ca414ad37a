#!/usr/bin/env python
"""bench.py -- images/sec of the CycleDiffusion hot path (BASELINE.json config 2) on N B200s of one node.

One "step" = one full cycle over one batch of synthetic (image, source-text, target-text) triplets per GPU:
    VAE encode (+posterior sample) -> 50-step DPM-Encoder under the source condition (scale 1)
    -> 50-step decode under the target condition with classifier-free guidance 7.5 -> VAE decode -> (x+1)/2
on the Stable Diffusion v1-4 topology (859.5 M-param U-Net, KL-f8 VAE, random-init weights -- there are no checkpoints
offline), 512x512, batch 4 per GPU (README.md:153), fp32 end to end.

    python bench.py --gpus 1 --steps K --warmup W                  # our engine
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...                           # the reference's CPU path (oracle port), rank 0 only

Prints ONE JSON line (see the keys at the bottom).  `value` is timed with inputs resident in HBM; `e2e` goes through
the drop-in wrapper API with HOST buffers (H2D of image / conditioning / noise and D2H of the result inside the timed
region).  `roofline` comes from a separate untimed profiling pass (CUDA events around every launch of each kernel
family, inside libcdx); `cpu_baseline` times the CPU oracle on a bounded sample on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

if os.environ.get('NCCL_DEBUG', '').upper() == 'VERSION':     # NCCL prints its banner to stdout: keep stdout = the one JSON line
    os.environ['NCCL_DEBUG'] = 'WARN'
import torch  # noqa: E402

METRIC = 'images/sec (512x512, 50-step encode+decode)'
UNIT = 'images/s'
TFLOP_PER_IMAGE = 124.1            # 1.117 + 150 * 0.80327 + 2.515
UNET_GFLOP = 803.27                # per sample-forward
S_STEPS, ETA, DEC_SCALE, ENC_SCALE = 50, 0.1, 7.5, 1.0
B_PER_GPU, RES, LAT = 4, 512, 64


def peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm_gbs=p['hbm_gbs'], tflops=p['bf16_tflops'], tflops_sustained=p.get('bf16_tflops_sustained', p['bf16_tflops']),
                    source='measured (MEASURED_PEAKS.json: copy GB/s, cuBLAS bf16 TF/s)')
    return dict(hbm_gbs=6650.0, tflops=1590.0, tflops_sustained=1400.0, source='fallback (B200_PROFILING.md)')


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = 'clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,' \
            'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap'
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--id={self.index}', f'--query-gpu={q}', '--format=csv,noheader,nounits', '-lms', '200'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(',')])

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace('.', '').isdigit())
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace('.', '').isdigit()]
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith('active')})
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': max(mx) if mx else None, 'reasons': reasons, 'samples': len(sm)}


def synthetic_inputs(B, seed=0):
    """image U[0,1] (seed 0), conditioning N(0,1) [B,77,768] (seed 1)."""
    g0, g1 = torch.Generator().manual_seed(seed), torch.Generator().manual_seed(1 + seed)
    image = torch.rand(B, 3, RES, RES, generator=g0)
    c_src = torch.randn(B, 77, 768, generator=g1)
    c_tgt = torch.randn(B, 77, 768, generator=g1)
    uc = torch.randn(1, 77, 768, generator=g1).expand(B, 77, 768).contiguous()
    return image, c_src, c_tgt, uc


def encode_noise(sched, n_rec, shape, gen):
    noise = torch.zeros((n_rec + 1,) + tuple(shape))
    noise[0] = torch.randn(shape, generator=gen)
    for i in range(n_rec):
        if sched.refine_steps - 1 - i != 0:
            noise[1 + i] = torch.randn(shape, generator=gen)
    return noise


# ================================================================================================ our arm
def run_ours(args):
    import torch.distributed as dist
    from cycle_diffusion_b200 import specs
    from cycle_diffusion_b200.engine import Engine, UNet, VAE
    from cycle_diffusion_b200.schedule import DDIMSchedule
    from cycle_diffusion_b200.wrappers import SDStochasticTextWrapper, _LatentGenerator

    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    assert world == args.gpus, f'--gpus {args.gpus} but WORLD_SIZE={world}'
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    eng = Engine(local)
    if args.mma is not None:
        eng.set_mma_mode(args.mma)
    ucfg, vcfg = specs.sd_unet_config(768), specs.kl_f8_config()
    unet, vae = UNet(eng, ucfg, 'openai'), VAE(eng, vcfg)
    # frozen weights: rank 0 builds them, one NCCL broadcast of the packed blobs over NVLink, no other collective
    t0 = time.time()
    if rank == 0:
        unet.load_state_dict(specs.synth_state_dict(specs.openai_unet_params(ucfg), 1234))
        vae.load_state_dict(specs.synth_state_dict(specs.kl_vae_params(vcfg), 1235))
    bcast_ms = None
    if world > 1:
        torch.cuda.synchronize()
        dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for net in (unet, vae):
            dist.broadcast(net.blob_tensor(), src=0)
        e1.record()
        torch.cuda.synchronize()
        bcast_ms = e0.elapsed_time(e1)
        if rank != 0:
            unet.adopt_blob()
            vae.adopt_blob()
    load_s = time.time() - t0

    B = B_PER_GPU
    image, c_src, c_tgt, uc = synthetic_inputs(B, seed=rank)
    sched = DDIMSchedule(S_STEPS, ETA, 0)
    n_rec = sched.refine_steps
    gen = torch.Generator().manual_seed(7 + rank)
    post_noise = torch.randn(B, 4, LAT, LAT, generator=gen)
    enc_noise = encode_noise(sched, n_rec, (B, 4, LAT, LAT), gen)
    d = eng.device
    dev = dict(image=image.to(d), c_src=c_src.to(d), c_tgt=c_tgt.to(d), uc=uc.to(d), post=post_noise.to(d), noise=enc_noise.to(d))

    def cycle_resident():
        x = eng.shift_scale(dev['image'], -0.5, 2.0)
        x0 = eng.vae_posterior(vae.encode_moments(x), dev['post'], 0.18215)
        z = unet.latent_encode(x0, dev['c_src'], dev['uc'], ENC_SCALE, sched, n_rec, dev['noise'])
        s = unet.latent_decode(z, dev['c_tgt'], dev['uc'], DEC_SCALE, sched)
        return eng.shift_scale(vae.decode(eng.affine(s, 1. / 0.18215, 0.0)), 1.0, 0.5)

    # the drop-in wrapper over the SAME engine objects, fed with host tensors
    class _Cond:
        def __call__(self, texts):
            return pinned['uc'] if texts[0] == '' else (pinned['c_src'] if texts[0] == 'src' else pinned['c_tgt'])
    genr = _LatentGenerator(eng, unet, vae, _Cond(), 4, LAT, 0.18215, True)
    wrap = SDStochasticTextWrapper('synthetic', custom_steps=S_STEPS, eta=ETA, white_box_steps=S_STEPS + 1, skip_steps=[0],
                                   encoder_unconditional_guidance_scales=[ENC_SCALE], decoder_unconditional_guidance_scales=[DEC_SCALE],
                                   n_trials=1, generator=genr)
    pinned = {k: v.pin_memory() for k, v in dict(image=image, c_src=c_src, c_tgt=c_tgt, uc=uc).items()}
    out_host = torch.empty(B, 3, RES, RES).pin_memory()
    h2d = [0]

    def cycle_e2e():
        torch.manual_seed(99)
        img_d = pinned['image'].to(d, non_blocking=True)
        z = wrap.encode(img_d, B * ['src'])
        img = wrap(z, img_d, B * ['src'], B * ['tgt'])
        out_host.copy_(img, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        # image + (uc, c) for encode + (uc, c) for decode + VAE posterior noise + DPM-Encoder noise
        h2d[0] = 4 * (image.numel() + 4 * c_src.numel() + post_noise.numel() + enc_noise.numel())
        return img

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        l0 = eng.launches
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=d)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), eng.launches - l0

    clocks = ClockSampler(local)
    clocks.start()
    ms_total, launches = timed(cycle_resident, args.steps, args.warmup)
    clk = clocks.stop()
    e2e_steps = max(1, min(args.steps, 2))
    ms_e2e, _ = timed(cycle_e2e, e2e_steps, 1)
    value = world * B * args.steps / (ms_total / 1e3)
    e2e_value = world * B * e2e_steps / (ms_e2e / 1e3)

    # ---- roofline pass (untimed): per-kernel-family CUDA-event timing of one CFG U-Net call (batch 2B) on rank 0
    roof, families, unet_ms, stage_ms = None, {}, None, None
    if rank == 0:
        pk = peaks()
        x = torch.randn(2 * B, 4, LAT, LAT, device=d)
        t = torch.full((2 * B,), 501., device=d)
        ctx = torch.cat([dev['uc'], dev['c_tgt']])
        for _ in range(2):
            unet(x, t, ctx)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            unet(x, t, ctx)
        e1.record()
        torch.cuda.synchronize()
        unet_ms = e0.elapsed_time(e1) / 3
        eng.profile(True)
        unet(x, t, ctx)
        families = eng.profile_read()
        eng.profile(False)
        tensor_fams = {k: v for k, v in families.items() if v['flops'] > 0}
        if tensor_fams:
            top = max(tensor_fams, key=lambda k: tensor_fams[k]['ms'])
            f = tensor_fams[top]
            ach = f['flops'] / (f['ms'] * 1e-3) / 1e12
            traffic = None      # DRAM bytes of one captured launch of this family (ncu --set full; profiles/ncu_traffic.json)
            try:
                with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'profiles', 'ncu_traffic.json')) as fh:
                    traffic = json.load(fh).get(top)
            except (OSError, ValueError):
                pass
            roof = {'kernel': top, 'bound': 'tensor', 'achieved': round(ach, 2), 'peak': pk['tflops_sustained'], 'unit': 'TFLOP/s',
                    'frac': round(ach / pk['tflops_sustained'], 4),
                    # the fp32-faithful path issues 3 TF32 MMAs per product at half the bf16 rate: its own ceiling is peak / 6
                    'frac_of_3xtf32_ceiling': round(ach / (pk['tflops_sustained'] / 6.0), 4), 'traffic': traffic, 'launches_per_unet_call': f['launches'],
                    'avg_launch_ms': round(f['ms'] / f['launches'], 4),
                    'peak_source': pk['source'] + ' -- sustained bf16 dense; this path is fp32-faithful (see DESIGN.md)',
                    'whole_job_tflops': round(value * TFLOP_PER_IMAGE, 2)}
        hb = {k: v for k, v in families.items() if v['flops'] == 0 and v['bytes'] > 0}
        for k, v in hb.items():
            v['gbs'] = round(v['bytes'] / (v['ms'] * 1e-3) / 1e9, 1)
            v['frac_hbm'] = round(v['gbs'] / pk['hbm_gbs'], 4)
        for v in families.values():
            v['ms'] = round(v['ms'], 3)
            if v['flops']:
                v['tflops'] = round(v['flops'] / (v['ms'] * 1e-3) / 1e12, 2)

        # stage breakdown of one cycle (untimed extra pass, CUDA events on the launching stream)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
        ev[0].record()
        x_ = eng.shift_scale(dev['image'], -0.5, 2.0)
        x0_ = eng.vae_posterior(vae.encode_moments(x_), dev['post'], 0.18215)
        ev[1].record()
        z_ = unet.latent_encode(x0_, dev['c_src'], dev['uc'], ENC_SCALE, sched, n_rec, dev['noise'])
        ev[2].record()
        s_ = unet.latent_decode(z_, dev['c_tgt'], dev['uc'], DEC_SCALE, sched)
        ev[3].record()
        eng.shift_scale(vae.decode(eng.affine(s_, 1. / 0.18215, 0.0)), 1.0, 0.5)
        ev[4].record()
        torch.cuda.synchronize()
        stage_ms = {k: round(ev[i].elapsed_time(ev[i + 1]), 1) for i, k in
                    enumerate(['vae_encode', f'dpm_encode_{S_STEPS}x_unet_b{B}', f'decode_{S_STEPS}x_unet_b{2 * B}', 'vae_decode'])}

    # the CPU leg is timed on rank 0 of the single-GPU run only (the other ranks would just wait at the closing barrier)
    cpu = cpu_baseline_sample(quick=True) if (rank == 0 and world == 1 and not args.no_cpu) else None

    if rank == 0:
        line = {
            'metric': METRIC, 'value': round(value, 4), 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': round(ms_total / args.steps, 2), 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'fp32', 'data': 'synthetic (images U[0,1], conditioning N(0,1), random-init SD v1-4-topology weights)',
            'config': {'workload': 'BASELINE configs[1]: Stable Diffusion v1-4 512x512, 50-step DPMEncoder (scale 1) + 50-step CFG decode (scale 7.5), '
                                   'batch 4 per GPU', 'global_batch': world * B, 'steps_encode': S_STEPS, 'steps_decode': S_STEPS, 'eta': ETA,
                       'parallelism': f'dp{world} (images sharded, one NCCL weight broadcast)', 'mma_mode': 'ffma-fp32' if args.mma == 0 else 'tcgen05-3xTF32 (fp32-faithful)',
                       'l2': 'no flush: 3.8 GB of weights + >1 GB activations per U-Net call are streamed every call (>> 126 MB L2)',
                       'unet_calls_per_step': 2 * S_STEPS, 'unet_ms_cfg_batch8': round(unet_ms, 2) if unet_ms else None, 'stage_ms': stage_ms},
            'e2e': {'value': round(e2e_value, 4), 'unit': UNIT, 'h2d_bytes_per_step': h2d[0], 'd2h_bytes_per_step': 4 * out_host.numel(),
                    'steps': e2e_steps, 'api': 'SDStochasticTextWrapper.encode + forward (host tensors in, pinned host tensor out)'},
            'gpu_launches': launches, 'clocks': clk, 'roofline': roof, 'kernel_families': families, 'cpu_baseline': cpu,
            'weights_broadcast_ms': bcast_ms, 'setup_s': round(load_s, 1), 'workspace_gb': round(eng.workspace_bytes / 2 ** 30, 2),
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# ================================================================================================ CPU arms
def _host_threads():
    """CPU threads this process may really use: affinity mask, capped by the cgroup CPU quota (a container that shows
    128 CPUs but is throttled to a few cores thrashes with 128 ATen threads)."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    try:
        with open('/sys/fs/cgroup/cpu.max') as fh:
            quota, period = fh.read().split()
        if quota != 'max':
            n = min(n, max(1, int(float(quota) / float(period) + 0.5)))
    except (OSError, ValueError):
        pass
    return max(1, min(n, 64))


_CPU_STATE = {}


def cpu_baseline_sample(quick=True):
    """The reference's CPU path (oracle port: same ATen CPU kernels, fp32) on a bounded sample of the same workload.

    Sample: one warm SD U-Net sample-forward at batch 1 and (once per process) one VAE encode + decode of one 512x512
    image; images/s is extrapolated as 1 / (150 sample-forwards * t_unet + t_vae).  Weights and
    the discarded cold oneDNN call are set up once per process."""
    from cycle_diffusion_b200 import specs
    from oracle import unet_openai, vae_kl
    st = _CPU_STATE
    if not st:
        st['threads'] = _host_threads()
        torch.set_num_threads(st['threads'])
        st['ucfg'], st['vcfg'] = specs.sd_unet_config(768), specs.kl_f8_config()
        st['usd'] = specs.synth_state_dict(specs.openai_unet_params(st['ucfg']), 1234)
        st['vsd'] = specs.synth_state_dict(specs.kl_vae_params(st['vcfg']), 1235)
        g = torch.Generator().manual_seed(0)
        st['x'] = torch.randn(1, 4, LAT, LAT, generator=g)
        st['ctx'] = torch.randn(1, 77, 768, generator=g)
        st['img'] = torch.rand(1, 3, RES, RES, generator=g) * 2 - 1
        with torch.no_grad():
            unet_openai.unet_forward(st['usd'], st['ucfg'], st['x'], torch.tensor([501]), st['ctx'])      # cold call, discarded
            t0 = time.time()
            m = vae_kl.encode_moments(st['vsd'], st['vcfg'], st['img'])
            vae_kl.decode(st['vsd'], st['vcfg'], m[:, :4])
            st['t_vae'] = time.time() - t0
    n_calls = 1 if quick else 3
    with torch.no_grad():
        t0 = time.time()
        for _ in range(n_calls):
            unet_openai.unet_forward(st['usd'], st['ucfg'], st['x'], torch.tensor([501]), st['ctx'])
        t_unet = (time.time() - t0) / n_calls
    t_vae = st['t_vae']
    value = 1.0 / (3 * S_STEPS * t_unet + t_vae)
    return {'value': round(value, 6), 'unit': UNIT, 'cores': st['threads'], 'kind': 'port',
            'sample': f'{n_calls} warm SD U-Net sample-forward(s) at batch 1 ({t_unet:.2f} s each) + VAE enc/dec of one 512x512 image ({t_vae:.2f} s); '
                      f'extrapolated: 1 / (150 * t_unet + t_vae)', 'unet_s_per_sample_forward': round(t_unet, 3), 'vae_s': round(t_vae, 3)}


def run_reference(args):
    rank = int(os.environ.get('RANK', 0))
    if rank != 0:
        return
    vals = []
    t_all = time.time()
    for i in range(args.warmup + args.steps):
        r = cpu_baseline_sample(quick=True)
        if i >= args.warmup:
            vals.append(r)
        if time.time() - t_all > 200 and vals:   # keep the whole run within a few minutes
            break
    if not vals:
        vals = [r]
    v = sum(x['value'] for x in vals) / len(vals)
    cpu = dict(vals[-1])
    cpu['value'] = round(v, 6)
    line = {'impl': 'reference', 'metric': METRIC, 'value': round(v, 6), 'unit': UNIT, 'n_gpus': args.gpus, 'steps': len(vals), 'warmup': args.warmup,
            'ms_per_step': round(1e3 * B_PER_GPU / v, 1), 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'fp32',
            'data': 'synthetic', 'config': {'workload': 'BASELINE configs[1] (SD v1-4 512x512, 50+50 steps, CFG 7.5), CPU path on a bounded sample',
                                            'note': 'the reference is Python and cannot travel to the GPU box; this is its CPU restatement (oracle/), same ATen kernels'},
            'cpu_baseline': cpu, 'e2e': {'value': round(v, 6), 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}, 'gpu_launches': 0}
    print(json.dumps(line))


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--mma', type=int, default=None, help='0 = FFMA fp32 tiles, 1 = tcgen05 3xTF32 (default: engine default)')
    ap.add_argument('--no-cpu', action='store_true', help='skip the cpu_baseline leg')
    a = ap.parse_args()
    if a.impl == 'reference':
        run_reference(a)
    else:
        run_ours(a)
