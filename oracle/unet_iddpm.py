"""Oracle: improved-DDPM (AFHQ/FFHQ) pixel U-Net forward, functional fp32 restatement (test infrastructure only).

Follows:
  create_model / AFHQ_DICT     ref model/lib/ddpm_ddim/models/improved_ddpm/script_util.py:5-99
  UNetModel.forward            ref .../improved_ddpm/unet.py:639-668 (topology :476-626)
  ResBlock._forward            ref unet.py:241-261 (scale-shift norm, res-block up/down)
  AttentionBlock._forward      ref unet.py:304-310
  QKVAttentionLegacy.forward   ref unet.py:342-363 (q*d^-1/4, k*d^-1/4, fp32 softmax)
  timestep_embedding           ref .../improved_ddpm/nn.py:103-121 ([cos | sin], float timesteps)
"""
import math
import torch
import torch.nn.functional as F

from .unet_openai import timestep_embedding


def afhq_config(image_size=256):
    if image_size == 256:
        mult = (1, 1, 2, 2, 4, 4)
    elif image_size == 128:
        mult = (1, 1, 2, 3, 4)
    elif image_size == 64:
        mult = (1, 2, 3, 4)
    else:
        raise ValueError(image_size)
    return dict(image_size=image_size, in_channels=3, out_channels=6, model_channels=128, num_res_blocks=1,
                channel_mult=mult, attention_resolutions=(image_size // 16,), num_head_channels=64)


def plan(cfg):
    """('conv'|'res'|'attn'|'resdown'|'resup') lists mirroring unet.py:476-626, with channel bookkeeping."""
    mc, mult, nrb, ar = cfg['model_channels'], cfg['channel_mult'], cfg['num_res_blocks'], cfg['attention_resolutions']
    inp = [['conv']]
    ds = 1
    for level in range(len(mult)):
        for _ in range(nrb):
            inp.append(['res', 'attn'] if ds in ar else ['res'])
        if level != len(mult) - 1:
            inp.append(['resdown'])
            ds *= 2
    mid = ['res', 'attn', 'res']
    out = []
    for level in reversed(range(len(mult))):
        for i in range(nrb + 1):
            layers = ['res'] + (['attn'] if ds in ar else [])
            if level and i == nrb:
                layers.append('resup')
                ds //= 2
            out.append(layers)
    return inp, mid, out


def _gn(sd, p, x):
    return F.group_norm(x.float(), 32, sd[p + '.weight'], sd[p + '.bias'], 1e-5)


def _resblock(sd, p, x, emb, updown=None):
    h = F.silu(_gn(sd, p + '.in_layers.0', x))
    if updown == 'down':
        h = F.avg_pool2d(h, 2, 2)
        x = F.avg_pool2d(x, 2, 2)
    elif updown == 'up':
        h = F.interpolate(h, scale_factor=2, mode='nearest')
        x = F.interpolate(x, scale_factor=2, mode='nearest')
    h = F.conv2d(h, sd[p + '.in_layers.2.weight'], sd[p + '.in_layers.2.bias'], padding=1)
    emb_out = F.linear(F.silu(emb), sd[p + '.emb_layers.1.weight'], sd[p + '.emb_layers.1.bias'])[..., None, None]
    scale, shift = torch.chunk(emb_out, 2, dim=1)
    h = _gn(sd, p + '.out_layers.0', h) * (1 + scale) + shift
    h = F.conv2d(F.silu(h), sd[p + '.out_layers.3.weight'], sd[p + '.out_layers.3.bias'], padding=1)
    if (p + '.skip_connection.weight') in sd:
        x = F.conv2d(x, sd[p + '.skip_connection.weight'], sd[p + '.skip_connection.bias'])
    return x + h


def _attnblock(sd, p, x, head_ch):
    b, c, *spatial = x.shape
    x = x.reshape(b, c, -1)
    qkv = F.conv1d(_gn(sd, p + '.norm', x), sd[p + '.qkv.weight'], sd[p + '.qkv.bias'])
    n_heads = c // head_ch
    bs, width, length = qkv.shape
    ch = width // (3 * n_heads)
    q, k, v = qkv.reshape(bs * n_heads, ch * 3, length).split(ch, dim=1)
    scale = 1 / math.sqrt(math.sqrt(ch))
    weight = torch.einsum('bct,bcs->bts', q * scale, k * scale)
    weight = torch.softmax(weight.float(), dim=-1)
    a = torch.einsum('bts,bcs->bct', weight, v).reshape(bs, -1, length)
    h = F.conv1d(a, sd[p + '.proj_out.weight'], sd[p + '.proj_out.bias'])
    return (x + h).reshape(b, c, *spatial)


def unet_forward(sd, cfg, x, timesteps, prefix=''):
    """x [B,3,R,R], timesteps [B] float -> [B,6,R,R] (learn_sigma; callers keep the first 3 channels)."""
    P = prefix
    inp, mid, outb = plan(cfg)
    hc = cfg['num_head_channels']
    emb = timestep_embedding(timesteps, cfg['model_channels'])
    emb = F.linear(F.silu(F.linear(emb, sd[P + 'time_embed.0.weight'], sd[P + 'time_embed.0.bias'])),
                   sd[P + 'time_embed.2.weight'], sd[P + 'time_embed.2.bias'])

    def run(block, bp, h):
        for li, kind in enumerate(block):
            p = f'{bp}.{li}'
            if kind == 'conv':
                h = F.conv2d(h, sd[p + '.weight'], sd[p + '.bias'], padding=1)
            elif kind == 'res':
                h = _resblock(sd, p, h, emb)
            elif kind == 'resdown':
                h = _resblock(sd, p, h, emb, 'down')
            elif kind == 'resup':
                h = _resblock(sd, p, h, emb, 'up')
            elif kind == 'attn':
                h = _attnblock(sd, p, h, hc)
        return h

    hs = []
    h = x
    for i, block in enumerate(inp):
        h = run(block, f'{P}input_blocks.{i}', h)
        hs.append(h)
    h = run(mid, f'{P}middle_block', h)
    for i, block in enumerate(outb):
        h = torch.cat([h, hs.pop()], dim=1)
        h = run(block, f'{P}output_blocks.{i}', h)
    h = F.silu(_gn(sd, P + 'out.0', h))
    return F.conv2d(h, sd[P + 'out.2.weight'], sd[P + 'out.2.bias'], padding=1)
