"""Oracle: SD v1 / LDM text2img U-Net forward, functional fp32 restatement (test infrastructure only).

Follows, op for op:
  UNetModel.forward            ref ldm/modules/diffusionmodules/openaimodel.py:710-742 (topology built at :506-686)
  ResBlock._forward            ref openaimodel.py:255-275
  Downsample / Upsample        ref openaimodel.py:91-160
  timestep_embedding           ref ldm/modules/diffusionmodules/util.py:152-172  ([cos | sin])
  GroupNorm32 (eps 1e-5)       ref util.py:215-217
  SpatialTransformer.forward   ref ldm/modules/attention.py:250-261 (Normalize eps 1e-6 at :76-77)
  BasicTransformerBlock        ref attention.py:211-215
  CrossAttention.forward       ref attention.py:170-193
  GEGLU / FeedForward          ref attention.py:37-64 (exact erf GELU)
The state_dict uses the reference's parameter names (time_embed.0.weight, input_blocks.1.0.in_layers.0.weight, ...).
"""
import math
import torch
import torch.nn.functional as F


def default_sd_config(context_dim=768):
    """v1-inference.yaml:29-44 (SD) / txt2img-1p4B-eval.yaml:20-42 (LDM: context_dim 1280)."""
    return dict(in_channels=4, out_channels=4, model_channels=320, attention_resolutions=(4, 2, 1),
                num_res_blocks=2, channel_mult=(1, 2, 4, 4), num_heads=8, context_dim=context_dim)


def timestep_embedding(timesteps, dim, max_period=10000):
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(start=0, end=half, dtype=torch.float32) / half)
    args = timesteps[:, None].float() * freqs[None]
    emb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
    if dim % 2:
        emb = torch.cat([emb, torch.zeros_like(emb[:, :1])], dim=-1)
    return emb


def _gn(sd, p, x, eps):
    return F.group_norm(x.float(), 32, sd[p + '.weight'], sd[p + '.bias'], eps)


def _conv(sd, p, x, stride=1, padding=1):
    return F.conv2d(x, sd[p + '.weight'], sd.get(p + '.bias'), stride=stride, padding=padding)


def _lin(sd, p, x):
    return F.linear(x, sd[p + '.weight'], sd.get(p + '.bias'))


def _resblock(sd, p, x, emb):
    h = _conv(sd, p + '.in_layers.2', F.silu(_gn(sd, p + '.in_layers.0', x, 1e-5)))
    emb_out = _lin(sd, p + '.emb_layers.1', F.silu(emb))[..., None, None]
    h = h + emb_out
    h = _conv(sd, p + '.out_layers.3', F.silu(_gn(sd, p + '.out_layers.0', h, 1e-5)))
    if (p + '.skip_connection.weight') in sd:
        x = _conv(sd, p + '.skip_connection', x, padding=0)
    return x + h


def _attention(sd, p, x, context, heads):
    q = _lin(sd, p + '.to_q', x)
    ctx = x if context is None else context
    k = _lin(sd, p + '.to_k', ctx)
    v = _lin(sd, p + '.to_v', ctx)
    b, n, inner = q.shape
    d = inner // heads
    scale = d ** -0.5

    def split(t):  # 'b n (h d) -> (b h) n d'
        return t.reshape(b, t.shape[1], heads, d).permute(0, 2, 1, 3).reshape(b * heads, t.shape[1], d)

    q, k, v = split(q), split(k), split(v)
    sim = torch.einsum('bid,bjd->bij', q, k) * scale
    attn = sim.softmax(dim=-1)
    out = torch.einsum('bij,bjd->bid', attn, v)
    out = out.reshape(b, heads, n, d).permute(0, 2, 1, 3).reshape(b, n, inner)
    return _lin(sd, p + '.to_out.0', out)


def _spatial_transformer(sd, p, x, context, heads):
    b, c, h, w = x.shape
    x_in = x
    x = _gn(sd, p + '.norm', x, 1e-6)
    x = _conv(sd, p + '.proj_in', x, padding=0)
    x = x.permute(0, 2, 3, 1).reshape(b, h * w, c)
    t = p + '.transformer_blocks.0'
    x = _attention(sd, t + '.attn1', F.layer_norm(x, (c,), sd[t + '.norm1.weight'], sd[t + '.norm1.bias']), None, heads) + x
    x = _attention(sd, t + '.attn2', F.layer_norm(x, (c,), sd[t + '.norm2.weight'], sd[t + '.norm2.bias']), context, heads) + x
    y = F.layer_norm(x, (c,), sd[t + '.norm3.weight'], sd[t + '.norm3.bias'])
    y, gate = _lin(sd, t + '.ff.net.0.proj', y).chunk(2, dim=-1)
    y = y * F.gelu(gate)
    x = _lin(sd, t + '.ff.net.2', y) + x
    x = x.reshape(b, h, w, c).permute(0, 3, 1, 2)
    x = _conv(sd, p + '.proj_out', x, padding=0)
    return x + x_in


def _attention_block(sd, p, x, cfg):
    """AttentionBlock._forward + QKVAttentionLegacy (openaimodel.py:304-315, 318-351): the use_spatial_transformer=False attention of
    the unconditional LDM U-Nets; qkv channels are [head][q|k|v][d]."""
    b, c, hh, ww = x.shape
    d = cfg['num_head_channels'] if cfg.get('num_head_channels', 0) > 0 else c // cfg['num_heads']
    heads = c // d
    xf = x.reshape(b, c, -1)
    qkv = F.conv1d(_gn(sd, p + '.norm', xf, 1e-5), sd[p + '.qkv.weight'], sd[p + '.qkv.bias'])
    bs, width, length = qkv.shape
    ch = width // (3 * heads)
    q, k, v = qkv.reshape(bs * heads, ch * 3, length).split(ch, dim=1)
    scale = 1 / math.sqrt(math.sqrt(ch))
    weight = torch.einsum('bct,bcs->bts', q * scale, k * scale)
    weight = torch.softmax(weight.float(), dim=-1)
    a = torch.einsum('bts,bcs->bct', weight, v).reshape(bs, -1, length)
    h = F.conv1d(a, sd[p + '.proj_out.weight'], sd[p + '.proj_out.bias'])
    return (xf + h).reshape(b, c, hh, ww)


def plan(cfg):
    """Block list mirroring UNetModel.__init__ (openaimodel.py:516-686).

    Returns (input_blocks, middle, output_blocks); each block is a list of ('conv'|'res'|'st'|'down'|'up').
    """
    mc, mult, nrb, ar = cfg['model_channels'], cfg['channel_mult'], cfg['num_res_blocks'], cfg['attention_resolutions']
    inp = [['conv']]
    ds = 1
    for level in range(len(mult)):
        for _ in range(nrb):
            inp.append(['res', 'st'] if ds in ar else ['res'])
        if level != len(mult) - 1:
            inp.append(['down'])
            ds *= 2
    mid = ['res', 'st', 'res']
    out = []
    for level in reversed(range(len(mult))):
        for i in range(nrb + 1):
            layers = ['res'] + (['st'] if ds in ar else [])
            if level and i == nrb:
                layers.append('up')
                ds //= 2
            out.append(layers)
    return inp, mid, out


def unet_forward(sd, cfg, x, timesteps, context, prefix=''):
    """x [B,Cin,h,w] fp32, timesteps [B] (long or float), context [B,77,D] -> [B,Cout,h,w]."""
    P = prefix
    heads = cfg.get('num_heads', 0)
    inp, mid, outb = plan(cfg)
    emb = _lin(sd, P + 'time_embed.2', F.silu(_lin(sd, P + 'time_embed.0', timestep_embedding(timesteps, cfg['model_channels']))))

    def run(block, bp, h):
        for li, kind in enumerate(block):
            p = f'{bp}.{li}'
            if kind == 'conv':
                h = _conv(sd, p, h)
            elif kind == 'res':
                h = _resblock(sd, p, h, emb)
            elif kind == 'st':
                h = _spatial_transformer(sd, p, h, context, heads) if cfg.get('context_dim', 0) else _attention_block(sd, p, h, cfg)
            elif kind == 'down':
                h = _conv(sd, p + '.op', h, stride=2, padding=1)
            elif kind == 'up':
                h = _conv(sd, p + '.conv', F.interpolate(h, scale_factor=2, mode='nearest'))
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
    h = F.silu(_gn(sd, P + 'out.0', h, 1e-5))
    return _conv(sd, P + 'out.2', h)
