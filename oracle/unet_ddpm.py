"""Oracle: Ho et al. DDPM pixel U-Net forward (CelebA-HQ / LSUN checkpoints), functional fp32 restatement -- TEST INFRASTRUCTURE ONLY.

Follows model/lib/ddpm_ddim/models/ddpm/diffusion.py:
  get_timestep_embedding   :6-25   ([sin | cos], log(10000) / (half - 1))
  ResnetBlock.forward      :117-139
  AttnBlock.forward        :168-189 (single head, scale C^-1/2)
  Downsample / Upsample    :36-70   (asymmetric pad (0,1,0,1) stride-2 conv; nearest x2 + conv)
  DDPM.forward             :299-337 (topology :192-297)
Selected by DDPMDDIMWrapper when config.data.dataset is CelebA_HQ or LSUN (ddpm_ddim_wrapper.py:360-369).  cfg['attention_resolutions']
holds downsample factors (image_size / attn resolution).  Pinned by tests/golden/unet_ddpm.npz (the reference class itself)."""
import math
import torch
import torch.nn.functional as F


def timestep_embedding(timesteps, dim):
    half = dim // 2
    emb = math.log(10000) / (half - 1)
    emb = torch.exp(torch.arange(half, dtype=torch.float32) * -emb)
    emb = timesteps.float()[:, None] * emb[None, :]
    return torch.cat([torch.sin(emb), torch.cos(emb)], dim=1)


def _gn(sd, p, x):
    return F.group_norm(x, 32, sd[p + '.weight'], sd[p + '.bias'], 1e-6)


def _conv(sd, p, x, stride=1, padding=1):
    return F.conv2d(x, sd[p + '.weight'], sd[p + '.bias'], stride=stride, padding=padding)


def _swish(x):
    return x * torch.sigmoid(x)


def _res(sd, p, x, temb):
    h = _conv(sd, p + '.conv1', _swish(_gn(sd, p + '.norm1', x)))
    h = h + F.linear(_swish(temb), sd[p + '.temb_proj.weight'], sd[p + '.temb_proj.bias'])[:, :, None, None]
    h = _conv(sd, p + '.conv2', _swish(_gn(sd, p + '.norm2', h)))
    if (p + '.nin_shortcut.weight') in sd:
        x = _conv(sd, p + '.nin_shortcut', x, padding=0)
    return x + h


def _attn(sd, p, x):
    h_ = _gn(sd, p + '.norm', x)
    q, k, v = (_conv(sd, f'{p}.{n}', h_, padding=0) for n in ('q', 'k', 'v'))
    b, c, h, w = q.shape
    w_ = torch.bmm(q.reshape(b, c, h * w).permute(0, 2, 1), k.reshape(b, c, h * w)) * (int(c) ** (-0.5))
    w_ = F.softmax(w_, dim=2)
    h_ = torch.bmm(v.reshape(b, c, h * w), w_.permute(0, 2, 1)).reshape(b, c, h, w)
    return x + _conv(sd, p + '.proj_out', h_, padding=0)


def unet_forward(sd, cfg, x, t):
    ch, mult, nrb, ar = cfg['model_channels'], cfg['channel_mult'], cfg['num_res_blocks'], cfg['attention_resolutions']
    temb = timestep_embedding(t, ch)
    temb = F.linear(_swish(F.linear(temb, sd['temb.dense.0.weight'], sd['temb.dense.0.bias'])), sd['temb.dense.1.weight'], sd['temb.dense.1.bias'])
    hs = [_conv(sd, 'conv_in', x)]
    ds = 1
    L = len(mult)
    for lvl in range(L):
        for b in range(nrb):
            h = _res(sd, f'down.{lvl}.block.{b}', hs[-1], temb)
            if ds in ar:
                h = _attn(sd, f'down.{lvl}.attn.{b}', h)
            hs.append(h)
        if lvl != L - 1:
            hs.append(_conv(sd, f'down.{lvl}.downsample.conv', F.pad(hs[-1], (0, 1, 0, 1)), stride=2, padding=0))
            ds *= 2
    h = hs[-1]
    h = _res(sd, 'mid.block_1', h, temb)
    h = _attn(sd, 'mid.attn_1', h)
    h = _res(sd, 'mid.block_2', h, temb)
    for lvl in reversed(range(L)):
        for b in range(nrb + 1):
            h = _res(sd, f'up.{lvl}.block.{b}', torch.cat([h, hs.pop()], dim=1), temb)
            if ds in ar:
                h = _attn(sd, f'up.{lvl}.attn.{b}', h)
        if lvl != 0:
            h = _conv(sd, f'up.{lvl}.upsample.conv', F.interpolate(h, scale_factor=2.0, mode='nearest'))
            ds //= 2
    return _conv(sd, 'conv_out', _swish(_gn(sd, 'norm_out', h)))
