"""Oracle: KL-f8 autoencoder encode / decode, functional fp32 restatement (test infrastructure only).

Follows:
  Encoder.forward              ref ldm/modules/diffusionmodules/model.py:434-459 (built at :368-432)
  Decoder.forward              ref model.py:535-568 (built at :462-533; ``up`` stored low-res-last, :525)
  ResnetBlock.forward          ref model.py:121-141 (temb is None)
  AttnBlock.forward            ref model.py:178-202 (single head, scale C^-1/2, 1x1 convs with bias)
  Downsample (asym pad 0,1,0,1; stride 2, pad 0)   ref model.py:72-76
  Upsample (nearest x2 + conv3x3)                  ref model.py:53-57
  Normalize (GroupNorm 32, eps 1e-6)               ref model.py:38-39
  AutoencoderKL.encode/decode (quant_conv / post_quant_conv 1x1)   ref ldm/models/autoencoder.py:302-303, 324-333
State-dict keys: ``encoder.*``, ``decoder.*``, ``quant_conv.*``, ``post_quant_conv.*``.
"""
import torch
import torch.nn.functional as F


def default_kl_f8_config():
    """v1-inference.yaml:51-65 ddconfig."""
    return dict(ch=128, ch_mult=(1, 2, 4, 4), num_res_blocks=2, in_channels=3, out_ch=3, z_channels=4, embed_dim=4)


def _gn(sd, p, x):
    return F.group_norm(x, 32, sd[p + '.weight'], sd[p + '.bias'], 1e-6)


def _conv(sd, p, x, stride=1, padding=1):
    return F.conv2d(x, sd[p + '.weight'], sd[p + '.bias'], stride=stride, padding=padding)


def _swish(x):
    return x * torch.sigmoid(x)


def _resnet(sd, p, x):
    h = _conv(sd, p + '.conv1', _swish(_gn(sd, p + '.norm1', x)))
    h = _conv(sd, p + '.conv2', _swish(_gn(sd, p + '.norm2', h)))
    if (p + '.nin_shortcut.weight') in sd:
        x = _conv(sd, p + '.nin_shortcut', x, padding=0)
    return x + h


def _attn(sd, p, x):
    h_ = _gn(sd, p + '.norm', x)
    q = _conv(sd, p + '.q', h_, padding=0)
    k = _conv(sd, p + '.k', h_, padding=0)
    v = _conv(sd, p + '.v', h_, padding=0)
    b, c, h, w = q.shape
    q = q.reshape(b, c, h * w).permute(0, 2, 1)
    k = k.reshape(b, c, h * w)
    w_ = torch.bmm(q, k) * (int(c) ** (-0.5))
    w_ = F.softmax(w_, dim=2)
    v = v.reshape(b, c, h * w)
    h_ = torch.bmm(v, w_.permute(0, 2, 1)).reshape(b, c, h, w)
    return x + _conv(sd, p + '.proj_out', h_, padding=0)


def encode_moments(sd, cfg, x, prefix=''):
    """x [B,3,R,R] in [-1,1] -> moments [B, 2*embed_dim, R/8, R/8] (mean | logvar)."""
    P = prefix + 'encoder.'
    nres = len(cfg['ch_mult'])
    h = _conv(sd, P + 'conv_in', x)
    for lvl in range(nres):
        for blk in range(cfg['num_res_blocks']):
            h = _resnet(sd, f'{P}down.{lvl}.block.{blk}', h)
        if lvl != nres - 1:
            h = _conv(sd, f'{P}down.{lvl}.downsample.conv', F.pad(h, (0, 1, 0, 1)), stride=2, padding=0)
    h = _resnet(sd, P + 'mid.block_1', h)
    h = _attn(sd, P + 'mid.attn_1', h)
    h = _resnet(sd, P + 'mid.block_2', h)
    h = _conv(sd, P + 'conv_out', _swish(_gn(sd, P + 'norm_out', h)))
    return _conv(sd, prefix + 'quant_conv', h, padding=0)


def vq_quantize(sd, z, prefix=''):
    """taming.modules.vqvae.quantize.VectorQuantizer2.forward (third party, absent from the tree; the published algorithm): nearest
    code by  |z|^2 + |e|^2 - 2 z.e , straight-through value  z + (z_q - z).  z [B, e_dim, h, w] -> same shape."""
    emb = sd[prefix + 'quantize.embedding.weight']
    zp = z.permute(0, 2, 3, 1).contiguous()
    zf = zp.view(-1, emb.shape[1])
    d = torch.sum(zf ** 2, dim=1, keepdim=True) + torch.sum(emb ** 2, dim=1) - 2 * torch.einsum('bd,dn->bn', zf, emb.t())
    idx = torch.argmin(d, dim=1)
    zq = emb[idx].view(zp.shape)
    zq = zp + (zq - zp)
    return zq.permute(0, 3, 1, 2).contiguous()


def decode(sd, cfg, z, prefix=''):
    """z [B,embed_dim,h,w] (already divided by scale_factor) -> image [B,3,8h,8w].  cfg['vq']: VQModelInterface.decode
    (ldm/models/autoencoder.py:272-281) quantises first."""
    P = prefix + 'decoder.'
    nres = len(cfg['ch_mult'])
    if cfg.get('vq'):
        z = vq_quantize(sd, z, prefix)
    h = _conv(sd, prefix + 'post_quant_conv', z, padding=0)
    h = _conv(sd, P + 'conv_in', h)
    h = _resnet(sd, P + 'mid.block_1', h)
    h = _attn(sd, P + 'mid.attn_1', h)
    h = _resnet(sd, P + 'mid.block_2', h)
    for lvl in reversed(range(nres)):
        for blk in range(cfg['num_res_blocks'] + 1):
            h = _resnet(sd, f'{P}up.{lvl}.block.{blk}', h)
        if lvl != 0:
            h = _conv(sd, f'{P}up.{lvl}.upsample.conv', F.interpolate(h, scale_factor=2.0, mode='nearest'))
    return _conv(sd, P + 'conv_out', _swish(_gn(sd, P + 'norm_out', h)))
