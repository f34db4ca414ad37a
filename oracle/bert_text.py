"""CPU restatement of the LDM `BERTEmbedder` transformer -- TEST INFRASTRUCTURE ONLY.

BERTEmbedder.forward (ldm/modules/encoders/modules.py:92-98) = TransformerWrapper(tokens, return_embeddings=True) with
attn_layers = Encoder(dim=n_embed, depth=n_layer) (modules.py:88-90).  Follows ldm/modules/x_transformer.py:
TransformerWrapper.forward 598-626 (token_emb + AbsolutePositionalEmbedding 25-36, final LayerNorm), AttentionLayers.forward
481-523 (pre-norm, layer types ('a','f') * depth, Residual 163-165), Attention 215-266 / 268-368 (bias-free to_q/k/v with
heads=8, dim_head=64, scale dim_head^-1/2, softmax, to_out Linear), FeedForward 194-211 (Linear, exact GELU, Linear).
Pinned by tests/golden/bert_text.npz, generated from the reference's own x_transformer module."""
import torch
import torch.nn.functional as F


def text_forward(sd, cfg, tokens, prefix=''):
    T = prefix + 'transformer.'
    W, H, dh = cfg['width'], cfg['heads'], cfg['dim_head']
    B, L = tokens.shape
    x = sd[T + 'token_emb.weight'][tokens] + sd[T + 'pos_emb.emb.weight'][:L][None]
    for l in range(cfg['layers']):
        pa, pf = f'{T}attn_layers.layers.{2 * l}', f'{T}attn_layers.layers.{2 * l + 1}'
        h = F.layer_norm(x, (W,), sd[pa + '.0.weight'], sd[pa + '.0.bias'], 1e-5)
        sp = lambda t: t.view(B, L, H, dh).transpose(1, 2)
        q, k, v = (sp(F.linear(h, sd[f'{pa}.1.to_{n}.weight'])) for n in 'qkv')
        w = (torch.einsum('bhid,bhjd->bhij', q, k) * dh ** -0.5).softmax(-1)
        a = torch.einsum('bhij,bhjd->bhid', w, v).transpose(1, 2).reshape(B, L, H * dh)
        x = x + F.linear(a, sd[pa + '.1.to_out.weight'], sd[pa + '.1.to_out.bias'])
        h = F.layer_norm(x, (W,), sd[pf + '.0.weight'], sd[pf + '.0.bias'], 1e-5)
        h = F.gelu(F.linear(h, sd[pf + '.1.net.0.0.weight'], sd[pf + '.1.net.0.0.bias']))
        x = x + F.linear(h, sd[pf + '.1.net.2.weight'], sd[pf + '.1.net.2.bias'])
    return F.layer_norm(x, (W,), sd[T + 'norm.weight'], sd[T + 'norm.bias'], 1e-5)
