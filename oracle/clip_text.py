"""CPU restatement of the CLIP text tower the reference conditions Stable Diffusion on -- TEST INFRASTRUCTURE ONLY.

The reference calls a third-party model: ``FrozenCLIPEmbedder`` (ldm/modules/encoders/modules.py:140-158) tokenises and runs
HF ``CLIPTextModel("openai/clip-vit-large-patch14")`` (transformers==4.19.2, environment.yml:466) and returns
``last_hidden_state``.  The arithmetic below follows transformers' modeling_clip.py: CLIPTextEmbeddings (token + position
embedding), CLIPEncoderLayer (pre-LN; CLIPAttention scales q by d^-1/2 and adds a causal mask; CLIPMLP with quick-GELU),
final_layer_norm.  Pinned by tests/golden/clip_text_*.npz, generated from the installed transformers CLIPTextModel
(tests/golden/make_golden.py: golden_clip_text)."""
import torch
import torch.nn.functional as F


def text_forward(sd, cfg, input_ids, prefix=''):
    """input_ids [B, L] long -> last_hidden_state [B, L, width] (fp32)."""
    T = prefix + 'text_model.'
    W, H = cfg['width'], cfg['heads']
    d = W // H
    B, L = input_ids.shape
    x = sd[T + 'embeddings.token_embedding.weight'][input_ids] + sd[T + 'embeddings.position_embedding.weight'][:L][None]
    causal = torch.full((L, L), float('-inf')).triu(1)                 # query i sees keys 0..i
    for l in range(cfg['layers']):
        p = f'{T}encoder.layers.{l}'
        lin = lambda t, n: F.linear(t, sd[f'{p}.{n}.weight'], sd[f'{p}.{n}.bias'])
        h = F.layer_norm(x, (W,), sd[f'{p}.layer_norm1.weight'], sd[f'{p}.layer_norm1.bias'], 1e-5)
        q = lin(h, 'self_attn.q_proj') * d ** -0.5
        k, v = lin(h, 'self_attn.k_proj'), lin(h, 'self_attn.v_proj')
        sp = lambda t: t.view(B, L, H, d).transpose(1, 2)
        w = (sp(q) @ sp(k).transpose(-1, -2) + causal).softmax(-1)
        a = (w @ sp(v)).transpose(1, 2).reshape(B, L, W)
        x = x + lin(a, 'self_attn.out_proj')
        h = F.layer_norm(x, (W,), sd[f'{p}.layer_norm2.weight'], sd[f'{p}.layer_norm2.bias'], 1e-5)
        h = lin(h, 'mlp.fc1')
        h = h * torch.sigmoid(1.702 * h)                                # quick_gelu
        x = x + lin(h, 'mlp.fc2')
    return F.layer_norm(x, (W,), sd[T + 'final_layer_norm.weight'], sd[T + 'final_layer_norm.bias'], 1e-5)
