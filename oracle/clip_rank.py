"""CPU restatement of Directional-CLIP ranking and the text-task metrics -- TEST INFRASTRUCTURE ONLY.

  model/energy/clean_clip.py:7-41   DirectionalCLIP.__call__: preprocess, encode_image / encode_text (OpenAI CLIP ViT-B/32), scores
  evaluation/utils.py:13-66         calculate_ssim / ssim / calculate_psnr;  evaluation/translate_text.py:76-89 the call site

CLIP itself is third-party (``clip`` @ git+openai/CLIP, README.md:82; absent here).  The arithmetic follows the published model
(clip/model.py VisionTransformer / encode_text), which transformers' CLIPModel restates under HF key names; the fixtures
tests/golden/clip_rank_*.npz are generated from the installed transformers CLIPModel (random-init reduced configs) -- parity for the
real OpenAI weights is unpinned.  Preprocessing: the reference applies torchvision Resize(bicubic) + CenterCrop to a float TENSOR
batch (clean_clip.py:14-17); in the torchvision release of its environment that is F.interpolate(mode='bicubic', align_corners=False)
without antialiasing.  The metric functions are pinned against the reference's own evaluation/utils.py (importable here: numpy + cv2)."""
import numpy as np
import torch
import torch.nn.functional as F

from .clip_text import text_forward

MEAN = torch.tensor([0.48145466, 0.4578275, 0.40821073]).view(1, 3, 1, 1)      # clip.py _transform
STD = torch.tensor([0.26862954, 0.26130258, 0.27577711]).view(1, 3, 1, 1)


def preprocess(img01, size):
    assert img01.shape[2] == img01.shape[3], 'square batches'
    x = F.interpolate(img01, size=(size, size), mode='bicubic', align_corners=False)
    return (x - MEAN) / STD


def image_features(sd, cfg, pixels):
    """CLIPVisionTransformer + visual_projection (HF modeling_clip.py) == clip/model.py VisionTransformer.forward."""
    V = 'vision_model.'
    W, H, P = cfg['width'], cfg['heads'], cfg['patch']
    d = W // H
    B = pixels.shape[0]
    x = F.conv2d(pixels, sd[V + 'embeddings.patch_embedding.weight'], stride=P).flatten(2).transpose(1, 2)          # [B, N, W]
    x = torch.cat([sd[V + 'embeddings.class_embedding'].expand(B, 1, W), x], dim=1) + sd[V + 'embeddings.position_embedding.weight'][None]
    L = x.shape[1]
    x = F.layer_norm(x, (W,), sd[V + 'pre_layrnorm.weight'], sd[V + 'pre_layrnorm.bias'], 1e-5)
    for l in range(cfg['layers']):
        p = f'{V}encoder.layers.{l}'
        lin = lambda t, n: F.linear(t, sd[f'{p}.{n}.weight'], sd[f'{p}.{n}.bias'])
        h = F.layer_norm(x, (W,), sd[f'{p}.layer_norm1.weight'], sd[f'{p}.layer_norm1.bias'], 1e-5)
        q = lin(h, 'self_attn.q_proj') * d ** -0.5
        k, v = lin(h, 'self_attn.k_proj'), lin(h, 'self_attn.v_proj')
        sp = lambda t: t.view(B, L, H, d).transpose(1, 2)
        w = (sp(q) @ sp(k).transpose(-1, -2)).softmax(-1)
        a = (w @ sp(v)).transpose(1, 2).reshape(B, L, W)
        x = x + lin(a, 'self_attn.out_proj')
        h = F.layer_norm(x, (W,), sd[f'{p}.layer_norm2.weight'], sd[f'{p}.layer_norm2.bias'], 1e-5)
        h = lin(h, 'mlp.fc1')
        h = h * torch.sigmoid(1.702 * h)
        x = x + lin(h, 'mlp.fc2')
    pooled = F.layer_norm(x[:, 0], (W,), sd[V + 'post_layernorm.weight'], sd[V + 'post_layernorm.bias'], 1e-5)
    return F.linear(pooled, sd['visual_projection.weight'])


def text_features(sd, cfg, ids):
    """clip/model.py encode_text: final-LN states, EOT token (argmax of the ids), @ text_projection."""
    hs = text_forward(sd, cfg, ids)
    pooled = hs[torch.arange(ids.shape[0]), ids.argmax(dim=-1)]
    return F.linear(pooled, sd['text_projection.weight'])


def dclip_scores(img_f, orig_f, enc_f, dec_f):
    """clean_clip.py:24-39."""
    n = lambda t: t / t.norm(dim=-1, keepdim=True)
    img_f, orig_f, enc_f, dec_f = n(img_f), n(orig_f), n(enc_f), n(dec_f)
    img_dir, txt_dir = n(img_f - orig_f), n(dec_f - enc_f)
    return torch.einsum('bz,bz->b', img_f, dec_f), torch.einsum('bz,bz->b', img_dir, txt_dir)


def directional_clip(sd, vcfg, tcfg, img, original_img, enc_ids, dec_ids):
    f = lambda im: image_features(sd, vcfg, preprocess(im, vcfg['image_size']))
    return dclip_scores(f(img), f(original_img), text_features(sd, tcfg, enc_ids), text_features(sd, tcfg, dec_ids))


# ---- metrics (numpy restatement of evaluation/utils.py; cv2.filter2D on the valid region == correlation with the outer product)
def _gauss(n=11, sigma=1.5):
    k = np.exp(-((np.arange(n) - (n - 1) / 2) ** 2) / (2 * sigma ** 2))
    return k / k.sum()


def ssim(img1, img2):
    """evaluation/utils.py:35-57 (img1, img2: [H, W] in [0, 255])."""
    C1, C2 = (0.01 * 255) ** 2, (0.03 * 255) ** 2
    img1, img2 = img1.astype(np.float64), img2.astype(np.float64)
    win = np.outer(_gauss(), _gauss())

    def filt(x):
        v = np.lib.stride_tricks.sliding_window_view(x, (11, 11))
        return np.einsum('ijkl,kl->ij', v, win)
    mu1, mu2 = filt(img1), filt(img2)
    s1, s2, s12 = filt(img1 ** 2) - mu1 ** 2, filt(img2 ** 2) - mu2 ** 2, filt(img1 * img2) - mu1 * mu2
    return (((2 * mu1 * mu2 + C1) * (2 * s12 + C2)) / ((mu1 ** 2 + mu2 ** 2 + C1) * (s1 + s2 + C2))).mean()


def metrics(img, original_img):
    """evaluation/translate_text.py:76-89 for one pair of [3,H,W] tensors in [0,1] (clamped here as there) -> psnr, ssim, l2."""
    img, original_img = img.clamp(0, 1), original_img.clamp(0, 1)
    mse = ((img - original_img) ** 2).mean(2).mean(1).mean(0)
    psnr = 100.0 if mse == 0 else float(10 * torch.log10(1 / mse))
    a, b = (img.numpy() * 255).transpose(1, 2, 0), (original_img.numpy() * 255).transpose(1, 2, 0)
    s = float(np.mean([ssim(a[:, :, i], b[:, :, i]) for i in range(3)]))
    l2 = float(torch.sqrt(((img - original_img) ** 2).sum(2).sum(1).sum(0)))
    return psnr, s, l2
