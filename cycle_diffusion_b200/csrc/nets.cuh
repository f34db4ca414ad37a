// nets.cuh -- network objects: parameter inventory, packed weight blob, forward executors.
#pragma once
#include <string>
#include <unordered_map>
#include <vector>

#include "common.cuh"

namespace cdx {

struct Param {
  std::string name;
  int64_t dims[4] = {0, 0, 0, 0};
  int rank = 0;
  size_t numel = 0;
  size_t off = 0;        // float offset into the blob
  int segment = 0;       // 0 general, 1 emb-proj weights (concatenated), 2 emb-proj biases (concatenated)
  bool conv3 = false;    // stored repacked O,kh,kw,I
  int cin_pad = 0;       // 3x3 conv with < 32 input channels: stored with Cin zero-padded to this (tensor-core K blocks are 32 wide)
  size_t store() const { return cin_pad ? (size_t)dims[0] * 9 * cin_pad : numel; }   // floats occupied in the blob
  bool geglu = false;    // GEGLU projection: rows stored as alternating blocks of 32 value rows | their 32 gate rows
  bool loaded = false;
};

enum NetKind { NET_UNET_OPENAI = 1, NET_UNET_IDDPM = 2, NET_VAE = 3, NET_CLIP_TEXT = 4, NET_UNET_DDPM = 5 };

struct Net {
  Engine* eng = nullptr;
  int kind = 0;
  cdx_unet_config ucfg{};
  cdx_vae_config vcfg{};
  cdx_text_config tcfg{};
  std::vector<Param> params;
  std::unordered_map<std::string, int> index;
  float* blob = nullptr;
  float* blob_hi = nullptr;     // rn_tf32(blob)            } pre-split planes for the tcgen05 TS kernel,
  float* blob_lo = nullptr;     // rn_tf32(blob - blob_hi)  } same offsets as `blob`, derived at finalize
  bool planes_valid = false;
  // fp16-split planes (MODE_H16): hi = fp16(w * 2^w_exp), lo = fp16(w * 2^w_exp - hi), element index = float index into `blob`;
  // w_exp from the largest |weight| of the GEMM operands (rank >= 2 parameters) so that every scaled weight is < 2^15
  void* blob_h_hi = nullptr;
  void* blob_h_lo = nullptr;
  int w_exp = 0;
  size_t blob_floats = 0;
  bool finalized = false;
  // timestep embedding
  std::vector<float> freqs_host;
  float* freqs_dev = nullptr;
  // cross-attention K / V of a context that stays fixed over a sampling loop (set up by the loop drivers in cabi.cu):
  // computed by the first U-Net call of the loop, reused by the others (the reference recomputes them every step)
  struct CtxKV {
    static constexpr int MAX_LAYERS = 63;
    bool valid = false; const float* ctx = nullptr; int L = 0, B = 0; float* buf = nullptr; size_t cap = 0;
    float* amax = nullptr;     // [0]: max |context|, [1 + layer]: max |V| of that layer's context projection (device)
  } ctxkv;
  // concatenated ResBlock emb projections: weights [emb_rows][ted] at emb_w_off, biases at emb_b_off
  size_t emb_w_off = 0, emb_b_off = 0;
  int emb_rows = 0, ted = 0;
  std::unordered_map<std::string, int> emb_off;   // ResBlock prefix -> row offset

  const Param& param(const std::string& name) const;
  float* P(const std::string& name) const { return blob + param(name).off; }
  bool has(const std::string& name) const { return index.find(name) != index.end(); }
  int dim0(const std::string& name) const { return (int)param(name).dims[0]; }
};

Net* make_unet(Engine* e, const cdx_unet_config& cfg);
Net* make_vae(Engine* e, const cdx_vae_config& cfg);
Net* make_text(Engine* e, const cdx_text_config& cfg);
void destroy_net(Net* n);
void net_load_param(Net& n, const char* name, const float* data, bool on_device, const int64_t* dims, int rank);
void net_finalize(Net& n);
void net_ensure_blob(Net& n);

// forward executors (enqueue only; caller handles arena dry-run)
// reuse_ctx: the caller guarantees `ctx` is unchanged since the previous call with reuse_ctx (and n.ctxkv was invalidated
// at the start of the loop) -> context K / V projections are taken from n.ctxkv instead of being recomputed
void unet_forward(Net& n, const float* x_nchw, const float* t_dev, const float* ctx, int ctx_len, float* out_nchw, int B, int H,
                  int W, cudaStream_t s, bool reuse_ctx = false);
void vae_encode(Net& n, const float* img_nchw, float* moments_nchw, int B, int R, cudaStream_t s);
void vae_decode(Net& n, const float* z_nchw, float* img_nchw, int B, int h, cudaStream_t s);
void text_encode(Net& n, const int* ids, float* out, int B, int L, cudaStream_t s);
void text_features(Net& n, const int* ids, float* out, int B, int L, cudaStream_t s);                // CLIP.encode_text   -> [B, proj_dim]
void clip_image_features(Net& n, const float* pixels, float* out, int B, cudaStream_t s);            // CLIP.encode_image  -> [B, proj_dim]

}  // namespace cdx
