// kernels_norm.cu -- HBM-bound normalisation kernels: GroupNorm(32)(+scale-shift)(+SiLU), LayerNorm, row softmax.
//
// GroupNorm follows GroupNorm32 / Normalize (util.py:215-217 eps 1e-5; attention.py:76-77 and model.py:38-39 eps 1e-6)
// on NHWC activations, optionally over the channel concatenation of two tensors (the U-Net skip `th.cat`, OAI:736)
// so that the concat is never materialised before the norm.  Statistics are accumulated in fp64 (one read),
// then one read + one write applies  y = (x - mean) * rstd * gamma + beta  [ * (1+scale) + shift ] [ SiLU ].
// Algorithmic HBM bytes: 2 reads + 1 write of the activation (stats pass + apply pass; the apply blocks fold the
// per-chunk partial sums themselves).
#include "common.cuh"

namespace cdx {
namespace {

constexpr int GN_GROUPS = 32;

__device__ __forceinline__ float silu_f(float v) { return v / (1.f + expf(-v)); }

// partial sums: part[((b*nchunk + chunk)*32 + g)*2 + {0,1}]
__global__ void __launch_bounds__(256) gn_stats_kernel(const float* __restrict__ x1, int C1, const float* __restrict__ x2,
                                                       int C2, int HW, int rows_per_chunk, double* __restrict__ part) {
  const int C = C1 + C2;
  const int cpg = C / GN_GROUPS;
  const int b = blockIdx.y, chunk = blockIdx.x;
  __shared__ double ssum[GN_GROUPS], ssq[GN_GROUPS];
  if (threadIdx.x < GN_GROUPS) { ssum[threadIdx.x] = 0.0; ssq[threadIdx.x] = 0.0; }
  __syncthreads();
  const int r0 = chunk * rows_per_chunk;
  const int r1 = min(HW, r0 + rows_per_chunk);
  // thread (tr, tc): tc owns float4 channel slots tc, tc+ncol, ...; tr strides over the rows of the chunk.  A float4 may
  // straddle two groups when cpg % 4 != 0 (C=320 -> cpg=10), so each of its 4 lanes accumulates separately and is
  // flushed to its own group's shared accumulator once per slot.
  const int C4 = C >> 2;
  const int ncol = min(C4, (int)blockDim.x);
  const int nrow_par = blockDim.x / ncol;
  const int tr = threadIdx.x / ncol, tc = threadIdx.x - tr * ncol;
  if (tr < nrow_par) {
    for (int c4 = tc; c4 < C4; c4 += ncol) {
      const int c = c4 * 4;
      double s[4] = {0.0, 0.0, 0.0, 0.0}, q[4] = {0.0, 0.0, 0.0, 0.0};
      const float* base = (c < C1) ? (x1 + (long long)b * HW * C1 + c) : (x2 + (long long)b * HW * C2 + (c - C1));
      const int ldx = (c < C1) ? C1 : C2;
      int r = r0 + tr;
      for (; r + 3 * nrow_par < r1; r += 4 * nrow_par) {      // 4 independent 128-bit loads in flight per thread
        float4 v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = *reinterpret_cast<const float4*>(base + (long long)(r + u * nrow_par) * ldx);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const double d0 = v[u].x, d1 = v[u].y, d2 = v[u].z, d3 = v[u].w;
          s[0] += d0; q[0] += d0 * d0;
          s[1] += d1; q[1] += d1 * d1;
          s[2] += d2; q[2] += d2 * d2;
          s[3] += d3; q[3] += d3 * d3;
        }
      }
      for (; r < r1; r += nrow_par) {
        const float4 v = *reinterpret_cast<const float4*>(base + (long long)r * ldx);
        const double d0 = v.x, d1 = v.y, d2 = v.z, d3 = v.w;
        s[0] += d0; q[0] += d0 * d0;
        s[1] += d1; q[1] += d1 * d1;
        s[2] += d2; q[2] += d2 * d2;
        s[3] += d3; q[3] += d3 * d3;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int g = (c + j) / cpg;
        atomicAdd(&ssum[g], s[j]);
        atomicAdd(&ssq[g], q[j]);
      }
    }
  }
  __syncthreads();
  if (threadIdx.x < GN_GROUPS) {
    double* o = part + (((long long)b * gridDim.x + chunk) * GN_GROUPS + threadIdx.x) * 2;
    o[0] = ssum[threadIdx.x];
    o[1] = ssq[threadIdx.x];
  }
}

// grid (row chunks, B); thread (tr, tc) owns channel vectors tc, tc+ncol, ... (so group / affine coefficients are hoisted out
// of the row loop as y = x * sc + sh, the form ATen's CPU kernel uses) and walks the chunk's rows 4 at a time
__global__ void __launch_bounds__(256) gn_apply_kernel(const float* __restrict__ x1, int C1, const float* __restrict__ x2, int C2,
                                                       const float* __restrict__ gamma, const float* __restrict__ beta,
                                                       const double* __restrict__ part, int nchunk, double inv_count,
                                                       float eps, int silu,
                                                       const float* __restrict__ scale, const float* __restrict__ shift,
                                                       int ld_ss, float* __restrict__ y, int HW, int rows_per_chunk) {
  const int C = C1 + C2;
  const int cpg = C / GN_GROUPS;
  const int C4 = C >> 2;
  const int b = blockIdx.y;
  const int r0 = blockIdx.x * rows_per_chunk;
  const int r1 = min(HW, r0 + rows_per_chunk);
  const int ncol = min(C4, (int)blockDim.x);
  const int nrow_par = blockDim.x / ncol;
  const int tr = threadIdx.x / ncol, tc = threadIdx.x - tr * ncol;
  // every block folds the stats partials of its image itself (fixed order, fp64): saves a launch per GroupNorm
  __shared__ float mean_rstd[GN_GROUPS * 2];
  if (threadIdx.x < GN_GROUPS) {
    const int g = threadIdx.x;
    double s = 0.0, q = 0.0;
    for (int c = 0; c < nchunk; ++c) {
      const double* o = part + (((long long)b * nchunk + c) * GN_GROUPS + g) * 2;
      s += o[0];
      q += o[1];
    }
    const double mean = s * inv_count;
    double var = q * inv_count - mean * mean;
    if (var < 0.0) var = 0.0;
    mean_rstd[g * 2 + 0] = (float)mean;
    mean_rstd[g * 2 + 1] = (float)(1.0 / sqrt(var + (double)eps));
  }
  __syncthreads();
  if (tr >= nrow_par) return;
  for (int c4 = tc; c4 < C4; c4 += ncol) {
    const int c = c4 * 4;
    float sc[4], sh[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int g = (c + j) / cpg;
      const float mean = mean_rstd[g * 2 + 0];
      const float rstd = mean_rstd[g * 2 + 1];
      float a = rstd * gamma[c + j];
      float o = beta[c + j] - mean * a;
      if (scale) {      // gn(x) * (1 + scale) + shift   (improved-DDPM scale-shift norm)
        const float s1 = 1.f + scale[(long long)b * ld_ss + c + j];
        a *= s1;
        o = o * s1 + shift[(long long)b * ld_ss + c + j];
      }
      sc[j] = a; sh[j] = o;
    }
    const float* src = (c < C1) ? (x1 + (long long)b * HW * C1 + c) : (x2 + (long long)b * HW * C2 + (c - C1));
    const int ldx = (c < C1) ? C1 : C2;
    float* dst = y + (long long)b * HW * C + c;
    auto act = [&](float4 v) {
      float t[4] = {fmaf(v.x, sc[0], sh[0]), fmaf(v.y, sc[1], sh[1]), fmaf(v.z, sc[2], sh[2]), fmaf(v.w, sc[3], sh[3])};
      if (silu) {
#pragma unroll
        for (int j = 0; j < 4; ++j) t[j] = __fdividef(t[j], 1.f + __expf(-t[j]));
      }
      return make_float4(t[0], t[1], t[2], t[3]);
    };
    int r = r0 + tr;
    for (; r + 3 * nrow_par < r1; r += 4 * nrow_par) {
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) v[u] = *reinterpret_cast<const float4*>(src + (long long)(r + u * nrow_par) * ldx);
#pragma unroll
      for (int u = 0; u < 4; ++u) *reinterpret_cast<float4*>(dst + (long long)(r + u * nrow_par) * C) = act(v[u]);
    }
    for (; r < r1; r += nrow_par)
      *reinterpret_cast<float4*>(dst + (long long)r * C) = act(*reinterpret_cast<const float4*>(src + (long long)r * ldx));
  }
}

// one warp per row
__global__ void __launch_bounds__(256) layernorm_kernel(const float* __restrict__ x, const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, float* __restrict__ y, int M, int C,
                                                        float eps) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= M) return;
  const float* xr = x + (long long)warp * C;
  float* yr = y + (long long)warp * C;
  const int C4 = C >> 2;
  float s = 0.f;
  for (int i = lane; i < C4; i += 32) {
    const float4 v = *reinterpret_cast<const float4*>(xr + i * 4);
    s += (v.x + v.y) + (v.z + v.w);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / (float)C;
  float q = 0.f;
  for (int i = lane; i < C4; i += 32) {
    const float4 v = *reinterpret_cast<const float4*>(xr + i * 4);
    const float a = v.x - mean, b = v.y - mean, c = v.z - mean, d = v.w - mean;
    q += (a * a + b * b) + (c * c + d * d);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float rstd = 1.f / sqrtf(q / (float)C + eps);
  for (int i = lane; i < C4; i += 32) {
    const float4 v = *reinterpret_cast<const float4*>(xr + i * 4);
    const float4 g = *reinterpret_cast<const float4*>(gamma + i * 4);
    const float4 b = *reinterpret_cast<const float4*>(beta + i * 4);
    float4 o;
    o.x = (v.x - mean) * rstd * g.x + b.x;
    o.y = (v.y - mean) * rstd * g.y + b.y;
    o.z = (v.z - mean) * rstd * g.z + b.z;
    o.w = (v.w - mean) * rstd * g.w + b.w;
    *reinterpret_cast<float4*>(yr + i * 4) = o;
  }
}

// in-place softmax over rows of length L (row stride ld); one warp per row, three passes (row stays in L1/L2)
__global__ void __launch_bounds__(256) softmax_kernel(float* __restrict__ x, long long rows, int L0, int ld, int causal_nq) {
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= rows) return;
  float* r = x + warp * ld;
  int L = L0;
  if (causal_nq > 0) {                       // causal mask (CLIP text tower): query i attends to keys 0..i, masked probabilities are 0
    L = min(L0, (int)(warp % causal_nq) + 1);
    for (int i = L + lane; i < L0; i += 32) r[i] = 0.f;
  }
  float mx = -INFINITY;
  for (int i = lane; i < L; i += 32) mx = fmaxf(mx, r[i]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  float s = 0.f;
  for (int i = lane; i < L; i += 32) {
    const float ev = expf(r[i] - mx);
    r[i] = ev;
    s += ev;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float inv = 1.f / s;
  for (int i = lane; i < L; i += 32) r[i] = r[i] * inv;
}

}  // namespace

void groupnorm(Engine& e, const float* x1, int C1, const float* x2, int C2, const float* gamma, const float* beta, float eps,
               bool silu, const float* scale, const float* shift, int ld_ss, float* y, int B, int HW, cudaStream_t s) {
  const int C = C1 + C2;
  CDX_CHECK(C % GN_GROUPS == 0, "groupnorm: C=%d not divisible by 32", C);
  CDX_CHECK(C1 % 4 == 0 && C2 % 4 == 0, "groupnorm: channel counts must be multiples of 4 (C1=%d C2=%d)", C1, C2);
  Scope sc(e.arena);
  int nchunk = cdiv(4LL * e.num_sms, B);
  if (nchunk > HW) nchunk = HW;
  if (nchunk < 1) nchunk = 1;
  const int rows_per_chunk = cdiv(HW, nchunk);
  nchunk = cdiv(HW, rows_per_chunk);
  double* part = (double*)e.arena.alloc((size_t)B * nchunk * GN_GROUPS * 2 * sizeof(double));
  if (e.dry()) return;
  ProfScope ps(e, s, PROF_GROUPNORM, 0.0, 2.0 * 4.0 * B * (double)HW * C, 2);   // algorithmic: one read + one write
  gn_stats_kernel<<<dim3(nchunk, B), 256, 0, s>>>(x1, C1, x2, C2, HW, rows_per_chunk, part);
  int achunk = cdiv(8LL * e.num_sms, B);
  if (achunk > HW) achunk = HW;
  const int arows = cdiv(HW, achunk);
  achunk = cdiv(HW, arows);
  gn_apply_kernel<<<dim3(achunk, B), 256, 0, s>>>(x1, C1, x2, C2, gamma, beta, part, nchunk, 1.0 / ((double)HW * (C / GN_GROUPS)), eps,
                                                  silu ? 1 : 0, scale, shift, ld_ss, y, HW, arows);
  CDX_CUDA(cudaGetLastError());
  e.launches += 2;
}

void layernorm(Engine& e, const float* x, const float* gamma, const float* beta, float* y, int M, int C, cudaStream_t s) {
  CDX_CHECK(C % 4 == 0, "layernorm: C=%d must be a multiple of 4", C);
  if (e.dry()) return;
  ProfScope ps(e, s, PROF_LAYERNORM, 0.0, 2.0 * 4.0 * (double)M * C, 1);
  layernorm_kernel<<<cdiv((long long)M * 32, 256), 256, 0, s>>>(x, gamma, beta, y, M, C, 1e-5f);
  CDX_CUDA(cudaGetLastError());
  e.launches++;
}

void softmax_rows(Engine& e, float* x, long long rows, int L, int ld, cudaStream_t s, int causal_nq) {
  if (e.dry()) return;
  ProfScope ps(e, s, PROF_SOFTMAX, 0.0, 2.0 * 4.0 * (double)rows * L, 1);
  softmax_kernel<<<cdiv(rows * 32, 256), 256, 0, s>>>(x, rows, L, ld, causal_nq);
  CDX_CUDA(cudaGetLastError());
  e.launches++;
}

}  // namespace cdx
