// kernels_tc.cu -- tcgen05 / TMEM / TMA back end for the dense contractions (3xTF32 split, fp32-faithful).
// Placeholder until the kernel lands: reports "not eligible" so that every contraction takes the FFMA path.
#include "common.cuh"

namespace cdx {
bool gemm_tc(Engine&, const GemmArgs&, cudaStream_t) { return false; }
}  // namespace cdx
