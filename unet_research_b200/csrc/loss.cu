// Fused masked binary-cross-entropy of the training step (reference utils_training.py:28-33 with nn.BCELoss,
// base_model_tests/training.py:195):
//     seg = out * mask;  gt = gt * mask;  loss = mean(BCE(seg, gt)) * numel / count_nonzero(mask)
// as ONE pass over the three fp32 maps (loss sum, mask count and the unscaled gradient d BCE / d out together) plus a
// one-block deterministic reduction, instead of ~10 ATen kernels forward and ~10 more through autograd.
// BCE follows ATen (Loss.cu): log terms clamped at -100; gradient (seg - gt) / max((1 - seg) * seg, 1e-12).
#include "b2u_common.cuh"

namespace b2u {

constexpr int kLossThreads = 256;

__global__ void __launch_bounds__(kLossThreads) masked_bce_partial_kernel(const float* __restrict__ out, const float* __restrict__ gt,
                                                                         const float* __restrict__ mask, long long n,
                                                                         float* __restrict__ grad_unscaled, double* __restrict__ partials) {
  double ls = 0.0;
  unsigned long long nz = 0;
  for (long long i = blockIdx.x * static_cast<long long>(kLossThreads) + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * kLossThreads) {
    const float m = mask[i];
    const float s = out[i] * m;
    const float g = gt[i] * m;
    const float l1 = fmaxf(logf(s), -100.f), l0 = fmaxf(logf(1.f - s), -100.f);
    ls += static_cast<double>(-(g * l1 + (1.f - g) * l0));
    nz += m != 0.f ? 1ull : 0ull;
    // d loss_i / d out_i = d BCE / d seg * mask  (the 1 / numel of the mean and the numel / nnz rescale are applied later)
    grad_unscaled[i] = (s - g) / fmaxf((1.f - s) * s, 1e-12f) * m;
  }
  __shared__ double sh_l[kLossThreads / 32];
  __shared__ unsigned long long sh_n[kLossThreads / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ls += __shfl_xor_sync(0xffffffffu, ls, o);
    nz += __shfl_xor_sync(0xffffffffu, nz, o);
  }
  if ((threadIdx.x & 31) == 0) { sh_l[threadIdx.x >> 5] = ls; sh_n[threadIdx.x >> 5] = nz; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0;
    unsigned long long b = 0;
    for (int w = 0; w < kLossThreads / 32; ++w) { a += sh_l[w]; b += sh_n[w]; }
    partials[2 * blockIdx.x] = a;
    partials[2 * blockIdx.x + 1] = static_cast<double>(b);
  }
}

// loss = (sum / numel) * (numel / nnz) evaluated like the reference: mean first, then the rescale in fp32
__global__ void masked_bce_final_kernel(const double* __restrict__ partials, int blocks, long long n, float* __restrict__ loss,
                                        float* __restrict__ scale_out) {
  double a = 0.0, b = 0.0;
  for (int i = 0; i < blocks; ++i) { a += partials[2 * i]; b += partials[2 * i + 1]; }   // fixed order: deterministic
  const float mean = static_cast<float>(a / static_cast<double>(n));
  const float resc = static_cast<float>(static_cast<double>(n) / b);
  *loss = mean * resc;
  *scale_out = resc / static_cast<float>(n);        // d loss / d (sum of BCE terms)
}

__global__ void __launch_bounds__(kLossThreads) masked_bce_bwd_kernel(const float* __restrict__ grad_unscaled, const float* __restrict__ upstream,
                                                                     const float* __restrict__ scale, float* __restrict__ grad_out, long long n) {
  const float k = *upstream * *scale;
  for (long long i = blockIdx.x * static_cast<long long>(kLossThreads) + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * kLossThreads)
    grad_out[i] = grad_unscaled[i] * k;
}

}  // namespace b2u

using namespace b2u;

extern "C" int b2u_masked_bce_blocks(void) { return 148; }

extern "C" int b2u_masked_bce_fwd(const float* out, const float* gt, const float* mask, long long n, float* grad_unscaled,
                                  double* partials, float* loss, float* scale, void* stream) {
  B2U_REQUIRE(out && gt && mask && grad_unscaled && partials && loss && scale && n > 0, "bad arguments");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int blocks = b2u_masked_bce_blocks();
  masked_bce_partial_kernel<<<blocks, kLossThreads, 0, st>>>(out, gt, mask, n, grad_unscaled, partials);
  B2U_LAUNCH_CHECK();
  masked_bce_final_kernel<<<1, 1, 0, st>>>(partials, blocks, n, loss, scale);
  B2U_LAUNCH_CHECK();
  return B2U_OK;
}

extern "C" int b2u_masked_bce_bwd(const float* grad_unscaled, const float* upstream, const float* scale, float* grad_out, long long n,
                                  void* stream) {
  B2U_REQUIRE(grad_unscaled && upstream && scale && grad_out && n > 0, "bad arguments");
  long long blocks = (n + kLossThreads - 1) / kLossThreads;
  if (blocks > 148 * 8) blocks = 148 * 8;
  masked_bce_bwd_kernel<<<static_cast<int>(blocks), kLossThreads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(grad_unscaled, upstream, scale, grad_out, n);
  B2U_LAUNCH_CHECK();
  return B2U_OK;
}
