// Fused optimiser step of the reference training recipe (base_model_tests/training.py:31-51 + the Lightning flag
// gradient_clip_val = 0.5): global-norm gradient clipping + SGD with momentum over ALL parameters in two launches
// instead of ~10 multi-tensor passes:
//
//   sgd_sumsq_kernel : per 64 Ki-element chunk, sum of squared gradients (fp32 lanes, fp64 block total) -> partial[chunk]
//   sgd_step_kernel  : every block re-reduces the (few hundred) partials in the same fixed order (deterministic, no
//                      extra launch, no atomics), forms clip = min(1, max_norm / (norm + 1e-6)) as
//                      torch.nn.utils.clip_grad_norm_ does, then   g <- g * clip;  m <- mu * m + g  (m <- g on the first
//                      step, torch.optim.SGD semantics with dampening 0);  p <- p - lr * m.
//
// Tensors are described by a device table of (param, grad, momentum, numel); a chunk table maps blockIdx -> (tensor,
// offset).  HBM-bound: 4 B read per element in pass 1; 12 B read + 8..12 B written per element in pass 2.
#include "b2u_common.cuh"

namespace b2u {

struct SgdChunk {
  int tensor;
  int count;                 // elements of this chunk
  long long start;           // element offset inside the tensor
};

__global__ void __launch_bounds__(256) sgd_sumsq_kernel(const b2u_sgd_tensor* __restrict__ tensors, const SgdChunk* __restrict__ chunks,
                                                        double* __restrict__ partial) {
  pdl_wait();       // predecessor grid complete + visible (no-op without the PDL launch attribute)
  pdl_trigger();    // the successor may be scheduled once every CTA got here
  const SgdChunk ck = chunks[blockIdx.x];
  const float* g = tensors[ck.tensor].grad + ck.start;
  float acc = 0.f;
  const bool vec = ((reinterpret_cast<uintptr_t>(g) & 15) == 0);
  int i = threadIdx.x * 4;
  if (vec) {
    for (; i + 3 < ck.count; i += 256 * 4) {
      const float4 v = *reinterpret_cast<const float4*>(g + i);
      acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    }
    for (int j = (ck.count & ~3) + threadIdx.x; j < ck.count; j += 256) acc += g[j] * g[j];
  } else {
    for (int j = threadIdx.x; j < ck.count; j += 256) acc += g[j] * g[j];
  }
  __shared__ double sh[256];
  sh[threadIdx.x] = static_cast<double>(acc);
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}

__global__ void __launch_bounds__(256) sgd_step_kernel(const b2u_sgd_tensor* __restrict__ tensors, const SgdChunk* __restrict__ chunks,
                                                       const double* __restrict__ partial, int n_chunks, float lr, float momentum,
                                                       float max_norm, int first_step, float* __restrict__ norm_out) {
  pdl_wait();       // predecessor grid complete + visible (no-op without the PDL launch attribute)
  pdl_trigger();    // the successor may be scheduled once every CTA got here
  __shared__ double sh[256];
  __shared__ float clip_s;
  double a = 0.0;
  if (max_norm > 0.f) {
    for (int i = threadIdx.x; i < n_chunks; i += 256) a += partial[i];
    sh[threadIdx.x] = a;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
      if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
      __syncthreads();
    }
    if (threadIdx.x == 0) {
      const float norm = static_cast<float>(sqrt(sh[0]));
      const float c = max_norm / (norm + 1e-6f);
      clip_s = c < 1.f ? c : 1.f;
      if (blockIdx.x == 0 && norm_out) *norm_out = norm;
    }
    __syncthreads();
  } else if (threadIdx.x == 0) {
    clip_s = 1.f;
  }
  if (max_norm <= 0.f) __syncthreads();
  const float clip = clip_s;
  const SgdChunk ck = chunks[blockIdx.x];
  const b2u_sgd_tensor t = tensors[ck.tensor];
  float* p = t.param + ck.start;
  float* g = t.grad + ck.start;
  float* m = t.momentum + ck.start;
  auto upd = [&](float& pv, float& gv, float& mv) {
    gv = gv * clip;
    mv = (first_step || momentum == 0.f) ? gv : fmaf(momentum, mv, gv);
    pv = pv - lr * mv;
  };
  const bool vec = (((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m)) & 15) == 0);
  int tail_from = 0;
  if (vec) {
    const int n4 = ck.count >> 2;
    for (int j = threadIdx.x; j < n4; j += 256) {
      float4 pv = reinterpret_cast<float4*>(p)[j], gv = reinterpret_cast<float4*>(g)[j];
      float4 mv = (first_step || momentum == 0.f) ? make_float4(0.f, 0.f, 0.f, 0.f) : reinterpret_cast<float4*>(m)[j];
      upd(pv.x, gv.x, mv.x); upd(pv.y, gv.y, mv.y); upd(pv.z, gv.z, mv.z); upd(pv.w, gv.w, mv.w);
      reinterpret_cast<float4*>(p)[j] = pv;
      if (momentum != 0.f) reinterpret_cast<float4*>(m)[j] = mv;
      if (clip != 1.f) reinterpret_cast<float4*>(g)[j] = gv;
    }
    tail_from = n4 << 2;
  }
  for (int j = tail_from + threadIdx.x; j < ck.count; j += 256) {
    float pv = p[j], gv = g[j], mv = (first_step || momentum == 0.f) ? 0.f : m[j];
    upd(pv, gv, mv);
    p[j] = pv;
    if (momentum != 0.f) m[j] = mv;
    if (clip != 1.f) g[j] = gv;
  }
}

}  // namespace b2u

using namespace b2u;

extern "C" long long b2u_sgd_chunk_elems(void) { return 65536; }

extern "C" int b2u_sgd_step(const b2u_sgd_tensor* tensors_dev, const void* chunks_dev, int n_chunks, double* partial_dev, float lr,
                            float momentum, float max_grad_norm, int first_step, float* grad_norm_out, void* stream) {
  B2U_REQUIRE(tensors_dev && chunks_dev && n_chunks > 0, "bad arguments");
  B2U_REQUIRE(max_grad_norm <= 0.f || partial_dev, "clipping needs the partial-sum workspace (n_chunks doubles)");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const SgdChunk* chunks = reinterpret_cast<const SgdChunk*>(chunks_dev);
  if (max_grad_norm > 0.f) {
    B2U_PDL_LAUNCH((sgd_sumsq_kernel), n_chunks, 256, 0, st, tensors_dev, chunks, partial_dev);
    B2U_LAUNCH_CHECK();
  }
  B2U_PDL_LAUNCH((sgd_step_kernel), n_chunks, 256, 0, st, tensors_dev, chunks, partial_dev, n_chunks, lr, momentum, max_grad_norm, first_step, grad_norm_out);
  B2U_LAUNCH_CHECK();
  return B2U_OK;
}
