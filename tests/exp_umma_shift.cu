// Hardware experiment (not part of the product): can a K-major SWIZZLE_128B UMMA operand start at a
// 128-byte row offset that is NOT 1024-byte aligned, with an 8-row-group stride (SBO) other than 1024 B?
// That is what "load the (bh+2) x (bw+2) input patch once and view it through 9 shifted descriptors"
// needs.  A = TMA-loaded patch [R rows][64 bf16] (128B swizzle), B = identity [64][64], D = A_view.
// For each (row shift j, SBO, base_offset mode) the kernel dumps D; the host checks D[m][n] against
// patch[j + (m/8)*(SBO/128) + m%8][n].
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o exp_umma_shift tests/exp_umma_shift.cu -lcuda
#include "../unet_research_b200/csrc/b2u_common.cuh"
#include <vector>
#include <cstdlib>

void b2u_set_error(const char*, ...) {}
int b2u_num_sms() { return 148; }

using namespace b2u;

constexpr int R = 192;            // patch rows
constexpr int NV = 40;            // variants

struct Variant { int shift, sbo, bomode; };

__global__ void __launch_bounds__(128) exp_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                                   const Variant* __restrict__ vars, int nvars, float* __restrict__ out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smA = smem;                       // R*128 = 24 KB
  uint8_t* smB = smem + 32 * 1024;           // 64*128 = 8 KB
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 48 * 1024);
  uint64_t* mma_bar = bar + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    mbar_init(mma_bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) { tmem_alloc(tmem_slot, 64); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bar, R * 128 + 64 * 128);
    tma_load_2d(smA, &tmA, bar, 0, 0);
    tma_load_2d(smB, &tmB, bar, 0, 0);
  }
  mbar_wait(bar, 0);
  __syncthreads();
  constexpr uint32_t idesc = umma_idesc(128, 64, 1);
  for (int v = 0; v < nvars; ++v) {
    const Variant var = vars[v];
    if (threadIdx.x == 0) {
      tc_fence_after();
      const uint32_t a_addr = smem_u32(smA) + var.shift * 128;
      uint64_t ad = 0;
      ad |= static_cast<uint64_t>((a_addr & 0x3FFFF) >> 4);
      ad |= static_cast<uint64_t>(1) << 16;
      ad |= static_cast<uint64_t>(var.sbo >> 4) << 32;
      ad |= static_cast<uint64_t>(1) << 46;
      if (var.bomode == 1) ad |= static_cast<uint64_t>((a_addr >> 7) & 7) << 49;
      ad |= static_cast<uint64_t>(2) << 61;
      const uint64_t bd = umma_desc_k_sw128(smem_u32(smB));
      for (int k = 0; k < 4; ++k) umma_ss<false>(tmem_base, ad + 2 * k, bd + 2 * k, idesc, k != 0);
      umma_commit(mma_bar);
    }
    mbar_wait(mma_bar, v & 1);
    tc_fence_after();
    for (int chunk = 0; chunk < 2; ++chunk) {
      uint32_t r[32];
      tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + chunk * 32, r);
      tmem_ld_wait();
      const int row = warp * 32 + lane;
      for (int i = 0; i < 32; ++i) out[(static_cast<size_t>(v) * 128 + row) * 64 + chunk * 32 + i] = __uint_as_float(r[i]);
    }
    tc_fence_before();
    __syncthreads();
  }
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tmem_base, 64); }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  void* fnp = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fnp, cudaEnableDefault, &q);
  EncodeTiledFn enc = reinterpret_cast<EncodeTiledFn>(fnp);
  for (int pass = 0; pass < 2; ++pass) {
    std::vector<__nv_bfloat16> hA(R * 64), hB(64 * 64);
    for (int r = 0; r < R; ++r)
      for (int c = 0; c < 64; ++c) hA[r * 64 + c] = __float2bfloat16(pass == 0 ? float(r) : float(c));
    for (int n = 0; n < 64; ++n)
      for (int k = 0; k < 64; ++k) hB[n * 64 + k] = __float2bfloat16(n == k ? 1.f : 0.f);
    __nv_bfloat16 *dA, *dB;
    cudaMalloc(&dA, hA.size() * 2);
    cudaMalloc(&dB, hB.size() * 2);
    cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
    CUtensorMap ta, tb;
    cuuint32_t estr[2] = {1, 1};
    {
      cuuint64_t dims[2] = {64, R};
      cuuint64_t str[1] = {128};
      cuuint32_t box[2] = {64, R};
      CUresult rc = enc(&ta, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dA, dims, str, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (rc) { printf("encode A failed %d\n", (int)rc); return 1; }
    }
    {
      cuuint64_t dims[2] = {64, 64};
      cuuint64_t str[1] = {128};
      cuuint32_t box[2] = {64, 64};
      CUresult rc = enc(&tb, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dB, dims, str, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (rc) { printf("encode B failed %d\n", (int)rc); return 1; }
    }
    std::vector<Variant> vars;
    const int shifts[] = {0, 1, 2, 3, 7, 8, 10, 11, 12, 21};
    for (int s : shifts)
      for (int sbo : {1024, 1280})
        for (int bo : {0, 1}) vars.push_back({s, sbo, bo});
    Variant* dV;
    cudaMalloc(&dV, vars.size() * sizeof(Variant));
    cudaMemcpy(dV, vars.data(), vars.size() * sizeof(Variant), cudaMemcpyHostToDevice);
    float* dOut;
    cudaMalloc(&dOut, vars.size() * 128 * 64 * 4);
    cudaMemset(dOut, 0, vars.size() * 128 * 64 * 4);
    cudaFuncSetAttribute(exp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    exp_kernel<<<1, 128, 64 * 1024>>>(ta, tb, dV, (int)vars.size(), dOut);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(e)); return 1; }
    std::vector<float> hO(vars.size() * 128 * 64);
    cudaMemcpy(hO.data(), dOut, hO.size() * 4, cudaMemcpyDeviceToHost);
    for (size_t v = 0; v < vars.size(); ++v) {
      int bad = 0, first_bad_m = -1;
      for (int m = 0; m < 128; ++m) {
        const int src_row = vars[v].shift + (m / 8) * (vars[v].sbo / 128) + (m % 8);
        for (int n = 0; n < 64; ++n) {
          const float want = src_row < R ? (pass == 0 ? float(src_row) : float(n)) : -1.f;
          if (src_row < R && hO[(v * 128 + m) * 64 + n] != want) { if (!bad) first_bad_m = m; ++bad; }
        }
      }
      printf("pass %d shift %2d sbo %4d base_offset_mode %d : %s (bad %d, first bad row %d; D[1][0..2]= %.0f %.0f %.0f, D[9][8]= %.0f)\n",
             pass, vars[v].shift, vars[v].sbo, vars[v].bomode, bad ? "MISMATCH" : "ok", bad, first_bad_m,
             hO[(v * 128 + 1) * 64 + 0], hO[(v * 128 + 1) * 64 + 1], hO[(v * 128 + 1) * 64 + 2], hO[(v * 128 + 9) * 64 + 8]);
    }
    cudaFree(dA); cudaFree(dB); cudaFree(dV); cudaFree(dOut);
  }
  return 0;
}
