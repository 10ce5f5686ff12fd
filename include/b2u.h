/* b2u.h -- C ABI of libb2u.so, the B200 (sm_100a) kernels behind the U-Net / MC-DropBlock hot path.
 *
 * The reference (JohnDLee/Unet-Research) is pure Python: it has no FFI layer, its "operators" are
 * the torch.nn calls inside `UNet.forward` (unet_code/utils/utils_unet.py:408-449) and
 * `DropBlock2D.forward` (unet_code/utils/utils_modules.py:36-66).  Each entry point below replaces
 * the library op(s) named in its comment; the Python host (`unet_research_b200/`) binds them with
 * ctypes (see INTEGRATION.md for the stub a reference maintainer would add).
 *
 * Conventions
 *  - every pointer is a NON-OWNING DEVICE pointer into a caller-allocated buffer unless it is a
 *    `const b2u_*_desc*` (host POD struct) or says "host";
 *  - `stream` is a cudaStream_t passed as void*; all launches are asynchronous, allocate nothing,
 *    never synchronise, and are CUDA-graph capturable;
 *  - return 0 on success, B2U_ERR_* otherwise; `b2u_last_error()` returns a thread-local message;
 *    nothing throws or exits across the ABI;
 *  - activations are NHWC, `dtype` B2U_BF16 (bf16 storage, tcgen05 kind::f16) or B2U_F32 (fp32
 *    storage, tcgen05 kind::tf32); GroupNorm statistics, coefficients and accumulators are fp32/fp64.
 */
#ifndef B2U_H_
#define B2U_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2U_VERSION 1

#define B2U_OK 0
#define B2U_ERR_ARG 1
#define B2U_ERR_CUDA 2
#define B2U_ERR_UNSUPPORTED 3

#define B2U_BF16 0
#define B2U_F32 1
#define B2U_F16 2   /* fp16 storage and operands (kind::f16, F16 format), fp32 accumulate: inference only */

const char* b2u_last_error(void);
int b2u_version(void);
/* SM count and max threads per SM of the current device (needed to reproduce torch.rand's launch
 * geometry, ATen/native/cuda/DistributionTemplates.h:50-62). */
int b2u_device_info(int* sm_count, int* max_threads_per_sm);

/* ------------------------------------------------------------------ weight packing (once per model)
 * nn.Conv2d weight [Cout,Cin,3,3] fp32 -> [9][Cout][Cin] (K-major per tap) in `dtype`.
 * transpose_flip != 0 packs the data-gradient operand instead: [9][Cin][Cout] with taps rotated 180. */
int b2u_pack_conv3x3_weight(const float* w, void* packed, int cout, int cin, int dtype, int transpose_flip,
                            void* stream);
/* Both layouts in one pass over the fp32 weight (training repacks after every optimiser step); packed_dgrad may be
 * NULL.  cout and cin must be multiples of 32. */
int b2u_pack_conv3x3_weight_pair(const float* w, void* packed_fwd, void* packed_dgrad, int cout, int cin, int dtype,
                                 void* stream);
/* nn.ConvTranspose2d weight [Cin,Cout,2,2] fp32 -> [4][Cout][Cin] (tap = 2*i+j) in `dtype`. */
int b2u_pack_convT2x2_weight(const float* w, void* packed, int cin, int cout, int dtype, void* stream);
/* Every tensor-core weight of the model in ONE launch (training repacks all of them after every optimiser step:
 * 17 Conv2d + 4 ConvTranspose2d tensors of the canonical U-Net = 25 launches, 0.23 ms of a 4.7 ms step).
 * Same outputs, bit for bit, as the per-tensor entry points above.  The table lives in DEVICE memory (pointers are
 * stable: parameters are updated in place, packed buffers are rewritten in place); `first_block` is filled on the host
 * copy by b2u_pack_batched_plan before the upload.  cout and cin must be multiples of 32. */
typedef struct {
  const float* w;           /* fp32 parameter in its PyTorch layout                                              */
  void* out0;               /* kind 0: forward operand [9][Cout][Cin]        kind 1: forward operand [4][Cout][Cin] */
  void* out1;               /* kind 0: data-gradient operand [9][Cin][Cout]  kind 1: data-gradient operand [Cin][4*Cout]; may be NULL */
  int32_t kind;             /* 0 = nn.Conv2d 3x3 [Cout][Cin][3][3], 1 = nn.ConvTranspose2d 2x2 [Cin][Cout][2][2]  */
  int32_t cout, cin;
  int32_t first_block;      /* first block of this tensor in the flat grid (b2u_pack_batched_plan)               */
} b2u_pack_entry;
int b2u_pack_batched_plan(b2u_pack_entry* entries_host, int n, int* total_blocks);
int b2u_pack_batched(const b2u_pack_entry* entries_dev, int n, int total_blocks, int dtype, void* stream);

/* ------------------------------------------------------------------ tensor-core convolutions
 * Implicit-GEMM 3x3 / stride 1 / zero "same" padding / no bias  (replaces nn.Conv2d at
 * utils_unet.py:166-171,188-193,216-229,244-249,329-342,355-360) and the 2x2 stride-2 transposed
 * convolution (nn.ConvTranspose2d, utils_unet.py:311-315) as TMA-fed tcgen05 GEMMs.
 * Output is the RAW convolution result; per-(image, channel-subgroup) sum / sum-of-squares partials of
 * the fp32 accumulators are written for the GroupNorm that follows (utils_unet.py:177 etc.). */
typedef struct {
  int32_t n, h, w;          /* input batch / height / width (pixels)                               */
  int32_t cin, cout;        /* channels; cin % (128/sizeof(elt)) == 0, cout % 64 == 0              */
  int32_t dtype;            /* B2U_BF16 | B2U_F32 | B2U_F16                                        */
  int32_t num_groups;       /* GroupNorm groups of the FOLLOWING norm; 0 = no statistics           */
  int32_t x_cstride;        /* channel count of the tensor x lives in (>= cin; concat buffers)     */
  int32_t reserved[4];
} b2u_conv_desc;

/* rows of the partial-statistics buffer per image and sub-group size:
 * partials is float[n][rows][cout / sgs][2]. */
int b2u_conv3x3_stat_layout(const b2u_conv_desc* d, int* rows_per_image, int* subgroup_size);
int b2u_convT2x2_stat_layout(const b2u_conv_desc* d, int* rows_per_image, int* subgroup_size);
/* x: [n,h,w,x_cstride]; wpacked: [9][cout][cin]; y: [n,h,w,cout]; partials may be NULL iff num_groups==0 */
int b2u_conv3x3_fwd(const void* x, const void* wpacked, void* y, float* partials, const b2u_conv_desc* d,
                    void* stream);
/* The same convolution with the producer's normalisation fused into its A-operand path (north_star: "Norm, ReLU,
 * DropBlock mask-apply ... fused into the conv prologues"; reference order Conv -> GroupNorm -> DropBlock -> ReLU -> next
 * Conv, utils_unet.py:166-182): x_raw is the RAW output of the producing conv (or max-pool) [n,h,w,cin]; transform warps
 * rewrite every TMA-landed halo patch in shared memory as  act = [relu]((a*x + b) * keep)  -- coef float2[n][cin] from
 * b2u_gn_finalize (DropBlock rescale folded in), mask_bits the NHWC keep bits of that unit's DropBlock site or NULL --
 * before the tensor core reads it; pixels outside the image stay zero (the zero padding applies to the ACTIVATED
 * tensor).  Bit-identical to b2u_gn_apply followed by b2u_conv3x3_fwd; the activated tensor is never written to HBM.
 * 16-bit formats only.  x_shared != 0: all n images read the raw tensor of image 0 (shared first conv of the MC loop). */
int b2u_conv3x3_pro_fwd(const void* x_raw, const float* coef, const void* mask_bits, const void* wpacked, void* y,
                        float* partials, const b2u_conv_desc* d, int relu, int x_shared, void* stream);
/* x: [n,h,w,x_cstride]; wpacked: [4][cout][cin]; y: [n,2h,2w,cout] */
int b2u_convT2x2_fwd(const void* x, const void* wpacked, void* y, float* partials, const b2u_conv_desc* d,
                     void* stream);

/* First layer (Cin = 1 or 3, K = 9*Cin is not a tensor-core job): reads the fp32 NCHW network input
 * [n,cin,h0,w0] directly, applies UNet.autopad (utils_unet.py:451-458) by treating everything outside
 * h0 x w0 as zero, writes raw NHWC [n,h,w,cout] + GroupNorm partials float[n][rows][cout/sgs][2]. */
int b2u_conv_first_stat_layout(int h, int w, int cout, int num_groups, int* rows_per_image, int* subgroup_size);
int b2u_conv_first_fwd(const float* x_nchw, const float* wgt /*[cout,cin,3,3] fp32*/, void* y, float* partials,
                       int n, int cin, int h0, int w0, int h, int w, int cout, int num_groups, int dtype,
                       void* stream);

/* ------------------------------------------------------------------ GroupNorm (nn.GroupNorm(32, C), training.py:191)
 * Reduce the partials of one tensor to per-(image, channel) affine coefficients
 *   a = gamma * rstd * s,  b = (beta - mean * gamma * rstd) * s
 * where s = numel_per_call / keep_count is the DropBlock rescale of the site that follows the norm
 * (utils_modules.py:64; s > 0 so relu(s*z) = s*relu(z)); keep_counts == NULL means s = 1.
 * coef: float2[n][c].  count = elements per (image, group) = (c/num_groups) * H * W.
 * mean_rstd (optional, training): float2[n][num_groups] (mean, rstd) kept for the backward pass. */
int b2u_gn_finalize(const float* partials, int rows_per_image, int subgroup_size, const float* gamma,
                    const float* beta, float* coef, int n, int c, int num_groups, double count, float eps,
                    const unsigned long long* keep_counts, int images_per_call, double numel_per_call,
                    float* mean_rstd, void* stream);
/* Same, with shared_partials = 1: all n images share the partial rows of image 0 (Monte-Carlo DropBlock: the n
 * batched iterations see the same image, so the first conv and its statistics are computed once; only the
 * DropBlock rescale s differs per iteration, Dropblock_Uncertainty.py:64). */
int b2u_gn_finalize_ex(const float* partials, int rows_per_image, int subgroup_size, const float* gamma,
                       const float* beta, float* coef, int n, int c, int num_groups, double count, float eps,
                       const unsigned long long* keep_counts, int images_per_call, double numel_per_call,
                       float* mean_rstd, int shared_partials, void* stream);

/* Fused normalise-affine [+ DropBlock mask] [+ ReLU] [+ second mask and rescale]: the elementwise tail
 * of a conv unit (GroupNorm -> DropBlock -> ReLU, utils_unet.py:177-182) and, with mask2, the DropBlock
 * applied to the concatenated tensor (utils_unet.py:382-383) folded into the producer's store.
 *   v   = a*x + b;  v = mask1 ? v : 0;  v = relu ? max(v,0) : v;  out = mask2 ? v * s2 : 0
 * Masks are bit-packed NHWC keep-masks (1 = keep), uint32[n][h][w][channels/32]. */
typedef struct {
  int32_t n, h, w, c;             /* logical tensor                                                 */
  int32_t dtype;
  int32_t relu;
  int32_t out_cstride, out_coffset; /* output tensor channel count and first channel written        */
  int32_t mask2_cstride, mask2_coffset; /* channel count / offset of the mask2 tensor (concat site)  */
  int32_t images_per_call2;       /* images sharing one keep count of site 2                        */
  int32_t reserved[3];            /* reserved[0] == 1: every image reads x of image 0 (b2u_gn_apply only) */
  double numel_per_call2;         /* numel of one DropBlock call of site 2                          */
} b2u_apply_desc;
int b2u_gn_apply(const void* x, const float* coef, const uint32_t* mask1, const uint32_t* mask2,
                 const unsigned long long* keep_counts2, void* out, const b2u_apply_desc* d, void* stream);

/* Tail of an encoder block: the same fused apply, PLUS the 2x2/stride-2 max-pool of the activated tensor
 * (nn.MaxPool2d, utils_unet.py:265-266) written raw with GroupNorm partials for the norm that follows the
 * pool (utils_unet.py:282), PLUS the skip tensor written (with the concat-site mask2) into channels
 * [out_coffset, out_coffset+c) of the decoder's concat buffer (replaces x.clone() :420 and torch.cat :382).
 * pooled: [n,h/2,w/2,c]; pool_partials: float[n][rows][c/sgs][2]; argmax (optional, training): uint8
 * [n,h/2,w/2,c] holding the window index 0..3 (row-major, first maximum wins as in ATen). */
int b2u_pool_stat_layout(int h, int w, int c, int num_groups, int* rows_per_image, int* subgroup_size);
int b2u_gn_apply_pool(const void* x, const float* coef, const uint32_t* mask1, const uint32_t* mask2,
                      const unsigned long long* keep_counts2, void* skip_out, void* pooled, float* pool_partials,
                      uint8_t* argmax, int pool_num_groups, const b2u_apply_desc* d, void* stream);

/* ------------------------------------------------------------------ output head
 * GroupNorm-affine + DropBlock mask + ReLU of the last conv unit, the 1x1 conv 64->1 (utils_unet.py:397-402),
 * sigmoid (:404), crop to h0 x w0 (:440), clamp(0,1) (:443), NaN->0 (:444); then either
 *   out != NULL : out[n][h0][w0] fp32                         (UNet.forward result), and / or
 *   acc != NULL : Monte-Carlo accumulation of v = result * fov (Dropblock_Uncertainty.py:64-67):
 *                 acc[0][pix] += sum_n v, acc[1][pix] += sum_n v*v (fp64), and the sample with global
 *                 index (*iter_base + n) < return_num is stored to samples[idx][h0][w0].
 * logits (optional): pre-sigmoid values [n][h0][w0] fp32 for parity checks. */
typedef struct {
  int32_t n, h, w, c;             /* padded tensor [n,h,w,c], c % 8 == 0, c <= 512                  */
  int32_t h0, w0;                 /* un-padded output size                                           */
  int32_t dtype;
  int32_t return_num;
  int32_t fov_per_image;          /* 0: one fov [h0][w0] shared by all images, 1: fov[n][h0][w0]    */
  int32_t reserved[3];
} b2u_head_desc;
int b2u_head_fwd(const void* x, const float* coef, const uint32_t* mask1, const float* w_head, float* out,
                 float* logits, const float* fov, double* acc, float* samples, const long long* iter_base,
                 const b2u_head_desc* d, void* stream);
/* Launch shape b2u_head_fwd uses for the c = 64, 16-bit head of an h0 x w0 output on a device with num_sms SMs (host
 * only, no device work): plan[0] = 1 cp.async ring kernel / 0 register kernel, [1] threads per block, [2] blocks,
 * [3] trips per block (a trip = 8 pixels per 8-lane group), [4] ring stages, [5] dynamic shared memory per block. */
int b2u_head_plan(int h0, int w0, int num_sms, int* plan);
/* mean = S1/T, std = sqrt(max((S2 - S1*S1/T)/(T-1), 0)) (unbiased, torch.std default) */
int b2u_mc_finalize(const double* acc, float* mean, float* std, long long npix, long long t, void* stream);
/* acc[0] += v, acc[1] += v*v for v = x[i] * fov (rotation ensemble accumulation of already-final samples) */
int b2u_mc_accumulate(const float* x, const float* fov, double* acc, float* samples, const long long* iter_base,
                      int n, long long npix, int return_num, void* stream);
int b2u_advance_counter(long long* counter, long long delta, void* stream);

/* ------------------------------------------------------------------ DropBlock masks (DropBlock2D.forward,
 * utils_modules.py:46-58).  Two phases, both for a TABLE of calls so one launch covers all 22 sites of
 * several Monte-Carlo iterations:
 *  centers: bit p of call k = (torch.rand(...)[p] < gamma_k), reproducing torch's CUDA Philox stream for
 *           (seed, offset_k) bit-exactly (Philox4x32-10; element order of
 *           distribution_elementwise_grid_stride_kernel; raw-word thresholds instead of float compares);
 *  dilate : zero-pad(bs/2) + bs x bs stride-1 max-pool + (1 - .) as a bitwise OR-smear, transposed to the
 *           NHWC keep-mask layout, and the keep count (block_mask.sum()) per call.                     */
typedef struct {
  uint64_t philox_offset;     /* generator offset of this torch.rand call (multiple of 4)            */
  uint64_t center_word_off;   /* first uint32 word of this call's centre bitmap                       */
  uint64_t mask_word_off;     /* first uint32 word of this call's NHWC keep-mask                      */
  uint32_t numel;             /* n_img * c * (h-bs+1) * (w-bs+1)                                      */
  uint32_t grid;              /* torch's grid.x for this numel (host computes min(SMs*8, ceil(numel/256))) */
  uint32_t thresh_lo;         /* centre = word < thresh_lo || word >= thresh_hi                       */
  uint32_t thresh_hi;
  int32_t n_img, c, h, w;     /* tensor this call masks                                               */
  int32_t block_size;         /* odd, <= 31                                                           */
  int32_t count_index;        /* slot of keep_counts this call adds to                                */
  int32_t dilate_first_block; /* first block of this call in the flat dilate grid: filled by b2u_dropblock_plan */
  int32_t reserved;           /* first block of this call in the v2 (NHWC) dilate grid: filled by b2u_dropblock_plan */
} b2u_dropblock_call;
/* Fills dilate_first_block of every descriptor of a HOST table (prefix sum of the blocks each call needs in
 * b2u_dropblock_dilate's flat 1-D grid) and returns the total in *total_blocks.  Call it once, after the shapes are
 * set and before the table is copied to the device; b2u_dropblock_dilate rejects a table that was not planned.
 * (A one-entry table is planned by construction: the field is 0.) */
int b2u_dropblock_plan(b2u_dropblock_call* host_table, int n_calls, long long* total_blocks);
/* table: device array of n_calls descriptors; seed: host value; offset_base: device scalar added to every
 * philox_offset (lets a captured CUDA graph advance the stream between replays); may be NULL. */
int b2u_dropblock_centers(const b2u_dropblock_call* table, int n_calls, uint64_t seed,
                          const unsigned long long* offset_base, uint32_t* center_bits, void* stream);
/* Same stream of draws; the host copy of the table lets the launcher split every thread's trips over short blocks
 * (grid.z): the Monte-Carlo loop overlaps this kernel with the forward on a low-priority stream, and a forward kernel can
 * only start on an SM once resident mask blocks have drained. */
int b2u_dropblock_centers_ex(const b2u_dropblock_call* table, int n_calls, const b2u_dropblock_call* host_table,
                             uint64_t seed, const unsigned long long* offset_base, uint32_t* center_bits, void* stream);
/* Dropblock2d_ichan (utils_modules.py:86-139): centres = torch.bernoulli(p = gamma) over the full [N,C,H,W] tensor
 * (ATen bernoulli_tensor_cuda_kernel: thread idx draws one curand_uniform4 for elements 4*idx..4*idx+3, `u <= p`),
 * border of bs/2 zeroed; written into the same compact centre bitmap, so b2u_dropblock_dilate follows unchanged.
 * call.numel = n_img*c*h*w, call.thresh_lo = largest raw Philox word whose uniform is <= float(gamma).
 * center_words_total: size of center_bits in uint32 words (zeroed by this call). */
int b2u_dropblock_centers_ichan(const b2u_dropblock_call* table, int n_calls, const b2u_dropblock_call* host_table,
                                uint64_t seed, const unsigned long long* offset_base, uint32_t* center_bits,
                                long long center_words_total, void* stream);
int b2u_dropblock_dilate(const b2u_dropblock_call* table, int n_calls, const b2u_dropblock_call* host_table,
                         const uint32_t* center_bits, uint32_t* mask_bits, unsigned long long* keep_counts,
                         void* stream);
/* The same dilation for block_size 7 without the 32x32 bit transpose: the (sparse, gamma ~ 0.3 %) centres are scattered
 * with atomicOr into a zeroed NHWC word bitmap `scatter_bits` (same size and offsets as mask_bits: mask_words_total uint32
 * words, zeroed by this call), then keep = ~(7x7 box OR) runs word-parallel over 32 channels.  Bit-identical output. */
int b2u_dropblock_dilate_v2(const b2u_dropblock_call* table, int n_calls, const b2u_dropblock_call* host_table,
                            const uint32_t* center_bits, uint32_t* scatter_bits, long long mask_words_total,
                            uint32_t* mask_bits, unsigned long long* keep_counts, void* stream);
/* centre bitmap from caller-supplied uniforms (parity mode: feed the oracle's captured torch.rand values) */
int b2u_dropblock_centers_from_uniform(const float* u, uint32_t* center_bits, long long numel, float gamma,
                                       void* stream);

/* ------------------------------------------------------------------ backward pass (training step,
 * utils_training.py:21-39: loss.backward() through UNet.forward).
 *
 * Elementwise backward of one unit (GroupNorm -> DropBlock -> ReLU, utils_unet.py:177-182), two passes over the
 * same inputs plus a finalize; the upstream gradient is the sum of up to three sources that are never materialised:
 *   grad_a   : dense NHWC tensor (channel window [a_coffset, a_coffset+c) of a tensor with a_cstride channels),
 *              optionally times the concat-site keep mask / rescale (backward of dropblock(cat([up, skip])), :382-383);
 *   grad_pool: pooled-resolution gradient routed through the stored max-pool argmax (nn.MaxPool2d, :265-266);
 *   grad_out : head mode, dlogit * w_head with dlogit = grad_out * out * (1 - out) inside h0 x w0
 *              (Conv2d 1x1 + Sigmoid + crop, :397-404,440).
 * pass 1 writes partials float[n][rows][c][3] = (sum dZ, sum dZ*xhat, sum dlogit*act); finalize reduces them to
 * dgamma / dbeta / dw_head and to the per-(image, group) GroupNorm-backward coefficients float2[n][G];
 * pass 2 writes dY (gradient w.r.t. the raw conv output), NHWC or space-to-depth [n,h/2,w/2,4,c]. */
typedef struct {
  int32_t n, h, w, c, dtype, relu, num_groups, s2d;
  int32_t images_per_call1, a_cstride, a_coffset, mask2_cstride, mask2_coffset, images_per_call2, h0, w0;
  double numel_per_call1, numel_per_call2;
  const void* y;                         /* raw conv output of the unit [n,h,w,c]                              */
  const float* coef;                     /* float2[n][c] from b2u_gn_finalize                                  */
  const float* mean_rstd;                /* float2[n][G]                                                       */
  const float* gamma;                    /* [c]                                                                */
  const uint32_t* mask1;                 /* own-site keep mask or NULL                                         */
  const unsigned long long* keep_counts1;
  const void* grad_a;
  const uint32_t* mask2;
  const unsigned long long* keep_counts2;
  const void* grad_pool;                 /* [n,h/2,w/2,c]                                                      */
  const uint8_t* argmax;                 /* [n,h/2,w/2,c]                                                      */
  const float* grad_out;                 /* [n,1,h0,w0]                                                        */
  const float* out;                      /* [n,1,h0,w0]                                                        */
  const float* w_head;                   /* [c]                                                                */
} b2u_unit_bwd_desc;
int b2u_unit_bwd_rows(int h, int w, int c, int* rows_per_image);
int b2u_unit_bwd_stats(const b2u_unit_bwd_desc* d, float* partials, void* stream);
int b2u_unit_bwd_finalize(const float* partials, int n, int rows_per_image, int c, int num_groups, const float* gamma,
                          double count, float* group_coef, float* dgamma, float* dbeta, float* dw_head, void* stream);
int b2u_unit_bwd_apply(const b2u_unit_bwd_desc* d, const float* group_coef, void* dy, void* stream);

/* Weight gradients as tcgen05 GEMMs with K = pixels (both operands MN-major, straight from the NHWC tensors):
 *   conv3x3 (taps = 9): dW[co][ci][r][s] = sum_p g[p][co] * x[p + (r-1, s-1)][ci]   (nn.Conv2d backward)
 *   1x1     (taps = 1): dW[cg][cx]       = sum_p g[p][cg] * x[p][cx]                (transposed conv, with g in the
 *                                                                                    space-to-depth layout)
 * g: [n,h,w,cg] gradient w.r.t. the raw conv output, x: [n,h,w,x_cstride] the conv's input.  Split-K partial sums go to
 * `workspace` (float[slices][taps][cg][cx], size from b2u_wgrad_workspace_floats) and are reduced in a fixed order into
 * dw with the PyTorch layout selected by `layout`: 0 = Conv2d [cg][cx][3][3], 1 = ConvTranspose2d [cx][cg/4][2][2]
 * (cg = 4 * Cout, tap-major). */
typedef struct {
  int32_t n, h, w, cg, cx, x_cstride, taps, layout, dtype;
  int32_t reserved[3];
} b2u_wgrad_desc;
int b2u_wgrad_workspace_floats(const b2u_wgrad_desc* d, long long* floats);
int b2u_wgrad(const void* g, const void* x, float* workspace, float* dw, const b2u_wgrad_desc* d, void* stream);
/* first layer (Cin 1 or 3): dW[co][ci][3][3] = sum_p g[p][co] * x_nchw[p + tap][ci], direct kernel on the fp32 input */
int b2u_wgrad_first(const void* g, const float* x_nchw, float* workspace, float* dw, int n, int cin, int h0, int w0, int h,
                    int w, int cout, int dtype, void* stream);
/* partial rows per image of b2u_wgrad_first: workspace = float[n][rows][cout][cin*9] */
int b2u_wgrad_first_rows(void);
/* plain 1x1 GEMM y[p][co] = sum_k x[p][k] * w[co][k] on the tcgen05 path (data gradient of the transposed conv:
 * x = dY in space-to-depth layout with k = 4*Cout, w packed [1][cin][4*Cout]) */
int b2u_gemm1x1_fwd(const void* x, const void* wpacked, void* y, const b2u_conv_desc* d, void* stream);
/* nn.ConvTranspose2d weight [Cin,Cout,2,2] fp32 -> [1][Cin][4*Cout] (k = tap*Cout + co) for b2u_gemm1x1_fwd */
int b2u_pack_convT2x2_dgrad_weight(const float* w, void* packed, int cin, int cout, int dtype, void* stream);

/* ------------------------------------------------------------------ fused masked BCE loss of the training step
 * (utils_training.py:28-33 with nn.BCELoss(), base_model_tests/training.py:195):
 *   seg = out*mask; gt = gt*mask; loss = mean(BCE(seg, gt)) * numel / count_nonzero(mask)
 * b2u_masked_bce_fwd: one pass over the three fp32 maps [n elements] -> loss (scalar), scale (scalar: d loss / d sum of
 *   BCE terms) and grad_unscaled[n] = d BCE_i / d out_i (ATen's clamps: log >= -100, denominator >= 1e-12);
 *   partials: double[2 * b2u_masked_bce_blocks()] scratch; deterministic (fixed-order reduction).
 * b2u_masked_bce_bwd: grad_out[i] = grad_unscaled[i] * (*upstream) * (*scale)  (upstream = d L / d loss, a device scalar). */
int b2u_masked_bce_blocks(void);
int b2u_masked_bce_fwd(const float* out, const float* gt, const float* mask, long long n, float* grad_unscaled,
                       double* partials, float* loss, float* scale, void* stream);
int b2u_masked_bce_bwd(const float* grad_unscaled, const float* upstream, const float* scale, float* grad_out, long long n,
                       void* stream);

/* ------------------------------------------------------------------ rotation (torchvision TF.rotate, BILINEAR,
 * fill 0, as called at Rotational_Uncertainty.py:54,58): affine grid in fp32, grid_sample(bilinear, zeros,
 * align_corners=False) of the image and of a ones channel, result img*m.  x,out: [n][c][h][w] fp32;
 * angles_deg: host array of n angles (counter-clockwise degrees, as passed to TF.rotate). */
int b2u_rotate_bilinear(const float* x, float* out, int n, int c, int h, int w, const double* angles_deg,
                        int x_batch_stride_is_zero, void* stream);
/* CUDA-graph form of the rotation ensemble loop (Rotational_Uncertainty.py:51-63): the per-angle affine coefficients
 * live in a DEVICE table (6 floats per angle, filled on the host by b2u_rotation_table with exactly the arithmetic of
 * b2u_rotate_bilinear) indexed by the global angle index `*iter_base_dev + k`, so one captured graph serves every step.
 *   b2u_rotation_table:         host helper, table_host[n][6] for angles_deg[n] (no launch);
 *   b2u_rotate_in_table:        x [c][h][w] (ONE image) -> out [n][c][h][w], image k rotated by row min(base + k, len-1);
 *   b2u_rotate_back_accumulate: seg [n][h][w] -> rotate back by row base + k, * fov (may be NULL), fp64 per-pixel
 *                               (sum, sum of squares) into acc[2][h*w], the first return_num samples into
 *                               samples[return_num][h*w] (may be NULL); images with base + k >= *iter_limit_dev are skipped. */
int b2u_rotation_table(const double* angles_deg, int n, int h, int w, float* table_host);
int b2u_rotate_in_table(const float* x, float* out, int n, int c, int h, int w, const float* table_dev, int table_len,
                        const long long* iter_base_dev, void* stream);
int b2u_rotate_back_accumulate(const float* seg, const float* fov, double* acc, float* samples, int n, int h, int w,
                               int return_num, const float* table_dev, int table_len, const long long* iter_base_dev,
                               const long long* iter_limit_dev, void* stream);

/* ------------------------------------------------------------------ fused square_pad + TF.resize
 * (utils_general.py:32-43 + torchvision TF.resize on tensors = bilinear, align_corners=False, antialias=True; used by
 * the multi-fidelity steps MF-training-UNI.py:54-73 and DropBlockEval(resize=...) Dropblock_Uncertainty.py:52-61).
 * x: [planes][h][w] fp32 -> out: [planes][oh][ow]; square_pad != 0 first zero-pads (virtually) to max(h,w) with the
 * reference's split (top = d/2, left = d - d/2).  ATen's anti-aliased triangle filter, fp32. */
int b2u_square_pad_resize(const float* x, float* out, int planes, int h, int w, int square_pad, int oh, int ow,
                          void* stream);
/* Its backward (the multi-fidelity training steps resize the segmentation back up before the loss,
 * MF-training-UNI.py:66-69): grad_in[planes][h][w] = adjoint of the same filter applied to grad_out[planes][oh][ow]. */
int b2u_square_pad_resize_bwd(const float* grad_out, float* grad_in, int planes, int h, int w, int square_pad, int oh, int ow,
                              void* stream);

/* ------------------------------------------------------------------ evaluation metrics (utils_metrics.py:157-173)
 * TP, FP, FN, TN of round(seg) against (long)gt over the pixels with (long)mask != 0 (the FOV), in one pass on the
 * device: F1 (Dice) = 2TP / (2TP + FP + FN), accuracy = (TP + TN) / N.  counts4: 4 x uint64 (zeroed by the call). */
int b2u_confusion_counts(const float* seg, const float* gt, const float* mask, long long n,
                         unsigned long long* counts4, void* stream);

/* ------------------------------------------------------------------ fused optimiser step
 * torch.optim.SGD(lr, momentum) (base_model_tests/training.py:32) + Lightning's gradient_clip_val (global L2 norm,
 * torch.nn.utils.clip_grad_norm_ semantics: clip = min(1, max_norm / (norm + 1e-6)); grads are scaled in place).
 * tensors_dev: device array of b2u_sgd_tensor; chunks_dev: device array of {int32 tensor, int32 count, int64 start}
 * covering every tensor in pieces of at most b2u_sgd_chunk_elems() elements; partial_dev: n_chunks doubles.
 * max_grad_norm <= 0 disables clipping; first_step != 0 initialises the momentum buffers with the gradient. */
typedef struct {
  float* param;
  float* grad;
  float* momentum;
  long long numel;
} b2u_sgd_tensor;
long long b2u_sgd_chunk_elems(void);
int b2u_sgd_step(const b2u_sgd_tensor* tensors_dev, const void* chunks_dev, int n_chunks, double* partial_dev, float lr,
                 float momentum, float max_grad_norm, int first_step, float* grad_norm_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B2U_H_ */
