// Host-side helpers shared by the tensor-core convolution translation units.
#pragma once
#include "b2u_common.cuh"

namespace b2u {

int conv_encode_map(CUtensorMap* map, int dtype, int rank, const void* base, const cuuint64_t* dims,
                    const cuuint64_t* strides_bytes, const cuuint32_t* box, int swizzle_bytes = 128);
int conv_validate_desc(const b2u_conv_desc* d);
int conv_stat_subgroup(int cout, int num_groups);
int conv_ilog2(int v);

// conv3x3 v2 (conv3x3_v2.cu)
int conv3x3_v2_stat_layout(const b2u_conv_desc* d, int* rows_per_image, int* subgroup_size);
// fused A-operand prologue (b2u_conv3x3_pro_fwd): x is a RAW conv / pool output, activated on the fly in shared memory
struct V2Prologue {
  const void* coef;     // float2[n][cin]
  const void* mask;     // NHWC keep bits or nullptr
  int relu;
  int x_shared;
};
int conv3x3_v2_run(const void* x, const void* wpacked, void* y, float* partials, const b2u_conv_desc* d, void* stream,
                   const V2Prologue* pro = nullptr);

// convT2x2 v2 (convT_v2.cu): persistent, all taps of an N tile share one activation fetch
int convT_v2_stat_layout(const b2u_conv_desc* d, int* rows_per_image, int* subgroup_size);
int convT_v2_run(const void* x, const void* wpacked, void* y, float* partials, const b2u_conv_desc* d, void* stream);

}  // namespace b2u
