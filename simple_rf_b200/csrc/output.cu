// "Next" row f4: output tail of a rendered frame.
//
// Replaces src/data_preprocessors/DataPreprocessor10.py:775-803 (retrieve_inference_outputs) + :967-995
// (post_process_output / post_process_image / post_process_depth): the reference copies EVERY tensor of the model's output
// dict to the host through pageable memory (rays, per-ray camera matrices, coarse and fine maps: ~190 B/ray) and then clips /
// rounds / converts on the CPU.  Here one kernel builds exactly what the caller keeps — the uint8 image and the clipped
// depth maps — in one contiguous device record, which the host side ships with a single copy into pinned memory (19 B/ray).
//
// Arithmetic = numpy's: clip(rgb, 0, 1) * 255 in fp32, round half to even, cast to uint8; depth < 0 -> 0 (NaN preserved).
#include "common.cuh"

namespace srf {

struct FrameOutParams {
  const float* rgb;            // [N,3]
  const float* maps[4];        // [N] each, nullable
  uint8_t* image;              // [N,3]
  float* out_maps[4];          // [N] each
  long long N;
};

__device__ __forceinline__ uint8_t to_level(float c) {
  c = c < 0.f ? 0.f : (c > 1.f ? 1.f : c);            // numpy.clip (NaN falls through both comparisons)
  return (uint8_t)(int)rintf(__fmul_rn(c, 255.f));
}

__global__ void __launch_bounds__(256) frame_outputs_kernel(const FrameOutParams p) {
  // one thread per 4 rays: 12 colour values in (3 x float4), 12 bytes out (3 x 32-bit stores), 4 x float4 maps
  const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long r0 = q * 4;
  if (r0 >= p.N) return;
  if (r0 + 4 <= p.N) {
    const float4* src = reinterpret_cast<const float4*>(p.rgb + r0 * 3);
    uint32_t* dst = reinterpret_cast<uint32_t*>(p.image + r0 * 3);
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const float4 v = __ldg(src + j);
      dst[j] = (uint32_t)to_level(v.x) | ((uint32_t)to_level(v.y) << 8) | ((uint32_t)to_level(v.z) << 16) | ((uint32_t)to_level(v.w) << 24);
    }
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      if (p.maps[m] == nullptr) continue;
      float4 v = __ldg(reinterpret_cast<const float4*>(p.maps[m] + r0));
      v.x = v.x < 0.f ? 0.f : v.x; v.y = v.y < 0.f ? 0.f : v.y; v.z = v.z < 0.f ? 0.f : v.z; v.w = v.w < 0.f ? 0.f : v.w;
      *reinterpret_cast<float4*>(p.out_maps[m] + r0) = v;
    }
    return;
  }
  for (long long r = r0; r < p.N; ++r) {              // ragged tail (N % 4 rays)
    for (int c = 0; c < 3; ++c) p.image[r * 3 + c] = to_level(p.rgb[r * 3 + c]);
    for (int m = 0; m < 4; ++m)
      if (p.maps[m] != nullptr) { const float v = p.maps[m][r]; p.out_maps[m][r] = v < 0.f ? 0.f : v; }
  }
}

}  // namespace srf

using namespace srf;

SRF_API int64_t srf_frame_record_bytes(int64_t num_rays, int num_maps) {
  const int64_t image = (num_rays * 3 + 15) / 16 * 16;
  return image + (int64_t)num_maps * ((num_rays * 4 + 15) / 16 * 16);
}

SRF_API int srf_frame_outputs(const float* rgb, const float* depth, const float* depth_var, const float* depth_ndc,
                              const float* depth_var_ndc, int64_t num_rays, uint8_t* record, void* stream) {
  if (num_rays == 0) return 0;
  SRF_REQUIRE(rgb && depth && depth_var && record, "srf_frame_outputs", "null pointer");
  SRF_REQUIRE((depth_ndc == nullptr) == (depth_var_ndc == nullptr), "srf_frame_outputs", "depth_ndc and depth_var_ndc come together");
  SRF_REQUIRE(((uintptr_t)rgb | (uintptr_t)depth | (uintptr_t)depth_var | (uintptr_t)depth_ndc | (uintptr_t)depth_var_ndc |
               (uintptr_t)record) % 16 == 0, "srf_frame_outputs", "buffers must be 16-byte aligned");
  FrameOutParams p{};
  p.rgb = rgb; p.N = num_rays;
  p.maps[0] = depth; p.maps[1] = depth_var; p.maps[2] = depth_ndc; p.maps[3] = depth_var_ndc;
  p.image = record;
  const int64_t image = (num_rays * 3 + 15) / 16 * 16, map = (num_rays * 4 + 15) / 16 * 16;
  for (int m = 0; m < 4; ++m) p.out_maps[m] = reinterpret_cast<float*>(record + image + m * map);
  const long long quads = (num_rays + 3) / 4;
  frame_outputs_kernel<<<(int)((quads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(p);
  return check_launch("srf_frame_outputs");
}
