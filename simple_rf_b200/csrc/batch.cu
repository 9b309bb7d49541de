// "Next" row f2: device-side assembly of one training batch from the cached per-pixel tables.
//
// Replaces src/data_preprocessors/DataPreprocessor10.py:530-549 (load_nerf_cached_batch) and :568-595
// (load_sparse_depth_cached_batch): five "-1"-initialised tensors, two boolean-mask index reads and seven masked
// index_put_ writes per iteration (~30 launches with two host synchronisations for the mask sizes) become one gather:
// row b of the batch reads flat pixel index indices[b]; image rays (mask 0) take pixel_id + target_rgb, sparse-depth rays
// (mask 1) take pixel_id + depth / reprojection error / 3-D point; every field a ray kind does not carry stays -1.
#include "common.cuh"

namespace srf {

struct BatchParams {
  const long long* indices;    // [B] flat pixel indices (view * h * w + y * w + x)
  const uint8_t* is_sd;        // [B] 1: sparse-depth ray, 0: image ray; nullptr: all image rays
  const int* pixel_table;      // [N,3] (view, x, y)
  const float* rgb_table;      // [N,3]
  const float* depth_table;    // [N] or nullptr
  const float* error_table;    // [N] or nullptr
  const float* points_table;   // [N,3] or nullptr
  int* pixel_id;               // [B,3]
  float* target_rgb;           // [B,3]
  float* sd_depth;             // [B] or nullptr
  float* sd_error;             // [B] or nullptr
  float* sd_points;            // [B,3] or nullptr
  int* error_flag;             // nullable; set to 1 when an index falls outside [0, N) (that row is filled with -1)
  long long B, N;
};

__global__ void __launch_bounds__(256) assemble_batch_kernel(const BatchParams p) {
  const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= p.B) return;
  const long long i = p.indices[b];
  const bool sd = p.is_sd != nullptr && p.is_sd[b] != 0;
  if (i < 0 || i >= p.N) {      // the reference's fancy indexing raises here; never read out of bounds
    if (p.error_flag != nullptr) *p.error_flag = 1;
#pragma unroll
    for (int c = 0; c < 3; ++c) { p.pixel_id[b * 3 + c] = -1; p.target_rgb[b * 3 + c] = -1.f; }
    if (p.sd_depth != nullptr) p.sd_depth[b] = -1.f;
    if (p.sd_error != nullptr) p.sd_error[b] = -1.f;
    if (p.sd_points != nullptr) { for (int c = 0; c < 3; ++c) p.sd_points[b * 3 + c] = -1.f; }
    return;
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) p.pixel_id[b * 3 + c] = __ldg(p.pixel_table + i * 3 + c);
#pragma unroll
  for (int c = 0; c < 3; ++c) p.target_rgb[b * 3 + c] = sd ? -1.f : __ldg(p.rgb_table + i * 3 + c);
  if (p.sd_depth != nullptr) p.sd_depth[b] = sd ? __ldg(p.depth_table + i) : -1.f;
  if (p.sd_error != nullptr) p.sd_error[b] = sd ? __ldg(p.error_table + i) : -1.f;
  if (p.sd_points != nullptr) {
#pragma unroll
    for (int c = 0; c < 3; ++c) p.sd_points[b * 3 + c] = sd ? __ldg(p.points_table + i * 3 + c) : -1.f;
  }
}

}  // namespace srf

using namespace srf;

SRF_API int srf_assemble_batch(const int64_t* indices, const uint8_t* is_sparse_depth, int64_t batch, int64_t num_pixels,
                               const int* pixel_table, const float* rgb_table, const float* depth_table, const float* error_table,
                               const float* points_table, int* pixel_id, float* target_rgb, float* sd_depth, float* sd_error,
                               float* sd_points, int* error_flag, void* stream) {
  if (batch == 0) return 0;
  SRF_REQUIRE(indices && pixel_table && rgb_table && pixel_id && target_rgb, "srf_assemble_batch", "null pointer");
  SRF_REQUIRE((sd_depth == nullptr || depth_table) && (sd_error == nullptr || error_table) && (sd_points == nullptr || points_table),
              "srf_assemble_batch", "an output was requested without its table");
  SRF_REQUIRE(num_pixels > 0, "srf_assemble_batch", "empty tables");
  BatchParams p{};
  p.indices = reinterpret_cast<const long long*>(indices); p.is_sd = is_sparse_depth;
  p.pixel_table = pixel_table; p.rgb_table = rgb_table; p.depth_table = depth_table; p.error_table = error_table;
  p.points_table = points_table;
  p.pixel_id = pixel_id; p.target_rgb = target_rgb; p.sd_depth = sd_depth; p.sd_error = sd_error; p.sd_points = sd_points;
  p.error_flag = error_flag;
  p.B = batch; p.N = num_pixels;
  assemble_batch_kernel<<<(int)((batch + 255) / 256), 256, 0, (cudaStream_t)stream>>>(p);
  return check_launch("srf_assemble_batch");
}
