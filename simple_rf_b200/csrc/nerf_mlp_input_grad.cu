// Gradient of one fused NeRF MLP evaluation with respect to its INPUTS: the sample points and the view directions.
//
// Only learnable cameras need it (SimpleNeRF17.py:817-842: the pose correction r, t of ExtrinsicsLearner receives its gradient through
// the rays, i.e. through pts = o + z d and view_dirs, :210-214).  Every shipped configuration freezes the cameras, so this kernel is off the
// training hot path: it runs on the CUDA cores, one thread per sample row, and reads what the dgrad kernel left in HBM anyway -
// the pre-activation gradient images dZ_f [tile][slot][128 x 64 bf16, SW128] of every layer f.
//
//   g_enc[row, j]  = sum over the layers f that consume encoding column j:  sum_n dZ_f[row, n] * W_f[n, cols_f[j]]      (fp32 weights)
//   g_x[c]         = g_enc[c] + sum_k 2^k ( cos(2^k x_c) g_enc[3 + 6k + c] - sin(2^k x_c) g_enc[6 + 6k + c] )            (encoding backward)
//
// with the encoding column order of the forward kernel (nerf_mlp.cu, encoding warps).  Rows mode (num_samples == 0, the TensoRF colour MLP whose
// A operand is precomputed rows, SimpleTensoRF09.py:1411-1421): no encoding, g_points[row, c] = g_enc[c] of whatever three weight columns the
// source maps to image columns 0..2 (the view directions behind the products), rows >= *count written as zero.  The encodings were rounded to bf16 for the tensor
// cores; the derivative is taken of the unrounded encoding (straight-through), as the bf16 weight operands are in the other kernels.
#include <cuda_bf16.h>

#include "common.cuh"

namespace srf {

constexpr int IG_MAX_SOURCES = 6;
constexpr int IG_THREADS = 128;          // one thread per row of a 128-row tile
constexpr int IG_IMAGE_BYTES = 128 * 128;

struct InputGradSource {
  int32_t dz_slot;      // first dZ image of the consuming layer
  int32_t dz_images;    // its output width / 64
  int32_t in_total;     // row pitch of its fp32 weight matrix [64 * dz_images, in_total]
  int32_t target;       // 0: the points encoding image (64 columns), 1: the view encoding image (32 columns)
  int64_t w_offset;     // element offset of the weight matrix in `params`
  int32_t cols[64];     // encoding image column -> weight column (-1: the layer does not read that column)
};

struct InputGradParams {
  InputGradSource src[IG_MAX_SOURCES];
  int num_sources;
  const float* params;
  const uint8_t* dz;
  int dz_slots;
  const float* rays_o; const float* rays_d; const float* z; const float* view_dirs;
  long long total;
  const int* count;     // device, nullable: number of valid rows (the buffers are sized for a worst case)
  int S, points_degree, views_degree;
  float* g_points; float* g_views;
  long long num_tiles;
};

__device__ __forceinline__ float bf16_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf16_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }

// acc[0..N) += sum over the 64 units of one dZ image row of dz[n] * ws[n][0..N)
template <int N>
__device__ __forceinline__ void accumulate_image(const uint8_t* img, int row, const float (*ws)[64], float (&acc)[N]) {
  for (int u = 0; u < 8; ++u) {
    const uint4 q = *reinterpret_cast<const uint4*>(img + (size_t)row * 128 + ((u ^ (row & 7)) << 4));
    const float d[8] = {bf16_lo(q.x), bf16_hi(q.x), bf16_lo(q.y), bf16_hi(q.y), bf16_lo(q.z), bf16_hi(q.z), bf16_lo(q.w), bf16_hi(q.w)};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4* w4 = reinterpret_cast<const float4*>(ws[u * 8 + j]);
#pragma unroll
      for (int i = 0; i < N / 4; ++i) {
        const float4 w = w4[i];
        acc[4 * i + 0] = fmaf(d[j], w.x, acc[4 * i + 0]);
        acc[4 * i + 1] = fmaf(d[j], w.y, acc[4 * i + 1]);
        acc[4 * i + 2] = fmaf(d[j], w.z, acc[4 * i + 2]);
        acc[4 * i + 3] = fmaf(d[j], w.w, acc[4 * i + 3]);
      }
    }
  }
}

// g_x for one coordinate from the gradient of its encoding columns (x, then sin / cos of 2^k x, k < degree)
template <int N>
__device__ __forceinline__ float encoding_backward(const float (&g)[N], int c, float x, int degree) {
  float out = g[c];
#pragma unroll
  for (int k = 0; k < 10; ++k) {
    if (k < degree && 6 + 6 * k + c < N) {
      float s, co;
      sincosf(ldexpf(x, k), &s, &co);
      out = fmaf(ldexpf(1.f, k), co * g[3 + 6 * k + c] - s * g[6 + 6 * k + c], out);
    }
  }
  return out;
}

__global__ void __launch_bounds__(IG_THREADS) nerf_mlp_input_grad_kernel(const __grid_constant__ InputGradParams p) {
  __shared__ __align__(16) float ws[64][64];          // one 64-unit slice of a weight matrix, columns mapped to encoding columns
  const int row = threadIdx.x;
  const long long valid_rows = p.count != nullptr ? min((long long)__ldg(p.count), p.total) : p.total;
  for (long long tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
    if (tile * 128 >= valid_rows) {                    // nothing but padding: the dZ images of such tiles were never written
      const long long m = tile * 128 + row;
      if (m < p.total) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          p.g_points[m * 3 + c] = 0.f;
          if (p.g_views != nullptr) p.g_views[m * 3 + c] = 0.f;
        }
      }
      continue;
    }
    float ge[64], gv[32];
#pragma unroll
    for (int i = 0; i < 64; ++i) ge[i] = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) gv[i] = 0.f;
    for (int s = 0; s < p.num_sources; ++s) {
      const InputGradSource& src = p.src[s];
      const float* w = p.params + src.w_offset;
      for (int kb = 0; kb < src.dz_images; ++kb) {
        __syncthreads();                               // the previous slice has been consumed
        for (int e = threadIdx.x; e < 64 * 64; e += IG_THREADS) {
          const int n = e >> 6, ic = e & 63;
          const int col = src.cols[ic];
          ws[n][ic] = col >= 0 ? __ldg(w + (size_t)(kb * 64 + n) * src.in_total + col) : 0.f;
        }
        __syncthreads();
        const uint8_t* img = p.dz + ((size_t)tile * p.dz_slots + src.dz_slot + kb) * IG_IMAGE_BYTES;
        if (src.target == 0) accumulate_image<64>(img, row, ws, ge);
        else accumulate_image<32>(img, row, ws, gv);
      }
    }
    const long long m = tile * 128 + row;
    if (m < p.total && (p.S == 0 || m >= valid_rows)) {        // rows mode: the mapped columns as they are; padding rows: zero
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        p.g_points[m * 3 + c] = m < valid_rows ? ge[c] : 0.f;
        if (p.g_views != nullptr) p.g_views[m * 3 + c] = m < valid_rows ? gv[c] : 0.f;
      }
    } else if (m < p.total) {
      const long long r = m / p.S;
      const float zz = p.z[m];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float x = __fadd_rn(p.rays_o[r * 3 + c], __fmul_rn(p.rays_d[r * 3 + c], zz));     // the point the forward encoded
        p.g_points[m * 3 + c] = encoding_backward(ge, c, x, p.points_degree);
      }
      if (p.g_views != nullptr) {
#pragma unroll
        for (int c = 0; c < 3; ++c) p.g_views[m * 3 + c] = encoding_backward(gv, c, p.view_dirs[r * 3 + c], p.views_degree);
      }
    }
  }
}

}  // namespace srf

using namespace srf;

SRF_API int srf_input_grad_source_bytes(void) { return (int)sizeof(InputGradSource); }

SRF_API int srf_nerf_mlp_input_grad(const void* sources, int num_sources, const float* params, const void* dz, int dz_slots,
                                    const float* rays_o, const float* rays_d, const float* z, const float* view_dirs,
                                    int64_t num_rows, const int* count, int num_samples, int points_degree, int views_degree,
                                    float* g_points, float* g_views, void* stream) {
  const char* where = "srf_nerf_mlp_input_grad";
  if (num_rows == 0) return 0;
  SRF_REQUIRE(sources && params && dz && g_points, where, "null pointer");
  SRF_REQUIRE(num_samples == 0 || (rays_o && rays_d && z), where, "rays_o, rays_d, z required unless num_samples == 0 (rows mode)");
  SRF_REQUIRE(num_sources >= 1 && num_sources <= IG_MAX_SOURCES, where, "1..6 encoding consumers expected");
  SRF_REQUIRE(num_samples >= 0 && (num_samples == 0 || num_rows % num_samples == 0), where, "rows must be rays x samples");
  SRF_REQUIRE(points_degree >= 0 && points_degree <= 10 && views_degree <= 4, where, "encoding degree out of range (points <= 10, views <= 4)");
  SRF_REQUIRE((g_views == nullptr) || (view_dirs != nullptr && views_degree >= 0), where, "g_views needs view_dirs and a view encoding");
  InputGradParams p{};
  const InputGradSource* src = reinterpret_cast<const InputGradSource*>(sources);
  bool any_view = false;
  for (int i = 0; i < num_sources; ++i) {
    const InputGradSource& s = src[i];
    SRF_REQUIRE(s.dz_images >= 1 && s.dz_images <= 4 && s.dz_slot >= 0 && s.dz_slot + s.dz_images <= dz_slots, where, "image slot out of range");
    SRF_REQUIRE(s.target == 0 || s.target == 1, where, "target must be 0 (points) or 1 (views)");
    SRF_REQUIRE(s.in_total >= 1 && s.w_offset >= 0, where, "bad weight matrix");
    for (int c = 0; c < 64; ++c) {
      SRF_REQUIRE(s.cols[c] >= -1 && s.cols[c] < s.in_total, where, "weight column out of range");
      SRF_REQUIRE(s.target == 0 || c < 32 || s.cols[c] < 0, where, "the view encoding image holds 32 columns");
    }
    any_view = any_view || s.target == 1;
    p.src[i] = s;
  }
  SRF_REQUIRE(!any_view || g_views != nullptr, where, "a view-encoding consumer needs g_views");
  p.num_sources = num_sources; p.params = params; p.dz = reinterpret_cast<const uint8_t*>(dz); p.dz_slots = dz_slots;
  p.rays_o = rays_o; p.rays_d = rays_d; p.z = z; p.view_dirs = view_dirs;
  p.total = num_rows; p.count = count; p.S = num_samples; p.points_degree = points_degree; p.views_degree = views_degree < 0 ? 0 : views_degree;
  p.g_points = g_points; p.g_views = g_views;
  p.num_tiles = (num_rows + 127) / 128;
  long long grid = (long long)sm_count() * 8;
  if (grid > p.num_tiles) grid = p.num_tiles;
  nerf_mlp_input_grad_kernel<<<(int)grid, IG_THREADS, 0, (cudaStream_t)stream>>>(p);
  return check_launch(where);
}
