// Patch-reprojection masks of the depth-consistency losses that follow the render path in every training step
// (SURVEY.md §8f row f1).
//
// Replaces the mask computation of src/loss_functions/AugmentationsDepthLoss11.py:105-182 and
// src/loss_functions/CoarseFineConsistencyLoss34.py:89-164 (+ src/utils/CommonUtils04.py:227-253, `reproject`): per image
// ray, two candidate depths are turned into world points, reprojected into the nearest other training view, 5x5 rgb
// patches around the source pixel and the two reprojections are compared (RMSE), and two boolean masks say which model is
// the more accurate one.  The reference spends 3 x 25 fancy-index gathers in Python loops plus ~60 small elementwise
// kernels on this; here it is one launch, one warp per ray, lanes over the 75 patch elements.
//
// The images are read with the reference's zero-outside semantics (it pads the images and lets negative indices wrap into
// the padding).  The masks are not differentiable; the loss itself (two masked means) stays in torch.
#include "common.cuh"

namespace srf {

struct PatchParams {
  const float* rays_o; const float* rays_d; const float* depth1; const float* depth2;
  const int* pixel_id;          // [N,3] (view, x, y)
  const int* closest_view;      // [V]
  const float* poses;           // [V,4,4] camera-to-world
  const float* images;          // [V,H,W,3]
  uint8_t* mask1; uint8_t* mask2;
  float* rmse1; float* rmse2;   // nullable
  long long N;
  int V, H, W, hpx, hpy;
  const float* k;               // [3,3] intrinsics of the first ray (the reference hard-codes intrinsics[:1]), device pointer
  float thr;
  int both_invalid_rule;
};

// (.round().long()) of a reprojected coordinate: round half to even; non-finite or out-of-range values become INT64_MIN
// on the reference's x86 host, i.e. "invalid and clipped to 0"
__device__ __forceinline__ long long round_to_long(float v) {
  if (!(fabsf(v) < 9.0e18f)) return LLONG_MIN;
  return (long long)rintf(v);
}

__device__ __forceinline__ float texel(const PatchParams& p, int view, long long y, long long x, int c) {
  if (y < 0 || y >= p.H || x < 0 || x >= p.W) return 0.f;
  return __ldg(p.images + (((size_t)view * p.H + (size_t)y) * p.W + (size_t)x) * 3 + c);
}

constexpr int PATCH_WARPS = 8;

__global__ void __launch_bounds__(PATCH_WARPS * 32) patch_reprojection_kernel(PatchParams p) {
  const int warp = threadIdx.x >> 5, lane = lane_id();
  const long long r = (long long)blockIdx.x * PATCH_WARPS + warp;
  if (r >= p.N) return;
  const int va = p.pixel_id[r * 3 + 0];
  const long long xa = p.pixel_id[r * 3 + 1], ya = p.pixel_id[r * 3 + 2];
  const int vb = p.closest_view[va];
  const float* P = p.poses + (size_t)vb * 16;
  // M = K * diag(1,-1,-1) * R_b^T, evaluated left to right like the reference's matmul chain
  float kp[9], M[9];
#pragma unroll
  for (int i = 0; i < 3; ++i) { kp[i * 3 + 0] = __ldg(p.k + i * 3 + 0); kp[i * 3 + 1] = -__ldg(p.k + i * 3 + 1); kp[i * 3 + 2] = -__ldg(p.k + i * 3 + 2); }
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j)      // (R^T)[k][j] = R[j][k] = P[j*4 + k]
      M[i * 3 + j] = fmaf(kp[i * 3 + 2], P[j * 4 + 2], fmaf(kp[i * 3 + 1], P[j * 4 + 1], __fmul_rn(kp[i * 3 + 0], P[j * 4 + 0])));
  long long xb[2], yb[2];
  float dep[2] = {p.depth1[r], p.depth2[r]};
#pragma unroll
  for (int m = 0; m < 2; ++m) {
    float d[3];
#pragma unroll
    for (int a = 0; a < 3; ++a)
      d[a] = __fadd_rn(__fadd_rn(p.rays_o[r * 3 + a], __fmul_rn(p.rays_d[r * 3 + a], dep[m])), -P[a * 4 + 3]);
    float q[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) q[i] = fmaf(M[i * 3 + 2], d[2], fmaf(M[i * 3 + 1], d[1], __fmul_rn(M[i * 3 + 0], d[0])));
    xb[m] = round_to_long(__fdiv_rn(q[0], q[2]));
    yb[m] = round_to_long(__fdiv_rn(q[1], q[2]));
  }
  const int W = p.W, H = p.H, hpx = p.hpx, hpy = p.hpy;
  auto valid = [&](long long x, long long y) { return x >= hpx && x < W - hpx && y >= hpy && y < H - hpy; };
  const bool v_a = valid(xa, ya), v_1 = valid(xb[0], yb[0]), v_2 = valid(xb[1], yb[1]);
  long long xc[2], yc[2];
#pragma unroll
  for (int m = 0; m < 2; ++m) {
    xc[m] = xb[m] < 0 ? 0 : (xb[m] > W - 1 ? W - 1 : xb[m]);
    yc[m] = yb[m] < 0 ? 0 : (yb[m] > H - 1 ? H - 1 : yb[m]);
  }
  const int px = 2 * hpx + 1, py = 2 * hpy + 1, n = px * py * 3;
  float s1 = 0.f, s2 = 0.f;
  for (int e = lane; e < n; e += 32) {
    const int c = e % 3, j = (e / 3) % px, i = e / (3 * px);
    const int oy = i - hpy, ox = j - hpx;
    const float a = texel(p, va, ya + oy, xa + ox, c);
    const float b1 = texel(p, vb, yc[0] + oy, xc[0] + ox, c);
    const float b2 = texel(p, vb, yc[1] + oy, xc[1] + ox, c);
    const float e1 = a - b1, e2 = a - b2;
    s1 = fmaf(e1, e1, s1);
    s2 = fmaf(e2, e2, s2);
  }
  s1 = warp_sum(s1);
  s2 = warp_sum(s2);
  if (lane == 0) {
    const float r1 = sqrtf(__fdiv_rn(s1, (float)n)), r2 = sqrtf(__fdiv_rn(s2, (float)n));
    bool m1 = ((r1 < r2) || !v_2) && (r1 < p.thr) && v_1 && v_a;
    bool m2 = ((r2 < r1) || !v_1) && (r2 < p.thr) && v_2 && v_a;
    if (p.both_invalid_rule && !v_1 && !v_2) {
      m1 = m1 || (dep[0] > dep[1]);
      m2 = m2 || (dep[1] > dep[0]);
    }
    p.mask1[r] = m1 ? 1 : 0;
    p.mask2[r] = m2 ? 1 : 0;
    if (p.rmse1) p.rmse1[r] = r1;
    if (p.rmse2) p.rmse2[r] = r2;
  }
}

}  // namespace srf

using namespace srf;

SRF_API int srf_patch_reprojection_masks(const float* rays_o, const float* rays_d, const float* depth1, const float* depth2,
                                         const int32_t* pixel_id, int64_t num_rays, const int32_t* closest_view, const float* poses,
                                         const float* intrinsics_first, const float* images, int num_views, int height, int width,
                                         int patch_x, int patch_y, float rmse_threshold, int both_invalid_rule, uint8_t* mask1,
                                         uint8_t* mask2, float* rmse1, float* rmse2, void* stream) {
  if (num_rays == 0) return 0;
  SRF_REQUIRE(rays_o && rays_d && depth1 && depth2 && pixel_id && closest_view && poses && intrinsics_first && images && mask1 && mask2,
              "srf_patch_reprojection_masks", "null pointer");
  SRF_REQUIRE(num_views > 0 && height > 0 && width > 0 && patch_x >= 1 && patch_y >= 1 && (patch_x & 1) && (patch_y & 1),
              "srf_patch_reprojection_masks", "need positive sizes and odd patch sizes");
  PatchParams p{};
  p.rays_o = rays_o; p.rays_d = rays_d; p.depth1 = depth1; p.depth2 = depth2; p.pixel_id = pixel_id; p.closest_view = closest_view;
  p.poses = poses; p.images = images; p.mask1 = mask1; p.mask2 = mask2; p.rmse1 = rmse1; p.rmse2 = rmse2;
  p.N = num_rays; p.V = num_views; p.H = height; p.W = width; p.hpx = patch_x / 2; p.hpy = patch_y / 2;
  p.k = intrinsics_first;
  p.thr = rmse_threshold; p.both_invalid_rule = both_invalid_rule;
  const unsigned blocks = (unsigned)((num_rays + PATCH_WARPS - 1) / PATCH_WARPS);
  patch_reprojection_kernel<<<blocks, PATCH_WARPS * 32, 0, (cudaStream_t)stream>>>(p);
  return check_launch("srf_patch_reprojection_masks");
}

// ---------------------------------------------------------------------------------------------------------------------
// "Next" row f4: total-variation regulariser of the VM planes, forward AND gradient in one launch.
//
// Replaces TotalVariationLoss04.compute_tv_loss (src/loss_functions/TotalVariationLoss04.py:97-116) and what autograd derives
// from it: per plane [1,C,H,W]   tv = 2 * (sum (x[h+1]-x[h])^2 / numel_h + sum (x[w+1]-x[w])^2 / numel_w) * iter_weight,
// which eager PyTorch evaluates as ~25 launches per plane (2 slices, sub, pow, sum, div, and in the backward pass
// pow-backward, two slice-backward zero fills + copies and the accumulations): ~300 launches for the 12 planes of the
// shipped model pair, i.e. the training iteration is bound by the host's launch rate, not by the GPU.  Here every element is
// read once with its right and lower neighbour; the loss is accumulated in double (atomics: the rounding of the fp32
// result does not depend on the order) and the gradient  d tv / d x  is written in the same pass.
namespace srf {

struct TvPlane { const float* x; float* grad; int C, H, W; float scale_h, scale_w; };   // scale = 2 * iter_weight / numel
struct TvParams { TvPlane plane[12]; int n; double* loss; };

__global__ void __launch_bounds__(256) tv_loss_kernel(const TvParams p) {
  const TvPlane t = p.plane[blockIdx.y];
  const long long total = (long long)t.C * t.H * t.W;
  double local = 0.0;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
    const int w = (int)(e % t.W);
    const int h = (int)((e / t.W) % t.H);
    const float x = t.x[e];
    float g = 0.f;
    if (h + 1 < t.H) { const float d = t.x[e + t.W] - x; local += (double)(t.scale_h * d * d); g -= 2.f * t.scale_h * d; }
    if (h > 0) g += 2.f * t.scale_h * (x - t.x[e - t.W]);
    if (w + 1 < t.W) { const float d = t.x[e + 1] - x; local += (double)(t.scale_w * d * d); g -= 2.f * t.scale_w * d; }
    if (w > 0) g += 2.f * t.scale_w * (x - t.x[e - 1]);
    t.grad[e] = g;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(FULL, local, o);
  __shared__ double s_part[8];
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = local;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int k = 0; k < 8; ++k) s += s_part[k];
    if (s != 0.0) atomicAdd(p.loss, s);
  }
}

}  // namespace srf

SRF_API int srf_tv_loss(const float* const* planes, float* const* grads, const int* dims, int num_planes, float iter_weight,
                        double* loss, void* stream) {
  using namespace srf;
  if (num_planes == 0) return 0;
  SRF_REQUIRE(planes && grads && dims && loss, "srf_tv_loss", "null pointer");
  SRF_REQUIRE(num_planes <= 12, "srf_tv_loss", "at most 12 planes per call");
  TvParams p{};
  p.n = num_planes; p.loss = loss;
  long long largest = 0;
  for (int i = 0; i < num_planes; ++i) {
    TvPlane& t = p.plane[i];
    t.x = planes[i]; t.grad = grads[i]; t.C = dims[3 * i]; t.H = dims[3 * i + 1]; t.W = dims[3 * i + 2];
    SRF_REQUIRE(t.x && t.grad && t.C > 0 && t.H > 0 && t.W > 0, "srf_tv_loss", "bad plane");
    // numel of the difference tensors, at least 1 (:105-106); the factor 2 of matrix_tv_loss (:109) and iter_weight folded in
    const double nh = (double)t.C * (t.H - 1) * t.W, nw = (double)t.C * t.H * (t.W - 1);
    t.scale_h = (float)(2.0 * iter_weight / (nh < 1.0 ? 1.0 : nh));
    t.scale_w = (float)(2.0 * iter_weight / (nw < 1.0 ? 1.0 : nw));
    const long long total = (long long)t.C * t.H * t.W;
    if (total > largest) largest = total;
  }
  long long bx = (largest + 255) / 256;
  const long long cap = (long long)sm_count() * 4;
  if (bx > cap) bx = cap;
  if (bx < 1) bx = 1;
  tv_loss_kernel<<<dim3((unsigned)bx, (unsigned)num_planes), 256, 0, (cudaStream_t)stream>>>(p);
  return check_launch("srf_tv_loss");
}
