// Ray generation (+NDC warp, view directions), stratified depths, inverse-CDF resampling + merge.
//
// Replaces (reference file:line, relative to the upstream checkout):
//   srf_raygen            src/utils/CommonUtils04.py:73-95, :120-138, :147-149;
//                         x-flip of src/models/SimpleTensoRF09.py:205-207
//   srf_stratified_z      src/models/SimpleNeRF17.py:347-357 (== SimpleTensoRF09.py:371-381)
//   srf_sample_pdf_merge  src/models/SimpleNeRF17.py:360-371, :385-417
//
// Bit-exact contract of srf_sample_pdf_merge (indices and merged depths) is obtained by
// reproducing the evaluation order of the ATen CPU kernels the reference runs:
// 8-lane x 4-accumulator cascade row sum, IEEE division, sequential fp64 prefix sum rounded to
// fp32 per element, upper-bound search, un-fused mul/add lerp (see oracle/aten_order.py).
#include <curand_kernel.h>

#include "common.cuh"

namespace srf {

thread_local char g_last_error[512] = {0};

// ------------------------------------------------------------------------------------------
// raygen: one thread per ray, per-view camera tables in shared memory, outputs staged through
// shared memory so that every [R,3] array is written with fully coalesced 128-byte rows.
// ------------------------------------------------------------------------------------------
constexpr int RAYGEN_THREADS = 256;
constexpr int MAX_VIEWS_SMEM = 64;

struct RaygenParams {
  const int32_t* pixel_id;
  const float* k_inv;   // [F,9]
  const float* c2w;     // [F,16]
  const float* focal;   // [F,2] fx, fy
  float* rays_o;
  float* rays_d;
  float* rays_o_ndc;
  float* rays_d_ndc;
  float* view_dirs;
  long long R;
  int F;
  float height, width, near, two_near;
  int half_pixel, flip_x, ndc, viewdirs_from_ndc;
};

__global__ void __launch_bounds__(RAYGEN_THREADS) raygen_kernel(RaygenParams p) {
  __shared__ float s_cam[MAX_VIEWS_SMEM * 27];            // 9 k_inv + 16 c2w + 2 focal per view
  __shared__ float s_out[5][RAYGEN_THREADS * 3];
  const int nf = min(p.F, MAX_VIEWS_SMEM);
  for (int i = threadIdx.x; i < nf * 27; i += RAYGEN_THREADS) {
    int v = i / 27, e = i % 27;
    float val = e < 9 ? p.k_inv[v * 9 + e] : (e < 25 ? p.c2w[v * 16 + (e - 9)] : p.focal[v * 2 + (e - 25)]);
    // the NDC scale factors -1 / (w / (2 fx)), -1 / (h / (2 fy)) (CommonUtils04.py:129-134) depend on the view only: evaluated
    // once per view here (same operations, same bits) instead of two nested IEEE divisions per ray
    if (e >= 25) val = __fdiv_rn(-1.f, __fdiv_rn(e == 25 ? p.width : p.height, __fmul_rn(2.f, val)));
    s_cam[i] = val;
  }
  __syncthreads();
  const long long base = (long long)blockIdx.x * RAYGEN_THREADS;
  const long long r = base + threadIdx.x;
  if (r < p.R) {
    const int img = p.pixel_id[r * 3 + 0];
    float x = (float)p.pixel_id[r * 3 + 1];
    float y = (float)p.pixel_id[r * 3 + 2];
    if (p.half_pixel) { x += 0.5f; y += 0.5f; }
    float cam[27];
    if (img < nf) {
#pragma unroll
      for (int e = 0; e < 27; ++e) cam[e] = s_cam[img * 27 + e];
    } else {
#pragma unroll
      for (int e = 0; e < 27; ++e)
        cam[e] = e < 9 ? p.k_inv[img * 9 + e] : (e < 25 ? p.c2w[img * 16 + (e - 9)] : p.focal[img * 2 + (e - 25)]);
      cam[25] = __fdiv_rn(-1.f, __fdiv_rn(p.width, __fmul_rn(2.f, cam[25])));
      cam[26] = __fdiv_rn(-1.f, __fdiv_rn(p.height, __fmul_rn(2.f, cam[26])));
    }
    const float* ki = cam;
    const float* m = cam + 9;
    // dirs = K^-1 (x, y, 1); flip y, z  (CommonUtils04.py:86-88)
    float d0 = __fadd_rn(__fadd_rn(__fmul_rn(ki[0], x), __fmul_rn(ki[1], y)), ki[2]);
    float d1 = -__fadd_rn(__fadd_rn(__fmul_rn(ki[3], x), __fmul_rn(ki[4], y)), ki[5]);
    float d2 = -__fadd_rn(__fadd_rn(__fmul_rn(ki[6], x), __fmul_rn(ki[7], y)), ki[8]);
    // rays_d = R dirs, rays_o = t  (:92-94)
    float rd[3], ro[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      rd[i] = __fadd_rn(__fadd_rn(__fmul_rn(d0, m[i * 4 + 0]), __fmul_rn(d1, m[i * 4 + 1])), __fmul_rn(d2, m[i * 4 + 2]));
      ro[i] = m[i * 4 + 3];
    }
    if (p.flip_x) { ro[0] = -ro[0]; rd[0] = -rd[0]; }
    float vsrc[3] = {rd[0], rd[1], rd[2]};
    float on[3] = {0.f, 0.f, 0.f}, dn[3] = {0.f, 0.f, 0.f};
    if (p.ndc) {
      const float sx = cam[25], sy = cam[26];
      // CommonUtils04.py:124-134, same operation order, no FMA contraction
      const float t = __fdiv_rn(-__fadd_rn(p.near, ro[2]), rd[2]);
      const float ox = __fadd_rn(ro[0], __fmul_rn(t, rd[0]));
      const float oy = __fadd_rn(ro[1], __fmul_rn(t, rd[1]));
      const float oz = __fadd_rn(ro[2], __fmul_rn(t, rd[2]));
      on[0] = __fdiv_rn(__fmul_rn(sx, ox), oz);
      on[1] = __fdiv_rn(__fmul_rn(sy, oy), oz);
      on[2] = __fadd_rn(1.f, __fdiv_rn(p.two_near, oz));
      dn[0] = __fmul_rn(sx, __fadd_rn(__fdiv_rn(rd[0], rd[2]), -__fdiv_rn(ox, oz)));
      dn[1] = __fmul_rn(sy, __fadd_rn(__fdiv_rn(rd[1], rd[2]), -__fdiv_rn(oy, oz)));
      dn[2] = __fdiv_rn(-p.two_near, oz);
      if (p.viewdirs_from_ndc) { vsrc[0] = dn[0]; vsrc[1] = dn[1]; vsrc[2] = dn[2]; }
    }
    const float nrm = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(vsrc[0], vsrc[0]), __fmul_rn(vsrc[1], vsrc[1])), __fmul_rn(vsrc[2], vsrc[2])));
#pragma unroll
    for (int i = 0; i < 3; ++i) {
      s_out[0][threadIdx.x * 3 + i] = ro[i];
      s_out[1][threadIdx.x * 3 + i] = rd[i];
      s_out[2][threadIdx.x * 3 + i] = on[i];
      s_out[3][threadIdx.x * 3 + i] = dn[i];
      s_out[4][threadIdx.x * 3 + i] = __fdiv_rn(vsrc[i], nrm);
    }
  }
  __syncthreads();
  const long long nvalid = min((long long)RAYGEN_THREADS, p.R - base) * 3;
  float* outs[5] = {p.rays_o, p.rays_d, p.ndc ? p.rays_o_ndc : nullptr, p.ndc ? p.rays_d_ndc : nullptr, p.view_dirs};
#pragma unroll
  for (int a = 0; a < 5; ++a) {
    if (outs[a] == nullptr) continue;
    for (int i = threadIdx.x; i < nvalid; i += RAYGEN_THREADS) outs[a][base * 3 + i] = s_out[a][i];
  }
}

// ------------------------------------------------------------------------------------------
// stratified depths
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float philox_uniform(unsigned long long seed, unsigned long long idx) {
  curandStatePhilox4_32_10_t st;
  curand_init(seed, idx >> 2, 0, &st);
  float4 v = curand_uniform4(&st);                 // (0,1]
  const float u = (idx & 3) == 0 ? v.x : ((idx & 3) == 1 ? v.y : ((idx & 3) == 2 ? v.z : v.w));
  return 1.f - u;                                  // [0,1)
}

__global__ void stratified_kernel(const float* __restrict__ ladder, int S, long long total,
                                  const float* __restrict__ jitter, int use_philox,
                                  unsigned long long seed, float* __restrict__ z) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int s = (int)(i % S);
    const float c = ladder[s];
    float out = c;
    if (jitter != nullptr || use_philox) {
      const float lo = s == 0 ? c : __fmul_rn(0.5f, __fadd_rn(c, ladder[s - 1]));
      const float hi = s == S - 1 ? c : __fmul_rn(0.5f, __fadd_rn(ladder[s + 1], c));
      const float u = jitter != nullptr ? ldg_stream(jitter + i) : philox_uniform(seed, (unsigned long long)i);
      out = __fadd_rn(lo, __fmul_rn(__fadd_rn(hi, -lo), u));
    }
    z[i] = out;
  }
}

// ------------------------------------------------------------------------------------------
// box-march depths (Simple-TensoRF without NDC): z[r, s] = clamp(t_entry(r), near, far) + step_size * (s + jitter[r])
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) box_march_kernel(const float* __restrict__ rays_o, const float* __restrict__ rays_d,
                                                        const float* __restrict__ jitter, float b0x, float b0y, float b0z,
                                                        float b1x, float b1y, float b1z, float near, float far, float step_size,
                                                        int S, long long R, float* __restrict__ z) {
  const long long total = R * S;
  const float b0[3] = {b0x, b0y, b0z}, b1[3] = {b1x, b1y, b1z};
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long r = i / S;
    const int s = (int)(i - r * S);
    float t_min = -__int_as_float(0x7f800000);
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      const float o = __ldg(rays_o + r * 3 + a), d = __ldg(rays_d + r * 3 + a);
      const float vec = d == 0.f ? 1e-6f : d;
      const float ra = __fdiv_rn(__fadd_rn(b1[a], -o), vec), rb = __fdiv_rn(__fadd_rn(b0[a], -o), vec);
      t_min = fmaxf(t_min, fminf(ra, rb));
    }
    t_min = fminf(fmaxf(t_min, near), far);
    float rng = (float)s;
    if (jitter != nullptr) rng = __fadd_rn(rng, __ldg(jitter + r));
    z[i] = __fadd_rn(t_min, __fmul_rn(step_size, rng));
  }
}

// ------------------------------------------------------------------------------------------
// sample_pdf + merge: one warp per ray.
// shared memory per warp (floats): zc[S] | cdf[S] (first holds w/pdf) | bins[S] | sortbuf[npad]
// ------------------------------------------------------------------------------------------
constexpr int PDF_WARPS = 8;

// ATen vectorized_inner_sum order (8 lanes x 4 ILP accumulators x 4 cascade levels); executed by
// lanes 0..7 of the warp, each playing one SIMD lane.  Returns the row sum in every lane.
__device__ float aten_row_sum(const float* w, int n) {
  const int lane = lane_id();
  float total = 0.f;
  if (n < 8) {   // ATen scalar_inner_sum -> row_sum: four ILP partial sums, left-overs into the first
    float part[4] = {0.f, 0.f, 0.f, 0.f};
    const int full = n >> 2;
    for (int i = 0; i < full; ++i)
#pragma unroll
      for (int b = 0; b < 4; ++b) part[b] = __fadd_rn(part[b], w[i * 4 + b]);
    for (int k = full * 4; k < n; ++k) part[0] = __fadd_rn(part[0], w[k]);
    return __fadd_rn(__fadd_rn(__fadd_rn(part[0], part[1]), part[2]), part[3]);
  }
  const int nvec = n >> 3;
  const int groups = nvec >> 2;
  int clog = 0;
  while ((1 << clog) < groups) ++clog;
  const int power = max(4, clog / 4);
  const int step = 1 << power;
  float part0 = 0.f;
  if (lane < 8) {
    float acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
    int i = 0;
    while (i + step <= groups) {
      for (int j = 0; j < step; ++j, ++i) {
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[0][b] = __fadd_rn(acc[0][b], w[((i * 4 + b) << 3) + lane]);
      }
#pragma unroll
      for (int j = 1; j < 4; ++j) {
#pragma unroll
        for (int b = 0; b < 4; ++b) { acc[j][b] = __fadd_rn(acc[j][b], acc[j - 1][b]); acc[j - 1][b] = 0.f; }
        if ((i & ((step - 1) << (j * power))) != 0) break;
      }
    }
    for (; i < groups; ++i) {
#pragma unroll
      for (int b = 0; b < 4; ++b) acc[0][b] = __fadd_rn(acc[0][b], w[((i * 4 + b) << 3) + lane]);
    }
#pragma unroll
    for (int j = 1; j < 4; ++j)
#pragma unroll
      for (int b = 0; b < 4; ++b) acc[0][b] = __fadd_rn(acc[0][b], acc[j][b]);
    for (int k = groups * 4; k < nvec; ++k) acc[0][0] = __fadd_rn(acc[0][0], w[(k << 3) + lane]);
    part0 = __fadd_rn(__fadd_rn(__fadd_rn(acc[0][0], acc[0][1]), acc[0][2]), acc[0][3]);
  }
  for (int k = nvec << 3; k < n; ++k) total = __fadd_rn(total, w[k]);
#pragma unroll
  for (int k = 0; k < 8; ++k) total = __fadd_rn(total, __shfl_sync(FULL, part0, k));
  return total;
}

struct PdfParams {
  const float* z_coarse;   // [R,S]
  const float* weights;    // [R,S]
  const float* u;          // [R,N] or one shared row [N] (u_row_stride == 0) or nullptr (philox)
  float* z_fine;           // [R,S+N]
  float* samples;          // [R,N] or nullptr
  long long* below;        // [R,N] or nullptr
  long long* above;        // [R,N] or nullptr
  long long R;
  long long u_row_stride;
  unsigned long long seed;
  int S, N, npad;
};

// ascending bitonic sort of n (a power of two) floats in shared memory by one warp; j, k are powers of two, so pair
// indices are built with shifts (no integer division on the hot loop)
__device__ __forceinline__ void bitonic_sort_smem(float* v, int n, int lane) {
  for (int k = 2, lk = 1; k <= n; k <<= 1, ++lk) {
    for (int lj = lk - 1; lj >= 0; --lj) {
      const int j = 1 << lj;
      for (int q = lane; q < (n >> 1); q += 32) {
        const int i = ((q >> lj) << (lj + 1)) | (q & (j - 1));
        const int l = i | j;
        const float a = v[i], b = v[l];
        const bool up = (i & k) == 0;
        if ((a > b) == up) { v[i] = b; v[l] = a; }
      }
      __syncwarp();
    }
  }
}

__global__ void __launch_bounds__(PDF_WARPS * 32) sample_pdf_merge_kernel(PdfParams p) {
  extern __shared__ float smem[];
  const int warp = threadIdx.x >> 5, lane = lane_id();
  const long long r = (long long)blockIdx.x * PDF_WARPS + warp;
  if (r >= p.R) return;
  const int S = p.S, N = p.N, npad = p.npad;
  float* zc = smem + (size_t)warp * (3 * S + npad);
  float* cdf = zc + S;
  float* bins = cdf + S;
  float* sbuf = bins + S;
  const int n = S - 2;         // interior weights
  const int nb = S - 1;        // bins == cdf entries

  for (int i = lane; i < S; i += 32) {
    const float zi = ldg_stream(p.z_coarse + r * S + i);
    zc[i] = zi;
    sbuf[i] = zi;
    if (i >= 1 && i <= n) cdf[i - 1] = __fadd_rn(ldg_stream(p.weights + r * S + i), 1e-5f);
  }
  for (int i = S + N + lane; i < npad; i += 32) sbuf[i] = __int_as_float(0x7f800000);
  __syncwarp();
  const float total = aten_row_sum(cdf, n);
  __syncwarp();
  for (int i = lane; i < nb; i += 32) bins[i] = __fmul_rn(0.5f, __fadd_rn(zc[i + 1], zc[i]));
  // pdf, then ATen's cumsum (sequential fp64 accumulation, each prefix rounded to fp32), shifted by one for the leading 0.
  // The pdf values are fp32 numbers that sum to ~1: when the smallest one is >= 2^-28 every fp64 partial sum is EXACT
  // (48 significant bits at most), so the additions may be re-associated without changing a single bit and the prefix
  // sum runs as a warp scan.  Otherwise (weights far outside [0, 1]) one lane reproduces the sequential order.
  float pmin = __int_as_float(0x7f800000);
  for (int i = lane; i < n; i += 32) {
    const float pv = __fdiv_rn(cdf[i], total);
    cdf[i] = pv;
    pmin = fminf(pmin, pv);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) pmin = fminf(pmin, __shfl_xor_sync(FULL, pmin, o));
  __syncwarp();
  constexpr int MAXC = 8;
  const int per = (n + 31) >> 5;
  if (per <= MAXC && pmin >= 3.7252903e-09f) {
    float pv[MAXC];
    const int i0 = lane * per;
    double tot = 0.0;
#pragma unroll
    for (int q = 0; q < MAXC; ++q) {
      pv[q] = (q < per && i0 + q < n) ? cdf[i0 + q] : 0.f;
      tot += (double)pv[q];
    }
    double incl = tot;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double t = __shfl_up_sync(FULL, incl, o);
      if (lane >= o) incl += t;
    }
    double acc = incl - tot;
    __syncwarp();                            // every lane holds its pdf values in registers before cdf is overwritten
#pragma unroll
    for (int q = 0; q < MAXC; ++q) {
      if (q < per && i0 + q < n) {
        acc += (double)pv[q];
        cdf[i0 + q + 1] = (float)acc;
      }
    }
    if (lane == 0) cdf[0] = 0.f;
  } else if (lane == 0) {
    double acc = 0.0;
    float prev = 0.f;                     // cdf[0]
    for (int i = 0; i < n; ++i) {
      acc += (double)cdf[i];
      const float cur = (float)acc;       // round-to-nearest fp32 of the fp64 prefix
      cdf[i] = prev;
      prev = cur;
    }
    cdf[n] = prev;
  }
  __syncwarp();

  for (int j = lane; j < N; j += 32) {
    float u;
    if (p.u != nullptr) u = p.u[r * p.u_row_stride + j];
    else u = philox_uniform(p.seed, (unsigned long long)(r * N + j));
    int lo = 0, hi = nb;                  // upper bound: first index with cdf > u
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (cdf[mid] <= u) lo = mid + 1; else hi = mid;
    }
    const int b = max(lo - 1, 0);
    const int a = min(lo, nb - 1);
    const float cb = cdf[b];
    float denom = __fadd_rn(cdf[a], -cb);
    if (denom < 1e-5f) denom = 1.f;
    const float t = __fdiv_rn(__fadd_rn(u, -cb), denom);
    const float bb = bins[b];
    const float smp = __fadd_rn(bb, __fmul_rn(t, __fadd_rn(bins[a], -bb)));
    sbuf[S + j] = smp;
    if (p.samples) p.samples[r * N + j] = smp;
    if (p.below) p.below[r * N + j] = b;
    if (p.above) p.above[r * N + j] = a;
  }
  __syncwarp();
  // merge.  Both runs are normally sorted already (coarse depths by construction, the new samples whenever u is sorted:
  // the inverse CDF is monotone), so each element's final position is its own index plus its rank in the other run
  // (binary searches; ties: coarse depths first - equal values make the order immaterial for the sorted VALUES the
  // reference's torch.sort returns).  Unsorted input (random u in training) takes the bitonic network below.
  constexpr int MAXQ = 8;
  const bool small = (S + 31) / 32 <= MAXQ && (N + 31) / 32 <= MAXQ;
  bool z_sorted = true, s_sorted = true;
  for (int i = lane; i < S - 1; i += 32) z_sorted = z_sorted && (sbuf[i] <= sbuf[i + 1]);
  for (int j = lane; j < N - 1; j += 32) s_sorted = s_sorted && (sbuf[S + j] <= sbuf[S + j + 1]);
  z_sorted = __all_sync(FULL, z_sorted);
  s_sorted = __all_sync(FULL, s_sorted);
  int np2 = 2;
  while (np2 < N) np2 <<= 1;
  if (small && z_sorted && !s_sorted && S + np2 <= npad) {
    bitonic_sort_smem(sbuf + S, np2, lane);           // random u (training): sort the new samples only (+inf padded)
    s_sorted = true;
  }
  if (small && z_sorted && s_sorted) {
    float vz[MAXQ], vs[MAXQ];
    int pz[MAXQ], ps[MAXQ];
#pragma unroll
    for (int q = 0; q < MAXQ; ++q) {
      const int i = lane + 32 * q;
      pz[q] = -1; ps[q] = -1; vz[q] = 0.f; vs[q] = 0.f;
      if (i < S) {                                    // coarse depth i: rank = number of samples strictly below it
        const float v = sbuf[i];
        int lo = 0, hi = N;
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          if (sbuf[S + mid] < v) lo = mid + 1; else hi = mid;
        }
        vz[q] = v; pz[q] = i + lo;
      }
      if (i < N) {                                    // sample i: rank = number of coarse depths <= it
        const float v = sbuf[S + i];
        int lo = 0, hi = S;
        while (lo < hi) {
          const int mid = (lo + hi) >> 1;
          if (sbuf[mid] <= v) lo = mid + 1; else hi = mid;
        }
        vs[q] = v; ps[q] = i + lo;
      }
    }
    __syncwarp();                                     // all reads done before the scatter overwrites sbuf
#pragma unroll
    for (int q = 0; q < MAXQ; ++q) {
      if (pz[q] >= 0) sbuf[pz[q]] = vz[q];
      if (ps[q] >= 0) sbuf[ps[q]] = vs[q];
    }
    __syncwarp();
  } else {
    bitonic_sort_smem(sbuf, npad, lane);              // coarse depths + new samples + inf padding
  }
  const int M = S + N;
  for (int i = lane; i < M; i += 32) p.z_fine[r * M + i] = sbuf[i];
}

}  // namespace srf

using namespace srf;

SRF_API const char* srf_last_error(void) { return g_last_error; }

SRF_API int srf_abi_version(void) { return 1; }

SRF_API int srf_raygen(const int32_t* pixel_id, int64_t num_rays, const float* k_inv, const float* c2w,
                       const float* focal, int num_views, int height, int width, float near, float two_near,
                       int half_pixel, int flip_x, int ndc, int viewdirs_from_ndc, float* rays_o, float* rays_d,
                       float* rays_o_ndc, float* rays_d_ndc, float* view_dirs, void* stream) {
  if (num_rays == 0) return 0;
  SRF_REQUIRE(pixel_id && k_inv && c2w && focal && rays_o && rays_d && view_dirs, "srf_raygen", "null pointer");
  SRF_REQUIRE(!ndc || (rays_o_ndc && rays_d_ndc), "srf_raygen", "ndc outputs required when ndc != 0");
  SRF_REQUIRE(num_views > 0 && num_rays >= 0, "srf_raygen", "bad sizes");
  RaygenParams p{pixel_id, k_inv, c2w, focal, rays_o, rays_d, rays_o_ndc, rays_d_ndc, view_dirs, num_rays, num_views,
                 (float)height, (float)width, near, two_near, half_pixel, flip_x, ndc, viewdirs_from_ndc};
  const unsigned blocks = (unsigned)((num_rays + RAYGEN_THREADS - 1) / RAYGEN_THREADS);
  raygen_kernel<<<blocks, RAYGEN_THREADS, 0, (cudaStream_t)stream>>>(p);
  return check_launch("srf_raygen");
}

SRF_API int srf_stratified_z(const float* ladder, int num_samples, int64_t num_rays, const float* jitter,
                             int use_philox, uint64_t seed, float* z, void* stream) {
  if (num_rays == 0) return 0;
  SRF_REQUIRE(ladder && z, "srf_stratified_z", "null pointer");
  SRF_REQUIRE(num_samples > 0 && num_rays >= 0, "srf_stratified_z", "bad sizes");
  const long long total = (long long)num_rays * num_samples;
  const int threads = 256;
  long long blocks = (total + threads - 1) / threads;
  const long long cap = (long long)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  stratified_kernel<<<(unsigned)blocks, threads, 0, (cudaStream_t)stream>>>(ladder, num_samples, total, jitter,
                                                                             use_philox, seed, z);
  return check_launch("srf_stratified_z");
}

SRF_API int srf_box_march_z(const float* rays_o, const float* rays_d, int64_t num_rays, int num_samples, const float* bbox,
                            float near, float far, float step_size, const float* jitter, float* z, void* stream) {
  if (num_rays == 0) return 0;
  SRF_REQUIRE(rays_o && rays_d && bbox && z, "srf_box_march_z", "null pointer");
  SRF_REQUIRE(num_samples > 0 && num_rays >= 0, "srf_box_march_z", "bad sizes");
  const long long total = (long long)num_rays * num_samples;
  long long blocks = (total + 255) / 256;
  const long long cap = (long long)sm_count() * 16;
  if (blocks > cap) blocks = cap;
  box_march_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(rays_o, rays_d, jitter, bbox[0], bbox[1], bbox[2], bbox[3], bbox[4],
                                                                       bbox[5], near, far, step_size, num_samples, num_rays, z);
  return check_launch("srf_box_march_z");
}

SRF_API int srf_sample_pdf_merge(const float* z_coarse, const float* weights, const float* u, int64_t u_row_stride,
                                 uint64_t seed, int64_t num_rays, int num_coarse, int num_fine, float* z_fine,
                                 float* samples, int64_t* below, int64_t* above, void* stream) {
  if (num_rays == 0) return 0;
  SRF_REQUIRE(z_coarse && weights && z_fine, "srf_sample_pdf_merge", "null pointer");
  SRF_REQUIRE(num_coarse >= 3 && num_fine >= 1 && num_rays >= 0, "srf_sample_pdf_merge", "need S >= 3, N >= 1");
  int npad = 2;
  while (npad < num_coarse + num_fine) npad <<= 1;
  const size_t smem = (size_t)PDF_WARPS * (3 * num_coarse + npad) * sizeof(float);
  SRF_REQUIRE(smem <= 200 * 1024, "srf_sample_pdf_merge", "S + N too large for one warp's shared-memory slice");
  static size_t configured = 0;
  if (smem > 48 * 1024 && smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(sample_pdf_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail("srf_sample_pdf_merge", cudaGetErrorString(e));
    configured = smem;
  }
  PdfParams p{z_coarse, weights, u, z_fine, samples, (long long*)below, (long long*)above, num_rays,
              (long long)u_row_stride, seed, num_coarse, num_fine, npad};
  const unsigned blocks = (unsigned)((num_rays + PDF_WARPS - 1) / PDF_WARPS);
  sample_pdf_merge_kernel<<<blocks, PDF_WARPS * 32, smem, (cudaStream_t)stream>>>(p);
  return check_launch("srf_sample_pdf_merge");
}
