// Data-gradient chain of the fused NeRF MLP on tcgen05 tensor cores: from (d loss / d sigma, d loss / d rgb)
// back through the heads, the view branch, the feature layer and the 8x256 trunk, one persistent CTA per SM,
// 128 samples per tile — the mirror image of nerf_mlp.cu.
//
// Replaces what autograd derives for src/models/SimpleNeRF17.py:726-785 (MLP trunk / heads).  Per tile:
//   init warps : d_o = g_rgb * rgb * (1 - rgb), d_sigma_raw = g_sigma * [sigma > 0]; the gradient entering the top
//                hidden layer, dZ_top = (sum_c d_o[c] W_head[c,:]) * relu-mask, as bf16 A-operand K blocks; the head
//                gradient images [d_o | d_sigma_raw] for the weight-gradient kernel
//   MMA thread : dX = dZ * W  as  tcgen05.mma with the TRANSPOSED weights streamed as pre-swizzled images
//   epilogue   : tcgen05.ld -> (+ d_sigma_raw * w_sigma for the layer under the sigma head) -> bf16 -> AND with the
//                ReLU mask (one bit per value, written by the forward next to the activation tiles) -> next layer's A
//                operand, written back into TENSOR MEMORY over the accumulator columns just drained (tcgen05.st; the
//                next layer's MMAs take A from TMEM), AND the dZ tile image for srf_nerf_mlp_wgrad, staged in shared
//                memory and written to HBM by 16 KB bulk async copies.
// Sample points and encodings carry no gradient (z_samples.detach(), frozen cameras), so the chain stops at layer 1.
#include "common.cuh"
#include "tcgen05.cuh"

namespace srf {

constexpr int DG_MAX_LAYERS = 12;

struct DgradLayer {           // mirrors srf_dgrad_layer in include/simple_rf_b200.h
  int32_t num_kblocks;        // 64-wide K blocks of the incoming gradient (output width of the forward layer / 64)
  int32_t mask_slot;          // saved-activation slot (4 images) whose non-zero pattern is the ReLU mask of this layer's output, or -1
  int32_t rank1_offset;       // >= 0: side offset of w_sigma[256]; adds d_sigma_raw[row] * w_sigma[col] before the mask
  int32_t dz_slot;            // first dz image slot of the output (n_out / 64 images), or -1: not written
  int32_t n_out;              // output width of this backward layer: 128 or 256 (input width of the forward layer, padded)
  int32_t rows_cols;          // > 0 (last layer only): the first rows_cols output columns are ALSO written as fp32 rows (input gradient)
  int64_t weight_offset;      // byte offset into the transposed-weight blob
};

struct DgradProgram {
  int32_t num_layers, num_fwd_layers;
  int32_t top_width;          // 128 or 256
  int32_t top_mask_slot;      // saved-activation slot of the top hidden layer (its non-zero pattern gates dZ_top)
  int32_t top_slot;           // dz slot of dZ_top (top_width / 64 images)
  int32_t head_slot;          // dz slots [head_slot, head_slot + 2): head-gradient images
  int32_t head_kind;          // 1: 3-row rgb head (d_o in image 0, d_sigma_raw in image 1 col 0); 2: 4-row head [sigma, r, g, b]
  int32_t head_w_offset;      // side offset of the head weights [rows][top_width]
  int32_t side_count;
  int32_t pad_;
  DgradLayer layers[DG_MAX_LAYERS];
};

struct DgradArgs {
  const uint8_t* weights_t;   // transposed-weight images
  const float* side;
  const uint8_t* acts;        // saved activation tiles of the forward [tile][act_slots + mask images][16 KB]; only the masks are read
  int act_slots;
  const float* sigma; const float* rgb;       // forward outputs [M], [M,3]
  const float* g_sigma; const float* g_rgb;   // upstream gradients [M], [M,3] (nullable)
  uint8_t* dz;                // [tile][dz_slots][16 KB]
  int dz_slots;
  float* g_rows;              // [M, g_row_pitch] fp32 input gradient of the chain's last layer, or nullptr
  int g_row_pitch;
  long long total;
  const int* count;           // device-side row count (<= total), or nullptr: the chain stops there (rows beyond carry no gradient)
};

// SRF_MLP_TRACE: CTA 0 records clock64() at pipeline events of its 3rd tile (tools/dgrad_trace.py)
#ifndef SRF_MLP_TRACE
#define SRF_MLP_TRACE 0
#endif
#if SRF_MLP_TRACE
__device__ long long g_dg_trace[2048];
#define DTRACE(slot) do { if (blockIdx.x == 0 && t == 2) g_dg_trace[(slot)] = clock64(); } while (0)
#else
#define DTRACE(slot) do { } while (0)
#endif
constexpr int DG_GROUPS = 2;
constexpr int DG_COLS = 32;
constexpr int DG_EPI_WARP0 = 2;
constexpr int DG_EPI_THREADS = 128 * DG_GROUPS;
constexpr int DG_INIT_WARP0 = DG_EPI_WARP0 + 4 * DG_GROUPS;
constexpr int DG_THREADS = 32 * (DG_INIT_WARP0 + 4);
constexpr int DG_KBLOCK = 128 * 128;
constexpr int DG_STAGE = 2 * DG_KBLOCK;
constexpr int DG_STAGES = 3;
constexpr int DG_RING_MAX = 2 * DG_STAGES;     // resident mode: one 16 KB half-stage per K block of the tile
constexpr int DG_OUT_BUFS = 3;             // staging images of the dZ tiles on their way to HBM
constexpr int DG_MAX_SIDE = 2048;

struct alignas(1024) DgradSmem {
  uint8_t h[4][DG_KBLOCK];                 // dZ_top (A operand of the first backward layer; also the source of its HBM copy)
  uint8_t w[DG_STAGES][DG_STAGE];
  uint8_t out[DG_OUT_BUFS][DG_KBLOCK];
  float side[DG_MAX_SIDE];
  float dsig[2][128];                      // d_sigma_raw per row, double-buffered over tiles (the init warps run a tile ahead)
  uint64_t w_full[DG_RING_MAX], w_empty[DG_RING_MAX];
  uint64_t a_ready[4];        // K block of the next layer's A operand written to tensor memory by the epilogue
  uint64_t top_ready;         // H blocks written by the init warps
  uint64_t h_free;            // every MMA of the tile's FIRST layer (the only reader of `h`) has retired
  uint64_t d_full[2];
  uint32_t tmem_base;
};

__device__ __forceinline__ void dg_epi_bar() { asm volatile("bar.sync 1, %0;" ::"n"(DG_EPI_THREADS) : "memory"); }
__device__ __forceinline__ void dg_init_bar() { asm volatile("bar.sync 2, 128;" ::: "memory"); }

// 2 mask bits -> 0xFFFF per set bit (low / high bf16 of a packed pair)
__device__ __forceinline__ uint32_t pair_mask(uint32_t bits, int j) {
  const uint32_t b = bits >> (2 * j);
  return ((b & 1u) ? 0x0000FFFFu : 0u) | ((b & 2u) ? 0xFFFF0000u : 0u);
}

__global__ void __launch_bounds__(DG_THREADS, 1) nerf_mlp_dgrad_kernel(const __grid_constant__ DgradProgram prog,
                                                                       const DgradArgs args) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  DgradSmem& sm = *reinterpret_cast<DgradSmem*>(smem_raw);
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;   // warp-uniform for the compiler
  if ((ptx::smem_u32(smem_raw) & 1023u) != 0) __trap();
  for (int i = threadIdx.x; i < prog.side_count; i += DG_THREADS) sm.side[i] = args.side[i];
  if (warp == 0 && lane == 0) {
    for (int s = 0; s < DG_RING_MAX; ++s) { ptx::mbar_init(&sm.w_full[s], 1); ptx::mbar_init(&sm.w_empty[s], 1); }
    for (int r = 0; r < 4; ++r) ptx::mbar_init(&sm.a_ready[r], 4 * DG_GROUPS);
    ptx::mbar_init(&sm.top_ready, 4);
    ptx::mbar_init(&sm.h_free, 1);
    ptx::mbar_init(&sm.d_full[0], 1);
    ptx::mbar_init(&sm.d_full[1], 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc(&sm.tmem_base, 512);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = sm.tmem_base;
  long long total_rows = args.total;
  if (args.count != nullptr) { const long long c = *args.count; total_rows = c < total_rows ? c : total_rows; }
  const int num_tiles = (int)((total_rows + 127) / 128);
  const int my_tiles = num_tiles > (int)blockIdx.x ? (num_tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;
  const int NL = prog.num_layers;
  const size_t act_stride = (size_t)act_tile_images(args.act_slots) * DG_KBLOCK;
  // Resident weights (the TensoRF colour MLP: two 128-wide layers, four 16 KB transposed images): when every layer is 128 wide and the
  // tile's K blocks fit the ring as 16 KB half-stages, step s of a tile always uses half-stage s - the images are copied once per CTA
  // and the ring protocol keeps cycling without copies (a 3-stage ring of 32 KB stages made the second K block of every layer wait
  // ~750 cycles for its image, tools/dgrad_rows_trace.py)
  int steps_per_tile = 0;
  bool narrow = true;
  for (int l = 0; l < NL; ++l) { steps_per_tile += prog.layers[l].num_kblocks; narrow = narrow && prog.layers[l].n_out == 128; }
  const bool resident = narrow && steps_per_tile <= DG_RING_MAX;
  const uint32_t ring = resident ? (uint32_t)steps_per_tile : (uint32_t)DG_STAGES;
  const uint32_t stage_bytes = resident ? (uint32_t)DG_KBLOCK : (uint32_t)DG_STAGE;

  if (warp == 0) {
    // ------------------------------------------------------------ transposed-weight producer
    if (lane == 0) {
      uint32_t it = 0;
      const uint64_t keep = ptx::l2_policy_evict_last();      // the transposed weights are re-read per tile: keep them in L2
      for (int t = 0; t < my_tiles; ++t)
        for (int l = 0; l < NL; ++l) {
          const DgradLayer& L = prog.layers[l];
          for (int kb = 0; kb < L.num_kblocks; ++kb, ++it) {
            const uint32_t st = it % ring, ph = (it / ring) & 1;
            const int halves = L.n_out >> 7;
            ptx::mbar_wait(&sm.w_empty[st], ph ^ 1);
            if (resident && t > 0) { ptx::mbar_arrive(&sm.w_full[st]); continue; }       // the image is already there
            ptx::mbar_arrive_expect_tx(&sm.w_full[st], halves * DG_KBLOCK);
            for (int nh = 0; nh < halves; ++nh)
              ptx::bulk_g2s_hint(&sm.w[0][0] + (size_t)st * stage_bytes + nh * DG_KBLOCK,
                                 args.weights_t + L.weight_offset + (size_t)(nh * L.num_kblocks + kb) * DG_KBLOCK, DG_KBLOCK, &sm.w_full[st], keep);
          }
        }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer: converged warp, one elected lane issues
    // (uniform control flow keeps the descriptors in uniform registers, see nerf_mlp.cu).  The first layer reads dZ_top from
    // shared memory; every later layer reads the gradient the epilogue packed back into tensor memory.
    const uint32_t desc_hi = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);
    const uint32_t h_lo = (ptx::smem_u32(sm.h[0]) >> 4) & 0x3FFF;
    const uint32_t w_lo = (ptx::smem_u32(sm.w[0]) >> 4) & 0x3FFF;
    uint32_t st = 0, ph = 0, layer_count = 0, a_phase = 0;
    for (int t = 0; t < my_tiles; ++t) {
      ptx::mbar_wait(&sm.top_ready, t & 1);
      for (int l = 0; l < NL; ++l, ++layer_count) {
        const DgradLayer& L = prog.layers[l];
        const uint32_t buf = layer_count & 1;
        const uint32_t d_addr = tmem + buf * 256;
        const uint32_t idesc = ptx::make_idesc_bf16(128, (uint32_t)L.n_out);
        const int nkb = L.num_kblocks;
        for (int kb = 0; kb < nkb; ++kb) {
          if (l > 0) {
            ptx::mbar_wait(&sm.a_ready[kb], (a_phase >> kb) & 1);
            a_phase ^= 1u << kb;
          }
          if (lane == 0) DTRACE(16 + (l * 4 + kb) * 4 + 0);
          ptx::mbar_wait(&sm.w_full[st], ph);
          if (lane == 0) DTRACE(16 + (l * 4 + kb) * 4 + 1);
          ptx::tc_fence_after();
          const uint32_t issue = ptx::elect_one();
          if (l > 0)
            ptx::umma4_bf16_ts_if(issue, d_addr, tmem + (buf ^ 1u) * 256 + (uint32_t)kb * 64, 32u, w_lo + st * (stage_bytes >> 4), desc_hi,
                                  idesc, kb == 0 ? 0u : 1u, 4u);
          else
            ptx::umma4_bf16_if(issue, d_addr, h_lo + (uint32_t)kb * (DG_KBLOCK >> 4), w_lo + st * (stage_bytes >> 4), desc_hi, idesc,
                               kb == 0 ? 0u : 1u, 4u);
          ptx::umma_commit_if(issue, &sm.w_empty[st]);
          if (kb == nkb - 1) {
            ptx::umma_commit_if(issue, &sm.d_full[buf]);
            if (l == 0) ptx::umma_commit_if(issue, &sm.h_free);
          }
          if (lane == 0) DTRACE(16 + (l * 4 + kb) * 4 + 2);
          if (++st == ring) { st = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp >= DG_INIT_WARP0) {
    // ------------------------------------------------------------ init warps: heads -> dZ_top, head-gradient images
    const int row = (warp - DG_INIT_WARP0) * 32 + lane;
    const bool leader = threadIdx.x == DG_INIT_WARP0 * 32;
    const int tw = prog.top_width;
    const float* hw = sm.side + prog.head_w_offset;
    for (int t = 0; t < my_tiles; ++t) {
      const long long tile = (long long)blockIdx.x + (long long)t * gridDim.x;
      const long long m = tile * 128 + row;
      const bool valid = m < total_rows;
      float d_o[4] = {0.f, 0.f, 0.f, 0.f};            // head_kind 1: [r, g, b, -]; head_kind 2: [sigma, r, g, b]
      float dsig = 0.f;
      if (valid) {
        if (args.g_sigma != nullptr && args.sigma[m] > 0.f) dsig = args.g_sigma[m];
        if (args.g_rgb != nullptr) {
          const int b = prog.head_kind == 2 ? 1 : 0;
#pragma unroll
          for (int c = 0; c < 3; ++c) { const float y = args.rgb[m * 3 + c]; d_o[b + c] = args.g_rgb[m * 3 + c] * y * (1.f - y); }
        }
        if (prog.head_kind == 2) d_o[0] = dsig;
      }
      const int head_rows = prog.head_kind == 2 ? 4 : 3;
      // head-gradient images (image 0: d_o in the first columns; image 1: d_sigma_raw in column 0), rest zero
      {
        uint8_t* img = args.dz + ((size_t)tile * args.dz_slots + prog.head_slot) * DG_KBLOCK;
        const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
        for (int u = 1; u < 8; ++u) {
          *reinterpret_cast<uint4*>(img + ptx::sw128_offset(row, u)) = zero;
          *reinterpret_cast<uint4*>(img + DG_KBLOCK + ptx::sw128_offset(row, u)) = zero;
        }
        *reinterpret_cast<uint4*>(img + ptx::sw128_offset(row, 0)) = make_uint4(ptx::pack_bf16(d_o[0], d_o[1]), ptx::pack_bf16(d_o[2], d_o[3]), 0u, 0u);
        *reinterpret_cast<uint4*>(img + DG_KBLOCK + ptx::sw128_offset(row, 0)) = make_uint4(ptx::pack_bf16(dsig, 0.f), 0u, 0u, 0u);
      }
      // ReLU mask of the top hidden layer: one word per 32 columns
      const uint8_t* mrec = args.acts + (size_t)tile * act_stride;
      uint32_t mbits[8];
#pragma unroll
      for (int q = 0; q < 8; ++q)
        mbits[q] = q < tw / 32 ? __ldg(reinterpret_cast<const uint32_t*>(mrec + act_mask_offset(args.act_slots, prog.top_mask_slot + (q >> 1)) +
                                                                          (q & 1) * 512 + row * 4))
                               : 0u;
      // `h` of the previous tile: read by its first layer's MMAs and by the bulk copies of its dZ_top images
      if (leader) ptx::bulk_wait_read<0>();
      dg_init_bar();
      ptx::mbar_wait(&sm.h_free, (t & 1) ^ 1);
      sm.dsig[t & 1][row] = dsig;
#pragma unroll
      for (int q = 0; q < 8; ++q) {                    // 32 columns = four 16-byte units at a time
        if (q >= tw / 32) break;
#pragma unroll
        for (int uu = 0; uu < 4; ++uu) {
          const int u = q * 4 + uu;
          float v[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int col = u * 8 + j;
            float a = 0.f;
            for (int c = 0; c < head_rows; ++c) a = fmaf(d_o[c], hw[c * tw + col], a);
            v[j] = a;
          }
          const uint32_t off = ptx::sw128_offset(row, u & 7);
          const uint32_t bq = mbits[q] >> (8 * uu);
          *reinterpret_cast<uint4*>(sm.h[u >> 3] + off) =
              make_uint4(ptx::pack_bf16(v[0], v[1]) & pair_mask(bq, 0), ptx::pack_bf16(v[2], v[3]) & pair_mask(bq, 1),
                         ptx::pack_bf16(v[4], v[5]) & pair_mask(bq, 2), ptx::pack_bf16(v[6], v[7]) & pair_mask(bq, 3));
        }
      }
      ptx::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&sm.top_ready);
      // the same images go to HBM (operands of the weight-gradient kernel) as bulk copies
      dg_init_bar();
      if (leader) {
        uint8_t* top = args.dz + ((size_t)tile * args.dz_slots + prog.top_slot) * DG_KBLOCK;
        ptx::bulk_s2g(top, sm.h[0], (uint32_t)(tw / 64) * DG_KBLOCK);
        ptx::bulk_commit();
      }
    }
    if (leader) ptx::bulk_wait_all();
  } else {
    // ------------------------------------------------------------ epilogue warps
    const int quarter = warp & 3;
    const int grp = (warp - DG_EPI_WARP0) >> 2;
    const int row = quarter * 32 + lane;
    uint32_t layer_count = 0, d_phase = 0, out_count = 0;
    uint32_t mb_next[4] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu};
    if (my_tiles > 0 && prog.layers[0].mask_slot >= 0) {      // layer 0 of the first tile
      const uint8_t* rec0 = args.acts + (size_t)blockIdx.x * act_stride;
#pragma unroll
      for (int kb = 0; kb < 4; ++kb)
        if (kb < (prog.layers[0].n_out >> 6))
          mb_next[kb] = __ldg(reinterpret_cast<const uint32_t*>(rec0 + act_mask_offset(args.act_slots, prog.layers[0].mask_slot + kb) + grp * 512 + row * 4));
    }
    for (int t = 0; t < my_tiles; ++t) {
      const long long tile = (long long)blockIdx.x + (long long)t * gridDim.x;
      const uint8_t* mrec = args.acts + (size_t)tile * act_stride;
      for (int l = 0; l < NL; ++l, ++layer_count) {
        const DgradLayer& L = prog.layers[l];
        const uint32_t buf = layer_count & 1;
        const uint32_t t_row = tmem + ((uint32_t)(quarter * 32) << 16) + buf * 256 + grp * DG_COLS;
        const bool last = l == NL - 1;
        const float* wsig = L.rank1_offset >= 0 ? sm.side + L.rank1_offset : nullptr;
        uint8_t* out = args.dz + ((size_t)tile * args.dz_slots + L.dz_slot) * DG_KBLOCK;
        // the mask words this thread applies were requested ONE LAYER AHEAD (below): a request issued only when this layer starts
        // reaches its first use ~400 cycles later and leaves ~1300 cycles of global-memory latency exposed per layer (in-kernel trace)
        const int out_blocks = L.n_out >> 6;
        uint32_t mb[4];
#pragma unroll
        for (int kb = 0; kb < 4; ++kb) mb[kb] = mb_next[kb];
        {
          const bool wrap = l + 1 == NL;                      // next: layer l + 1 of this tile, or layer 0 of this CTA's next tile
          const DgradLayer& N = prog.layers[wrap ? 0 : l + 1];
          const uint8_t* nrec = wrap ? mrec + (size_t)gridDim.x * act_stride : mrec;
          const bool have = !wrap || t + 1 < my_tiles;
#pragma unroll
          for (int kb = 0; kb < 4; ++kb)
            mb_next[kb] = (have && N.mask_slot >= 0 && kb < (N.n_out >> 6))
                              ? __ldg(reinterpret_cast<const uint32_t*>(nrec + act_mask_offset(args.act_slots, N.mask_slot + kb) + grp * 512 + row * 4))
                              : 0xFFFFFFFFu;
        }
        const bool rows_out = L.rows_cols > 0 && args.g_rows != nullptr;
        ptx::mbar_wait(&sm.d_full[buf], (d_phase >> buf) & 1);
        d_phase ^= 1u << buf;
        ptx::tc_fence_after();
        if (warp == DG_EPI_WARP0 && lane == 0) DTRACE(1024 + l * 16);
        const float ds = wsig != nullptr ? sm.dsig[t & 1][row] : 0.f;
        // one K block of the layer's output: accumulator -> (+ rank-1 sigma term) -> masked bf16 pairs -> the next backward
        // layer's A operand in tensor memory
        auto produce = [&](int kb, uint32_t (&pk)[16]) {
          uint32_t v[32];
          ptx::tmem_ld32(t_row + kb * 64, v);
          ptx::tmem_ld_wait(v);
          const int col0 = kb * 64 + grp * DG_COLS;
          if (rows_out && col0 < L.rows_cols) {
            // fp32 input gradient (d loss / d product rows).  A thread owns a row, so storing its 32 columns directly makes every
            // 128-bit store instruction of the warp touch 32 different lines (~1 400 cycles per block, tools/dgrad_rows_trace.py).
            // The warp's 32 x 32 block goes through 4 KB of shared memory instead (16-byte units XOR-swizzled by the row: both
            // phases are conflict-free) and leaves as row segments: 8 lanes write 128 contiguous bytes, 4 rows per instruction.
            // The staging area is h[2..3], unused here: a chain with a row output has a 128-wide top (checked by the host).
            float* stg = reinterpret_cast<float*>(&sm.h[2][0]) + (warp - DG_EPI_WARP0) * 1024;
#pragma unroll
            for (int q = 0; q < 8; ++q)
              *reinterpret_cast<float4*>(stg + lane * 32 + ((q ^ (lane & 7)) << 2)) =
                  make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]), __uint_as_float(v[4 * q + 2]), __uint_as_float(v[4 * q + 3]));
            __syncwarp();
            const long long row0 = tile * 128 + quarter * 32;
            const int u = lane & 7;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int rr = 4 * i + (lane >> 3);
              const float4 x = *reinterpret_cast<const float4*>(stg + rr * 32 + ((u ^ (rr & 7)) << 2));
              if (row0 + rr < total_rows && col0 + 4 * u + 4 <= L.rows_cols)
                *reinterpret_cast<float4*>(args.g_rows + (size_t)(row0 + rr) * args.g_row_pitch + col0 + 4 * u) = x;
            }
            __syncwarp();
          }
          if (wsig != nullptr) {
            const float4* w4 = reinterpret_cast<const float4*>(wsig + col0);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const float4 w = w4[q];
              v[4 * q + 0] = __float_as_uint(fmaf(ds, w.x, __uint_as_float(v[4 * q + 0])));
              v[4 * q + 1] = __float_as_uint(fmaf(ds, w.y, __uint_as_float(v[4 * q + 1])));
              v[4 * q + 2] = __float_as_uint(fmaf(ds, w.z, __uint_as_float(v[4 * q + 2])));
              v[4 * q + 3] = __float_as_uint(fmaf(ds, w.w, __uint_as_float(v[4 * q + 3])));
            }
          }
#pragma unroll
          for (int j = 0; j < 16; ++j) pk[j] = ptx::pack_bf16(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1])) & pair_mask(mb[kb], j);
          if (!last) {
            // K block kb of the next backward layer's A operand: packed pairs over the accumulator columns just drained
            ptx::tmem_st16(t_row + kb * 64, pk);
            ptx::tmem_st_wait();
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&sm.a_ready[kb]);
          }
          if (warp == DG_EPI_WARP0 && lane == 0) DTRACE(1024 + l * 16 + 1 + 2 * kb);
        };
        // dZ tile image for the weight-gradient kernel: staged in shared memory and shipped per lane quarter - the 32 rows of a quarter
        // are 4 KB of the image, written by the TWO warps of that quarter (same SM sub-partition), which meet at their own 64-thread
        // barrier; no CTA-wide epilogue barrier per block.  (Shipping two blocks at a time behind both hand-offs, as the forward does,
        // measured 4 % slower here: 2.19 vs 2.10 ms per iteration.)
        auto ship = [&](int kb, const uint32_t (&pk)[16]) {
          uint8_t* stg = sm.out[out_count % DG_OUT_BUFS];
#pragma unroll
          for (int u = 0; u < 4; ++u)
            *reinterpret_cast<uint4*>(stg + ptx::sw128_offset(row, grp * 4 + u)) = make_uint4(pk[4 * u], pk[4 * u + 1], pk[4 * u + 2], pk[4 * u + 3]);
          ptx::fence_proxy_async_smem();
          const bool q_leader = grp == 0 && lane == 0;
          if (q_leader) ptx::bulk_wait_read<DG_OUT_BUFS - 2>();   // the next block's buffer quarter is free once both warps pass the barrier
          asm volatile("bar.sync %0, 64;" ::"r"(4 + quarter) : "memory");
          if (q_leader) {
            ptx::bulk_s2g_hint(out + (size_t)kb * DG_KBLOCK + quarter * 4096, stg + quarter * 4096, 4096,
                               ptx::l2_policy_evict_first());     // read once, by wgrad, much later
            ptx::bulk_commit();
          }
          ++out_count;
          if (warp == DG_EPI_WARP0 && lane == 0) DTRACE(1024 + l * 16 + 2 + 2 * kb);
        };
#pragma unroll
        for (int kb = 0; kb < 4; ++kb) {
          if (kb >= out_blocks) break;
          uint32_t pk[16];
          produce(kb, pk);
          if (L.dz_slot >= 0) ship(kb, pk);
        }
        ptx::tc_fence_before();
      }
    }
    if (grp == 0 && lane == 0) ptx::bulk_wait_all();        // every quarter's shipping thread drains its own bulk groups
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem, 512);
  }
}

}  // namespace srf

using namespace srf;

SRF_API int srf_nerf_mlp_dgrad(const void* program, const void* weights_t, const float* side, const void* acts, int act_slots,
                               const float* sigma, const float* rgb, const float* g_sigma, const float* g_rgb, int64_t num_rows,
                               const int* count, void* dz, int dz_slots, float* g_rows, int g_row_pitch, void* stream) {
  if (num_rows == 0) return 0;
  SRF_REQUIRE(program && weights_t && side && acts && rgb && dz && (sigma || !g_sigma), "srf_nerf_mlp_dgrad", "null pointer");
  DgradProgram prog = *reinterpret_cast<const DgradProgram*>(program);
  SRF_REQUIRE(prog.num_layers >= 1 && prog.num_layers <= DG_MAX_LAYERS, "srf_nerf_mlp_dgrad", "bad layer count");
  SRF_REQUIRE(prog.top_width == 128 || prog.top_width == 256, "srf_nerf_mlp_dgrad", "top width must be 128 or 256");
  SRF_REQUIRE(prog.head_kind == 1 || prog.head_kind == 2, "srf_nerf_mlp_dgrad", "bad head kind");
  SRF_REQUIRE(prog.side_count <= DG_MAX_SIDE, "srf_nerf_mlp_dgrad", "side table too large");
  SRF_REQUIRE(prog.layers[0].num_kblocks * 64 == prog.top_width, "srf_nerf_mlp_dgrad", "first layer must consume dZ_top");
  for (int l = 0; l < prog.num_layers; ++l) {
    const DgradLayer& L = prog.layers[l];
    SRF_REQUIRE(L.n_out == 128 || L.n_out == 256, "srf_nerf_mlp_dgrad", "layer output width must be 128 or 256");
    SRF_REQUIRE(L.num_kblocks >= 1 && L.num_kblocks <= 4 && (l == 0 || L.num_kblocks * 64 == prog.layers[l - 1].n_out), "srf_nerf_mlp_dgrad",
                "K blocks must cover the previous layer's output");
    SRF_REQUIRE(L.dz_slot + L.n_out / 64 <= dz_slots && L.mask_slot + L.n_out / 64 <= act_slots, "srf_nerf_mlp_dgrad", "bad slot / mask index");
    SRF_REQUIRE(L.dz_slot >= 0 || l == prog.num_layers - 1, "srf_nerf_mlp_dgrad", "only the last layer may skip its dZ images");
    SRF_REQUIRE(L.rank1_offset < 0 || ((L.rank1_offset & 3) == 0 && L.n_out == 256), "srf_nerf_mlp_dgrad", "rank-1 offset must be a multiple of 4 (256-wide layers)");
    SRF_REQUIRE(L.rows_cols == 0 || (l == prog.num_layers - 1 && g_rows != nullptr && L.rows_cols % 4 == 0 && L.rows_cols <= L.n_out &&
                                     L.rows_cols <= g_row_pitch && g_row_pitch % 4 == 0 && prog.top_width == 128),
                "srf_nerf_mlp_dgrad", "bad fp32 row output (last layer only, multiples of 4, a 128-wide top: the rows are staged in the upper half of h)");
  }
  SRF_REQUIRE(prog.top_mask_slot >= 0 && prog.top_mask_slot + prog.top_width / 64 <= act_slots, "srf_nerf_mlp_dgrad", "bad top mask slot");
  SRF_REQUIRE(prog.head_slot >= 0 && prog.head_slot + 2 <= dz_slots && prog.top_slot >= 0 && prog.top_slot + prog.top_width / 64 <= dz_slots,
              "srf_nerf_mlp_dgrad", "bad head / top slot");
  DgradArgs a{};
  a.weights_t = reinterpret_cast<const uint8_t*>(weights_t); a.side = side; a.acts = reinterpret_cast<const uint8_t*>(acts);
  a.act_slots = act_slots; a.sigma = sigma; a.rgb = rgb;
  a.g_sigma = g_sigma; a.g_rgb = g_rgb; a.dz = reinterpret_cast<uint8_t*>(dz); a.dz_slots = dz_slots; a.total = num_rows; a.count = count;
  a.g_rows = g_rows; a.g_row_pitch = g_row_pitch;
  const size_t smem = sizeof(DgradSmem);
  static unsigned long long configured = 0ull;      // one bit per device
  if (first_use_on_this_device(configured)) {
    cudaError_t e = cudaFuncSetAttribute(nerf_mlp_dgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail("srf_nerf_mlp_dgrad", cudaGetErrorString(e));
  }
  const long long tiles = (num_rows + 127) / 128;
  const int grid = tiles < sm_count() ? (int)tiles : sm_count();
  nerf_mlp_dgrad_kernel<<<grid, DG_THREADS, smem, (cudaStream_t)stream>>>(prog, a);
  return check_launch("srf_nerf_mlp_dgrad");
}

SRF_API int srf_dgrad_program_bytes(void) { return (int)sizeof(DgradProgram); }

#if SRF_MLP_TRACE
extern "C" __attribute__((visibility("default"))) int srf_debug_dgrad_trace(long long* host_out) {
  cudaDeviceSynchronize();
  return cudaMemcpyFromSymbol(host_out, srf::g_dg_trace, sizeof(long long) * 2048) == cudaSuccess ? 0 : 1;
}
#endif
