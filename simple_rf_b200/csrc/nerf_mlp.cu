// Fused NeRF MLP forward on tcgen05 tensor cores: sample points -> positional encoding -> 8x256 trunk
// (+skip) -> sigma head -> feature -> view branch -> rgb, one persistent CTA per SM, 128 samples per tile.
//
// Replaces src/models/SimpleNeRF17.py:213-215 (pts = o + d z), :581-613 (PositionalEncoder), :419-484
// (run_network / batchify chunk loop) and :696-785 (MLP.forward and both heads), for the three shipped
// variants (main, points_augmentation with the degree-3 sigma input, views_augmentation with the 4-wide
// head); the variant is a small "layer program" passed by the host (srf_mlp_layer in the header).
//
// Data flow per tile (nothing between the ray data and sigma/rgb touches HBM, and the hidden activations never
// leave tensor memory):
//   encoding warps: points + encoding -> bf16 A-operand K-blocks E (points) / V (views) in shared memory, one tile ahead
//   producer warp : streams pre-packed bf16 weight images (128 output units x 64-wide K block, 128B-swizzled) through a
//                   ring with cp.async.bulk + mbarrier complete_tx (TMA bulk engine), in schedule order
//   2 MMA issuers : alternate steps of a host-flattened schedule; a step = four tcgen05.mma M=128, N=128, K=16 of one
//                   K block of one 128-column half; B from shared memory, A from shared memory (E, V) or from tensor
//                   memory (hidden activations); fp32 accumulators in TMEM (two 256-column buffers, alternating per layer)
//   epilogue warps: tcgen05.ld -> +bias, ReLU -> bf16 pairs -> tcgen05.st over the accumulator columns just drained = the
//                   next layer's A operand, signalled per 64-column K-block so the next layer's MMAs start while the rest
//                   of the epilogue is still running; the 1/3/4-wide heads are thread-local dot products.
#include <stdlib.h>
#include <type_traits>
#include "common.cuh"
#include "tcgen05.cuh"

namespace srf {

constexpr int MLP_MAX_LAYERS = 12;
constexpr int MLP_MAX_KBLOCKS = 6;

struct MlpLayer {              // mirrors srf_mlp_layer in include/simple_rf_b200.h
  int32_t num_kblocks;
  int32_t kblock_region[MLP_MAX_KBLOCKS];   // 0 = E, 1..4 = H0..H3, 5 = V
  int32_t kblock_ksteps[MLP_MAX_KBLOCKS];   // MMA K-steps (16 wide) issued for that block, 1..4
  int32_t n;                                // 256 or 128
  int32_t relu;                             // apply ReLU in the epilogue
  int32_t write_h;                          // store the activation as the next A operand
  int32_t head;                             // 0 none, 1 sigma (1 row), 2 sigma+rgb (4 rows), 3 rgb (3 rows)
  int32_t bias_offset;                      // float offset in the side table
  int32_t head_offset;                      // float offset of head weights [rows][n] followed by head bias [rows]
  int32_t save_slot;                        // training: first image slot of this layer's output in the saved tile, or -1
  int64_t weight_offset;                    // byte offset of the layer's packed K-blocks in the blob
};

struct MlpProgram {
  int32_t num_layers;
  int32_t points_degree;        // positional-encoding degree of the points (<= 10)
  int32_t views_degree;         // of the view directions (<= 4); < 0: no view branch
  int32_t side_count;           // floats in the side table
  MlpLayer layers[MLP_MAX_LAYERS];
  // 0: bf16 operands.  > 0: split-bf16 ("bf16x3") operands for the fp32 contract: every activation and weight is a pair
  // hi = bf16(x), lo = bf16(x - hi); a K block contributes A_hi W_hi + A_lo W_hi + A_hi W_lo (relative error ~2^-16 per
  // product instead of 2^-8); the lo image of the weight image at byte offset o of the blob sits at o + lo_offset
  int64_t lo_offset;
};

struct MlpArgs {
  const uint8_t* weights;       // packed bf16 blob
  const float* side;            // fp32 side table: biases, head weights, head biases
  const float* rays_o;          // [R,3]   origin used for the sample points (NDC origin when ndc)
  const float* rays_d;          // [R,3]
  const float* z;               // [R,S]
  const float* view_dirs;       // [R,3] or nullptr
  const float* noise;           // [R*S] or nullptr: added to raw sigma before the ReLU
  const uint4* rows;            // rows mode: [total, 128] bf16 inputs: columns 0..63 fill region 0, 64..127 region 5
  const int* count;             // rows mode: device-side row count (<= total), or nullptr
  int row_units;                // rows mode: 16-byte units per row (units 0..7 -> region 0, 8..15 -> region 5)
  float* sigma;                 // [R*S]
  float* rgb;                   // [R*S,3]
  // training: every A-operand tile (encodings, post-activation layer outputs) is also written to HBM as the same
  // 128x64 bf16 swizzled images the tensor cores read ([tile][slot][16 KB]); the backward kernels read them both as
  // GEMM operands and as ReLU masks (post-activation value != 0)
  uint8_t* save_acts;           // or nullptr
  int act_slots, e_slot, v_slot;
  long long total;              // R*S (rows mode: capacity of `rows`)
  int S;
};

// SRF_MLP_TRACE: CTA 0 records clock64() at pipeline events of its 3rd and 4th tile into a global buffer (tools/mlp_trace.py)
#ifndef SRF_MLP_TRACE
#define SRF_MLP_TRACE 0
#endif
#if SRF_MLP_TRACE
__device__ long long g_mlp_trace[4096];
#define TRACE(slot) do { if (blockIdx.x == 0 && (t == 2 || t == 3)) g_mlp_trace[(slot) + (t - 2) * 2048] = clock64(); } while (0)
#else
#define TRACE(slot) do { } while (0)
#endif
// SRF_MLP_SPLIT 1: a 256-wide layer is issued as two 128-column halves (all K blocks of half 0, then of half 1), so the
// epilogue of half 0 - and the hand-off of the next layer's first two K blocks - runs under the MMAs of half 1
#ifndef SRF_MLP_SPLIT
#define SRF_MLP_SPLIT 1
#endif
#ifndef SRF_MLP_CHUNK
#define SRF_MLP_CHUNK 2
#endif
#ifndef SRF_PAIR_CLUSTER_ACQUIRE
#define SRF_PAIR_CLUSTER_ACQUIRE 0
#endif
#ifndef SRF_MLP_PAIR_DEFAULT
#define SRF_MLP_PAIR_DEFAULT 0
#endif
constexpr int GROUPS = 2;                         // column groups per 64-wide block = epilogue warps per TMEM lane quarter
constexpr int COLS = 64 / GROUPS;                 // columns of a 64-wide block owned by one epilogue warp
constexpr int EPI_THREADS = 128 * GROUPS;
// Warp roles: 0 weight producer (one lane), 1 first MMA issuer, 2..9 epilogue, 10..13 encoding, 14 second MMA issuer.
// Every SM sub-partition (warp id % 4) hosts two epilogue warps and one encoding warp; TMEM lane quarters follow
// warp id % 4, which any 8 consecutive warps cover twice.  Two other numberings were measured on hardware and lost 10 %
// (encoding warps below the epilogue warps; both issuers first): keep this one.
constexpr int PRODUCER_WARP = 0;
constexpr int ISSUER1_WARP = 1;
constexpr int EPI_WARP0 = 2;                      // 4 * GROUPS epilogue warps
constexpr int ENC_WARP0 = EPI_WARP0 + 4 * GROUPS; // 4 encoding warps (one row per thread) run one tile ahead
constexpr int ISSUER2_WARP = ENC_WARP0 + 4;
constexpr int MLP_THREADS = 32 * (ISSUER2_WARP + 1);
constexpr int KBLOCK_BYTES = 128 * 128;           // 128 rows x 64 bf16
constexpr int IMAGE_BYTES = 128 * 128;            // packed weight image: 128 output units x one 64-wide K block
#if SRF_MLP_SPLIT
constexpr int STAGE_BYTES = IMAGE_BYTES;          // one 128-row half of a K block per ring stage
constexpr int NUM_STAGES = 10;
#else
constexpr int STAGE_BYTES = 2 * IMAGE_BYTES;      // both 128-row halves of a K block per ring stage
constexpr int NUM_STAGES = 5;
#endif
// training (activation tiles are saved): the ring shrinks to SAVE_RING_BYTES and the rest of `w` holds SAVE_BUFS
// 16 KB staging images, written by the epilogue warps and drained to HBM by bulk async copies
constexpr int SAVE_BUFS = 4;
constexpr int SAVE_STAGES = NUM_STAGES - SAVE_BUFS * IMAGE_BYTES / STAGE_BYTES;
// rows mode (TensoRF colour MLP: <= 4 MMA steps per tile): the weight images stay RESIDENT in the first ROWS_RING stages (loaded once per
// CTA, the ring protocol keeps cycling without copies), and the input rows of the next tiles are staged by bulk async copies into
// ROW_BUF_BYTES buffers behind them: three tiles ahead at test time, one when the saved-tile staging images take stages SAVE_STAGES..
constexpr int ROWS_RING = 4;
constexpr int ROW_BUFS = 3;
constexpr int ROW_BUF_BYTES = 128 * 256;          // 128 rows x (at most) 16 units of 16 bytes
constexpr int A_REGIONS = 2;                      // E and V only; the hidden activations live in tensor memory
constexpr int V_REGION = 1;
// split-bf16 mode: the lo parts of E and V take the last two ring stages (the ring shrinks to NUM_STAGES - 2)
constexpr int LO_STAGES = 2 * KBLOCK_BYTES / STAGE_BYTES;
constexpr int E_LO_STAGE = NUM_STAGES - LO_STAGES, V_LO_STAGE = E_LO_STAGE + KBLOCK_BYTES / STAGE_BYTES;
constexpr int MAX_SIDE = 4096;                    // floats
constexpr int MAX_STEPS = 224;                    // bf16: <= 74 steps for the shipped variants; split-bf16 issues three per K block

// One step of the MMA issuers = the (up to) four K=16 MMAs of one 64-wide K block [of one 128-column half].  The host
// flattens the layer program into this schedule and passes it as a kernel parameter (constant bank); the producer streams
// the weight images in the same order.
struct alignas(16) MmaStep {
  uint32_t a_off;         // region 0 / 5: (byte offset of the A region inside MlpSmem::a) >> 4;  H block r: TMEM column 64 (r - 1)
  uint32_t idesc;         // instruction descriptor (M = 128, N = columns of this step)
  // bits 0-2: 16-wide K steps to issue; 3: first step of the accumulator half (overwrite); 4: last step of the half
  // (commit d_full); 5 / 6: last reader of region 0 / 5 (commit e_free / v_free); 7: column half (0 / 1);
  // 8-11: 1 + region whose a_ready barrier must be acquired first (0: none); 12: layer index & 1 (selects the accumulator
  // buffer); 13: A operand in tensor memory; 14: index (& 1) of this acquisition among the tile's acquisitions of that
  // region; 15: their count per tile (& 1); 16-17: weight images of this step (1 or 2); 18-20: K blocks of the layer
  uint32_t meta;
  uint32_t w_off;         // (byte offset of the step's first weight image in the blob) >> 4
};
struct MmaSchedule {
  int32_t num_steps;
  int32_t pad_[3];
  MmaStep steps[MAX_STEPS];
};

struct alignas(1024) MlpSmem {
  uint8_t a[A_REGIONS][KBLOCK_BYTES];     // E, V
  uint8_t w[NUM_STAGES][STAGE_BYTES];
  float side[MAX_SIDE];
  float part[2][GROUPS][128][4];          // head partial sums of each column group
  uint64_t w_full[NUM_STAGES], w_empty[NUM_STAGES];
  uint64_t rows_full[ROW_BUFS];           // rows mode: the bulk copy of a tile's input rows has landed in its staging buffer
  uint64_t pair_full[NUM_STAGES];         // CTA pair: the PEER CTA's half of the stage's weight image has landed (leader only)
  uint64_t a_ready[6];                    // per A region (0 E, 1..4 H blocks, 5 V): written and visible to the tensor core
  uint64_t d_full[4];                     // [accumulator buffer][column half] complete
  uint64_t e_free, v_free;                // every MMA reading region 0 / 5 of the current tile has retired
  uint32_t tmem_base;
};

__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory"); }

// sin/cos of 2^k x for k = k0 .. k0+count-1 by angle doubling from one accurate sincosf
template <typename F>
__device__ __forceinline__ void encode_octaves(float x, int k0, int count, F&& emit) {
  float s, c;
  sincosf(ldexpf(x, k0), &s, &c);
  for (int k = 0; k < count; ++k) {
    emit(k0 + k, s, c);
    const float s2 = 2.f * s * c;
    c = 1.f - 2.f * s * s;
    s = s2;
  }
}

// split-bf16 mode: one accurate sincosf per octave (2^k x is exact in fp32; the doubling recurrence above loses ~1 bit per octave,
// invisible under bf16 rounding but not against the 1e-3 fp32 contract)
template <typename F>
__device__ __forceinline__ void encode_octaves_exact(float x, int count, F&& emit) {
  for (int k = 0; k < count; ++k) {
    float s, c;
    sincosf(ldexpf(x, k), &s, &c);
    emit(k, s, c);
  }
}

__device__ __forceinline__ void store_bf16(uint8_t* block, int row, int col, float v) {
  const uint32_t off = ptx::sw128_offset(row, col >> 3) + ((col & 7) << 1);
  *reinterpret_cast<__nv_bfloat16*>(block + off) = __float2bfloat16_rn(v);
}

// split-bf16: hi into `block`, the bf16 of the remainder into `block_lo`
__device__ __forceinline__ void store_bf16_split(uint8_t* block, uint8_t* block_lo, int row, int col, float v) {
  const uint32_t off = ptx::sw128_offset(row, col >> 3) + ((col & 7) << 1);
  const __nv_bfloat16 hi = __float2bfloat16_rn(v);
  *reinterpret_cast<__nv_bfloat16*>(block + off) = hi;
  *reinterpret_cast<__nv_bfloat16*>(block_lo + off) = __float2bfloat16_rn(v - __bfloat162float(hi));
}

template <int N>
__device__ __forceinline__ void tmem_load(uint32_t taddr, uint32_t (&v)[N]);
template <>
__device__ __forceinline__ void tmem_load<32>(uint32_t taddr, uint32_t (&v)[32]) { ptx::tmem_ld32(taddr, v); }
template <>
__device__ __forceinline__ void tmem_load<16>(uint32_t taddr, uint32_t (&v)[16]) { ptx::tmem_ld16(taddr, v); }

// SPLIT: split-bf16 operands (prog.lo_offset != 0); a separate instantiation keeps the bf16 path free of it.
// PAIR: launched as clusters of two CTAs (the two SMs of a TPC) that share every MMA (tcgen05 cta_group::2, M = 256): each CTA keeps
// its own tile, encoder, epilogue and tensor memory, but holds only HALF of every weight image (64 of the step's 128 output units) -
// the shared-memory traffic of the weight stream, which is what bounds this kernel, halves in both directions, and so does the L2
// stream.  The leader CTA's issuers drive the MMAs of both tiles; the peer's arrivals reach the leader's barriers through
// shared::cluster addresses, completions come back as multicast commits.
template <bool SPLIT, bool PAIR>
__global__ void __launch_bounds__(MLP_THREADS, 1) nerf_mlp_fwd_kernel(const __grid_constant__ MlpProgram prog,
                                                                      const __grid_constant__ MmaSchedule sched, const MlpArgs args) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  MlpSmem& sm = *reinterpret_cast<MlpSmem*>(smem_raw);
  // warp index through a lane-0 shuffle: tells the compiler it is warp-uniform (role dispatch and the issuers' step
  // data can then live in uniform registers)
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
  if ((ptx::smem_u32(smem_raw) & 1023u) != 0) __trap();   // 128B-swizzle atoms need a 1024-byte aligned base

  const uint32_t cta_rank = PAIR ? ptx::cluster_ctarank() : 0u;
  const bool leader_cta = cta_rank == 0u;
  constexpr uint32_t CTAS = PAIR ? 2u : 1u;
  for (int i = threadIdx.x; i < prog.side_count; i += MLP_THREADS) sm.side[i] = args.side[i];
  if (warp == PRODUCER_WARP && lane == 0) {
    for (int s = 0; s < NUM_STAGES; ++s) { ptx::mbar_init(&sm.w_full[s], 1); ptx::mbar_init(&sm.w_empty[s], 1); ptx::mbar_init(&sm.pair_full[s], 1); }
    ptx::mbar_init(&sm.a_ready[0], 4 * CTAS);                    // the leader's barriers collect both CTAs' arrivals
    ptx::mbar_init(&sm.a_ready[5], 4 * CTAS);
    for (int r = 1; r < 5; ++r) ptx::mbar_init(&sm.a_ready[r], 4 * GROUPS * CTAS);
    for (int b = 0; b < 4; ++b) ptx::mbar_init(&sm.d_full[b], 1);
    ptx::mbar_init(&sm.e_free, 1);
    ptx::mbar_init(&sm.v_free, 1);
    for (int b = 0; b < ROW_BUFS; ++b) ptx::mbar_init(&sm.rows_full[b], 1);
    ptx::fence_barrier_init();
  }
  if constexpr (PAIR) {
    __syncthreads();
    ptx::cluster_sync_all();                                     // both CTAs' barriers exist before anything arrives remotely
    if (warp == ISSUER1_WARP) ptx::tmem_alloc_pair(&sm.tmem_base, 512);
  } else {
    if (warp == ISSUER1_WARP) ptx::tmem_alloc(&sm.tmem_base, 512);
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = sm.tmem_base;
  long long total = args.total;
  if (args.count != nullptr) { const long long c = *args.count; total = c < total ? c : total; }
  const int num_tiles = (int)((total + 127) / 128);
  // persistent loop: CTA b takes tiles b, b + grid, ...; a CTA pair takes the tile pairs (2j, 2j + 1) the same way (a missing second tile
  // of the last pair is processed as an all-invalid tile: both CTAs must run the same steps)
  const int units = PAIR ? (num_tiles + 1) / 2 : num_tiles;
  const int unit0 = PAIR ? (int)blockIdx.x / 2 : (int)blockIdx.x, unit_stride = PAIR ? (int)gridDim.x / 2 : (int)gridDim.x;
  const int my_tiles = units > unit0 ? (units - unit0 + unit_stride - 1) / unit_stride : 0;
  auto tile_of = [&](int t) { return (long long)(unit0 + t * unit_stride) * (long long)CTAS + cta_rank; };
  constexpr bool split = SPLIT;                    // (launcher: never together with save_acts / rows)
  const bool rows_mode = args.rows != nullptr;
  // rows mode: one ring stage per step of the tile, so stage s always holds step s's image and nothing is ever re-loaded
  const uint32_t num_stages = rows_mode ? (uint32_t)sched.num_steps
                                        : (args.save_acts != nullptr ? SAVE_STAGES : (split ? (uint32_t)E_LO_STAGE : NUM_STAGES));
  const bool has_views = prog.views_degree >= 0 || (args.rows != nullptr && prog.views_degree == -2);   // -2: rows mode, 2 blocks

  if (warp == PRODUCER_WARP) {
    // ------------------------------------------------------------ weight producer: one ring stage per schedule step
    if (lane == 0) {
      const int num_steps = sched.num_steps;
      uint32_t stage = 0, ph = 0;
      const uint64_t keep = ptx::l2_policy_evict_last();      // weight images are re-read per tile by every CTA: keep them in L2
      for (int t = 0; t < my_tiles; ++t) {
        for (int s = 0; s < num_steps; ++s) {
          const MmaStep st = sched.steps[s];
          const uint32_t images = (st.meta >> 16) & 3u;
          ptx::mbar_wait(&sm.w_empty[stage], ph ^ 1);
          if constexpr (PAIR) {
#if SRF_MLP_SPLIT
            // N = 128 steps: this CTA's half of the image, output units 64 * rank .. 64 * rank + 63 = eight 1 KB swizzle atoms, contiguous
            ptx::mbar_arrive_expect_tx(&sm.w_full[stage], IMAGE_BYTES / 2);
            ptx::bulk_g2s(sm.w[stage], args.weights + ((size_t)st.w_off << 4) + (size_t)cta_rank * (IMAGE_BYTES / 2), IMAGE_BYTES / 2,
                          &sm.w_full[stage]);
#else
            // N = 256 steps: the leader holds the image of output units 0..127, the peer the image of units 128..255
            const uint32_t half_stride = ((st.meta >> 18) & 7u) * IMAGE_BYTES;
            const uint32_t mine = images > 1 ? cta_rank : 0u;
            ptx::mbar_arrive_expect_tx(&sm.w_full[stage], images > 1 ? IMAGE_BYTES : IMAGE_BYTES / 2);
            if (images > 1) ptx::bulk_g2s(sm.w[stage], args.weights + ((size_t)st.w_off << 4) + (size_t)mine * half_stride, IMAGE_BYTES, &sm.w_full[stage]);
            else ptx::bulk_g2s(sm.w[stage], args.weights + ((size_t)st.w_off << 4) + (size_t)cta_rank * (IMAGE_BYTES / 2), IMAGE_BYTES / 2, &sm.w_full[stage]);
#endif
          } else if (rows_mode && t > 0) {
            ptx::mbar_arrive(&sm.w_full[stage]);        // resident image: the stage is "refilled" without a copy
          } else {
            ptx::mbar_arrive_expect_tx(&sm.w_full[stage], images * IMAGE_BYTES);
            const uint32_t half_stride = ((st.meta >> 18) & 7u) * IMAGE_BYTES;     // blob order: [layer][128-row half][K block]
            for (uint32_t i = 0; i < images; ++i)
              ptx::bulk_g2s_hint(sm.w[stage] + i * IMAGE_BYTES, args.weights + ((size_t)st.w_off << 4) + (size_t)i * half_stride, IMAGE_BYTES,
                                 &sm.w_full[stage], keep);
          }
          if (++stage == num_stages) { stage = 0; ph ^= 1; }
        }
      }
    }
  } else if (PAIR && !leader_cta && (warp == ISSUER1_WARP || warp == ISSUER2_WARP)) {
    // ------------------------------------------------------------ peer CTA: no MMA issue.  One warp tells the leader, stage by stage,
    // that this CTA's half of the weight image has landed (the leader's MMA reads it from here)
    if (warp == ISSUER1_WARP) {
      const int num_steps = sched.num_steps;
      const uint32_t remote = ptx::map_to_cta(ptx::smem_u32(&sm.pair_full[0]), 0u);
      uint32_t stage = 0, ph = 0;
      for (int t = 0; t < my_tiles; ++t) {
        for (int s = 0; s < num_steps; ++s) {
          ptx::mbar_wait(&sm.w_full[stage], ph);
          if (lane == 0) ptx::mbar_arrive_cluster_relaxed(remote + stage * 8u);     // the bytes were landed by the copy engine, not by this thread
          if (++stage == num_stages) { stage = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == ISSUER1_WARP || warp == ISSUER2_WARP) {
    // ------------------------------------------------------------ MMA issuers: two warps take alternate steps; each prepares
    // its next step (schedule entry, barrier acquisition, descriptors) while the other one issues, and a named-barrier token
    // keeps the MMAs in schedule order.  The whole warp runs the loop (uniform control flow: descriptors stay in uniform
    // registers), one elected lane issues tcgen05.mma / tcgen05.commit.
    const int num_steps = sched.num_steps;
    const uint32_t me = warp == ISSUER1_WARP ? 0u : 1u;
    const uint32_t desc_hi = (uint32_t)(1024 >> 4) | (1u << 14) | (2u << 29);   // SBO | version 1 | SWIZZLE_128B
    const uint32_t w_lo = (ptx::smem_u32(sm.w[0]) >> 4) & 0x3FFF;
    const uint32_t a_lo = (ptx::smem_u32(sm.a[0]) >> 4) & 0x3FFF;
    const uint32_t bar_full = ptx::smem_u32(&sm.w_full[0]), bar_empty = ptx::smem_u32(&sm.w_empty[0]);
    const uint32_t bar_pair = ptx::smem_u32(&sm.pair_full[0]);
    const uint32_t bar_a = ptx::smem_u32(&sm.a_ready[0]), bar_d = ptx::smem_u32(&sm.d_full[0]);
    const uint32_t bar_e = ptx::smem_u32(&sm.e_free), bar_v = ptx::smem_u32(&sm.v_free);
    const uint4* steps = reinterpret_cast<const uint4*>(sched.steps);
    const uint32_t odd_layers = (uint32_t)prog.num_layers & 1u;
    // global step g = t * num_steps + s; this warp owns g = me, me + 2, ...; ring stage g % num_stages, phase (g / num_stages) & 1
    uint32_t stage = me % num_stages, ph = (me / num_stages) & 1;
    int s = (int)me, t = 0;
    while (s >= num_steps) { s -= num_steps; ++t; }
    if (me == 1) asm volatile("bar.arrive 2, 64;" ::: "memory");       // the first warp may issue step 0
    while (t < my_tiles) {
      const uint4 st = steps[s];           // x: A offset, y: instruction descriptor, z: packed step data
      const uint32_t meta = st.z;
      const uint32_t wr = (meta >> 8) & 15u;
      const uint32_t tp = (uint32_t)t & 1u;
      if constexpr (PAIR) {
#if SRF_PAIR_CLUSTER_ACQUIRE
        if (wr) ptx::mbar_wait_addr_cluster(bar_a + (wr - 1) * 8, ((meta >> 14) ^ ((meta >> 15) & tp)) & 1u);
        ptx::mbar_wait_addr_cluster(bar_pair + stage * 8, ph);
#else
        // CTA-scope waits: nothing the peer wrote is read by THIS thread - the consumers are the tensor cores (async proxy), and the
        // producers ordered their writes with tcgen05.wait::st / fence.proxy.async before arriving
        if (wr) ptx::mbar_wait_addr(bar_a + (wr - 1) * 8, ((meta >> 14) ^ ((meta >> 15) & tp)) & 1u);
        ptx::mbar_wait_addr(bar_pair + stage * 8, ph);
#endif
      } else {
        if (wr) ptx::mbar_wait_addr(bar_a + (wr - 1) * 8, ((meta >> 14) ^ ((meta >> 15) & tp)) & 1u);
      }
      if (lane == 0) TRACE(16 + s * 4 + 0);
      ptx::mbar_wait_addr(bar_full + stage * 8, ph);
      if (lane == 0) TRACE(16 + s * 4 + 1);
      const uint32_t buf = ((meta >> 12) ^ (tp & odd_layers)) & 1;
      const uint32_t half = (meta >> 7) & 1u;
      const uint32_t issue = ptx::elect_one();
      const uint32_t d_addr = tmem + buf * 256 + half * 128;
      const uint32_t b_lo = w_lo + stage * (STAGE_BYTES >> 4);
      const uint32_t acc = (meta >> 3) & 1 ? 0u : 1u, ksteps = meta & 7u;
      const uint32_t bx = meta & 0x10u ? bar_d + (buf * 2 + half) * 8 : 0u, by = meta & 0x20u ? bar_e : 0u, bz = meta & 0x40u ? bar_v : 0u;
      // token (inside the statement): every MMA of the previous step has been issued by the other warp
      if (meta & 0x2000u) {          // A = H block of the previous layer, packed in the other accumulator buffer's columns
        const uint32_t a_t = tmem + (buf ^ 1u) * 256 + st.x;
        if constexpr (PAIR) {
          if (me == 0) ptx::umma4_step_pair<2, 3, true>(issue, d_addr, a_t, b_lo, desc_hi, st.y, acc, ksteps, bar_empty + stage * 8, bx, by, bz);
          else ptx::umma4_step_pair<3, 2, true>(issue, d_addr, a_t, b_lo, desc_hi, st.y, acc, ksteps, bar_empty + stage * 8, bx, by, bz);
        } else {
          if (me == 0) ptx::umma4_step<2, 3, true>(issue, d_addr, a_t, b_lo, desc_hi, st.y, acc, ksteps, bar_empty + stage * 8, bx, by, bz);
          else ptx::umma4_step<3, 2, true>(issue, d_addr, a_t, b_lo, desc_hi, st.y, acc, ksteps, bar_empty + stage * 8, bx, by, bz);
        }
      } else {
        if constexpr (PAIR) {
          if (me == 0) ptx::umma4_step_pair<2, 3, false>(issue, d_addr, a_lo + st.x, b_lo, desc_hi, st.y, acc, ksteps, bar_empty + stage * 8, bx, by, bz);
          else ptx::umma4_step_pair<3, 2, false>(issue, d_addr, a_lo + st.x, b_lo, desc_hi, st.y, acc, ksteps, bar_empty + stage * 8, bx, by, bz);
        } else {
          if (me == 0) ptx::umma4_step<2, 3, false>(issue, d_addr, a_lo + st.x, b_lo, desc_hi, st.y, acc, ksteps, bar_empty + stage * 8, bx, by, bz);
          else ptx::umma4_step<3, 2, false>(issue, d_addr, a_lo + st.x, b_lo, desc_hi, st.y, acc, ksteps, bar_empty + stage * 8, bx, by, bz);
        }
      }
      if (lane == 0) TRACE(16 + s * 4 + 2);
      stage += 2;
      if (stage >= num_stages) { stage -= num_stages; ph ^= 1; }
      s += 2;
      while (s >= num_steps) { s -= num_steps; ++t; }
    }
  } else if (warp >= ENC_WARP0 && warp < ENC_WARP0 + 4) {
    // ------------------------------------------------------------ encoding warps: region 0 (E) and 5 (V), one tile ahead
    const int row = (warp - ENC_WARP0) * 32 + lane;
    // rows mode: tile t's rows arrive in staging buffer t % row_bufs by ONE bulk copy issued row_bufs tiles earlier (the rows of a tile
    // are contiguous); a thread fetching its own row from global memory one tile ahead left the ~2 us load latency exposed on every
    // tile (4500 cycles per tile against ~1000 of MMA work)
    const int row_bufs = args.save_acts != nullptr ? 1 : ROW_BUFS;
    uint8_t* row_stage = &sm.w[0][0] + (size_t)ROWS_RING * STAGE_BYTES;
    auto fetch_rows = [&](int t) {                    // one thread: bulk copy of tile t's rows (clamped to the rows array)
      const long long row0 = tile_of(t) * 128;
      const long long avail = args.total - row0;
      const uint32_t bytes = (uint32_t)(avail < 128 ? (avail > 0 ? avail : 0) : 128) * (uint32_t)args.row_units * 16u;
      const int b = t % row_bufs;
      if (bytes == 0) { ptx::mbar_arrive(&sm.rows_full[b]); return; }
      ptx::mbar_arrive_expect_tx(&sm.rows_full[b], bytes);
      ptx::bulk_g2s(row_stage + (size_t)b * ROW_BUF_BYTES, args.rows + row0 * args.row_units, bytes, &sm.rows_full[b]);
    };
    if (rows_mode && warp == ENC_WARP0 && lane == 0)
      for (int t = 0; t < row_bufs && t < my_tiles; ++t) fetch_rows(t);
    for (int t = 0; t < my_tiles; ++t) {
      const long long tile = tile_of(t);
      const long long m = tile * 128 + row;
      const bool valid = m < total;
      float p[3] = {0.f, 0.f, 0.f}, vd[3] = {0.f, 0.f, 1.f};
      uint4 x[16];
      if (rows_mode) {
        const int b = t % row_bufs;
        ptx::mbar_wait(&sm.rows_full[b], (uint32_t)(t / row_bufs) & 1u);
        const uint4* mine = reinterpret_cast<const uint4*>(row_stage + (size_t)b * ROW_BUF_BYTES) + row * args.row_units;
#pragma unroll
        for (int q = 0; q < 16; ++q) x[q] = (valid && q < args.row_units) ? mine[q] : make_uint4(0u, 0u, 0u, 0u);
        asm volatile("bar.sync 4, 128;" ::: "memory");           // every encoder thread holds its row: the buffer may be refilled
        if (warp == ENC_WARP0 && lane == 0 && t + row_bufs < my_tiles) fetch_rows(t + row_bufs);
      } else if (valid) {
        const long long r = m / args.S;
        const float zz = args.z[m];
#pragma unroll
        for (int c = 0; c < 3; ++c) p[c] = __fadd_rn(args.rays_o[r * 3 + c], __fmul_rn(args.rays_d[r * 3 + c], zz));
        if (args.view_dirs != nullptr) {
#pragma unroll
          for (int c = 0; c < 3; ++c) vd[c] = args.view_dirs[r * 3 + c];
        }
      }
      // region 0 of the previous tile must have been read by its last MMA
      ptx::mbar_wait(&sm.e_free, (t & 1) ^ 1);
      if (warp == ENC_WARP0 && lane == 0) TRACE(0);
      uint8_t* E = sm.a[0];
      if (args.rows != nullptr) {
#pragma unroll
        for (int u = 0; u < 8; ++u) *reinterpret_cast<uint4*>(E + ptx::sw128_offset(row, u)) = x[u];
      } else {
        const int deg = prog.points_degree;
        if (split) {
          uint8_t* E_lo = sm.w[E_LO_STAGE];
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            store_bf16_split(E, E_lo, row, c, p[c]);
            if (deg > 0)
              encode_octaves_exact(p[c], deg, [&](int k, float s, float co) {
                store_bf16_split(E, E_lo, row, 3 + 6 * k + c, s);
                store_bf16_split(E, E_lo, row, 6 + 6 * k + c, co);
              });
          }
          for (int c = 3 + 6 * deg; c < 64; ++c) store_bf16_split(E, E_lo, row, c, 0.f);
        } else {
#pragma unroll
          for (int c = 0; c < 3; ++c) {
            store_bf16(E, row, c, p[c]);
            if (deg > 0)
              encode_octaves(p[c], 0, deg, [&](int k, float s, float co) {
                store_bf16(E, row, 3 + 6 * k + c, s);
                store_bf16(E, row, 6 + 6 * k + c, co);
              });
          }
          for (int c = 3 + 6 * deg; c < 64; ++c) store_bf16(E, row, c, 0.f);
        }
      }
      if (args.save_acts != nullptr) {                 // each thread re-reads the row it just wrote
        uint8_t* dst = args.save_acts + ((size_t)tile * act_tile_images(args.act_slots) + args.e_slot) * KBLOCK_BYTES;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const uint32_t off = ptx::sw128_offset(row, u);
          *reinterpret_cast<uint4*>(dst + off) = *reinterpret_cast<const uint4*>(E + off);
        }
      }
      ptx::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        // (pair) the peer signals the LEADER's barrier.  A relaxed arrive: what it announces was written to THIS SM's shared memory and
        // is read by THIS SM's tensor core; fence.proxy.async + __syncwarp above ordered it.  (release.cluster costs ~1000 cycles here.)
        if (PAIR && !leader_cta) ptx::mbar_arrive_cluster_relaxed(ptx::map_to_cta(ptx::smem_u32(&sm.a_ready[0]), 0u));
        else ptx::mbar_arrive(&sm.a_ready[0]);
      }
      if (warp == ENC_WARP0 && lane == 0) TRACE(1);
      if (has_views) {
        ptx::mbar_wait(&sm.v_free, (t & 1) ^ 1);
        uint8_t* V = sm.a[V_REGION];
        const int vdeg = prog.views_degree;
        if (args.rows != nullptr) {
#pragma unroll
          for (int u = 0; u < 8; ++u) *reinterpret_cast<uint4*>(V + ptx::sw128_offset(row, u)) = x[8 + u];
        } else {
          if (split) {
            uint8_t* V_lo = sm.w[V_LO_STAGE];
#pragma unroll
            for (int c = 0; c < 3; ++c) {
              store_bf16_split(V, V_lo, row, c, vd[c]);
              if (vdeg > 0)
                encode_octaves_exact(vd[c], vdeg, [&](int k, float s, float co) {
                  store_bf16_split(V, V_lo, row, 3 + 6 * k + c, s);
                  store_bf16_split(V, V_lo, row, 6 + 6 * k + c, co);
                });
            }
            for (int c = 3 + 6 * vdeg; c < 32; ++c) store_bf16_split(V, V_lo, row, c, 0.f);
          } else {
#pragma unroll
            for (int c = 0; c < 3; ++c) {
              store_bf16(V, row, c, vd[c]);
              if (vdeg > 0)
                encode_octaves(vd[c], 0, vdeg, [&](int k, float s, float co) {
                  store_bf16(V, row, 3 + 6 * k + c, s);
                  store_bf16(V, row, 6 + 6 * k + c, co);
                });
            }
            for (int c = 3 + 6 * vdeg; c < 32; ++c) store_bf16(V, row, c, 0.f);
          }
        }
        if (args.save_acts != nullptr) {
          uint8_t* dst = args.save_acts + ((size_t)tile * act_tile_images(args.act_slots) + args.v_slot) * KBLOCK_BYTES;
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const uint32_t off = ptx::sw128_offset(row, u);
            *reinterpret_cast<uint4*>(dst + off) = *reinterpret_cast<const uint4*>(V + off);
          }
        }
        ptx::fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) {
          if (PAIR && !leader_cta) ptx::mbar_arrive_cluster_relaxed(ptx::map_to_cta(ptx::smem_u32(&sm.a_ready[5]), 0u));
          else ptx::mbar_arrive(&sm.a_ready[5]);
        }
      }
    }
  } else {
    // ------------------------------------------------------------ epilogue warps
    const int quarter = warp & 3;        // TMEM lane quarter this warp may read: fixed by hardware to warp_id % 4
    const int grp = (warp - EPI_WARP0) >> 2;   // column group of every 64-wide block this warp owns
    const int row = quarter * 32 + lane;
    uint32_t layer_count = 0;
    const uint32_t side_addr = ptx::smem_u32(sm.side), d_full_addr = ptx::smem_u32(&sm.d_full[0]);
    const bool remote_arrive = PAIR && !leader_cta;              // the peer's epilogue signals the LEADER's barriers
    const uint32_t a_ready_addr = remote_arrive ? ptx::map_to_cta(ptx::smem_u32(&sm.a_ready[0]), 0u) : ptx::smem_u32(&sm.a_ready[0]);
    uint32_t d_phase = 0;                // bit b: parity to wait for on d_full[b]
    uint32_t save_count = 0;             // staged activation images so far (selects the staging buffer)
    for (int t = 0; t < my_tiles; ++t) {
      const long long tile = tile_of(t);
      const long long m = tile * 128 + row;
      const bool valid = m < total;
      for (int l = 0; l < prog.num_layers; ++l, ++layer_count) {
        const MlpLayer& L = prog.layers[l];
        const uint32_t buf = layer_count & 1;
        const int n = L.n, relu = L.relu, write_h = L.write_h, head = L.head;
        const uint32_t bias_addr = side_addr + (uint32_t)(L.bias_offset + grp * COLS) * 4u;   // this warp's 32 columns of block 0
        const int head_rows = head == 1 ? 1 : (head == 2 ? 4 : (head == 3 ? 3 : 0));
        const float* hw = sm.side + L.head_offset;
        float hacc[4] = {0.f, 0.f, 0.f, 0.f};
        const uint32_t t_row = tmem + ((uint32_t)(quarter * 32) << 16) + buf * 256 + grp * COLS;
        const int save_slot = args.save_acts != nullptr ? L.save_slot : -1;
        uint8_t* tile_base = args.save_acts != nullptr ? args.save_acts + (size_t)tile * act_tile_images(args.act_slots) * KBLOCK_BYTES : nullptr;
        uint8_t* save_base = save_slot >= 0 ? tile_base + (size_t)save_slot * KBLOCK_BYTES : nullptr;

        // bias + activation (+ head partial dot products) on one COLS-column slice, packed to bf16 pairs
        uint32_t mask_bits = 0;
        uint32_t pk_lo[COLS / 2];                                  // split-bf16: bf16 of the remainders x - float(bf16(x))
        auto split_pairs = [&](const float (&x)[COLS], uint32_t (&pk)[COLS / 2]) {
#pragma unroll
          for (int j = 0; j < COLS / 2; ++j) {
            const uint32_t h = ptx::pack_bf16(x[2 * j], x[2 * j + 1]);
            pk[j] = h;
            pk_lo[j] = ptx::pack_bf16(x[2 * j] - __uint_as_float(h << 16), x[2 * j + 1] - __uint_as_float(h & 0xffff0000u));
          }
        };
        auto compute = [&](uint32_t (&v)[COLS], uint32_t (&pk)[COLS / 2], int kb) {
          const int col0 = kb * 64 + grp * COLS;
          const uint32_t b4 = bias_addr + (uint32_t)kb * 256u;
          float x[COLS];
#pragma unroll
          for (int q = 0; q < COLS / 4; ++q) {                     // + bias as packed fp32 pairs (FADD2)
            const float4 b = ptx::lds4(b4 + q * 16);
            x[4 * q + 0] = __uint_as_float(v[4 * q + 0]); x[4 * q + 1] = __uint_as_float(v[4 * q + 1]);
            x[4 * q + 2] = __uint_as_float(v[4 * q + 2]); x[4 * q + 3] = __uint_as_float(v[4 * q + 3]);
            ptx::fadd2(x[4 * q + 0], x[4 * q + 1], b.x, b.y);
            ptx::fadd2(x[4 * q + 2], x[4 * q + 3], b.z, b.w);
          }
          if (save_base != nullptr) {
            // sign pattern of the pre-activation values = the ReLU mask the data-gradient chain applies (bit i: x[i] > 0):
            // top bit of -x (as integers) shifted in, last element first
            uint32_t bits = 0;
#pragma unroll
            for (int i = COLS - 1; i >= 0; --i) bits = __funnelshift_l(0u - __float_as_uint(x[i]), bits, 1);
            mask_bits = bits;
          }
          if (head_rows == 0 && split) {
            if (relu) {
#pragma unroll
              for (int j = 0; j < COLS; ++j) x[j] = fmaxf(x[j], 0.f);
            }
            split_pairs(x, pk);
            return;
          }
          if (head_rows == 0) {                                    // plain hidden layer: ReLU rides on the bf16 conversion (F2FP.RELU)
            if (relu) {
#pragma unroll
              for (int j = 0; j < COLS / 2; ++j) pk[j] = ptx::pack_bf16_relu(x[2 * j], x[2 * j + 1]);
            } else {
#pragma unroll
              for (int j = 0; j < COLS / 2; ++j) pk[j] = ptx::pack_bf16(x[2 * j], x[2 * j + 1]);
            }
            return;
          }
          if (relu) {
#pragma unroll
            for (int j = 0; j < COLS; ++j) x[j] = fmaxf(x[j], 0.f);
          }
          // head partial dot products on the fp32 activations (FFMA2): every row keeps four independent accumulator pairs and the
          // rows are interleaved, so the dependent chain is 4 packed FMAs deep instead of 16 per row (this epilogue sits on the
          // tile's critical path); fully unrolled per head width so hacc[] stays in registers
          auto head_dot = [&](auto rows_tag) {
            constexpr int ROWS = decltype(rows_tag)::value;
            float acc[ROWS][8];
#pragma unroll
            for (int hr = 0; hr < ROWS; ++hr) {
#pragma unroll
              for (int k = 0; k < 8; ++k) acc[hr][k] = 0.f;
            }
#pragma unroll
            for (int q = 0; q < COLS / 4; ++q) {
#pragma unroll
              for (int hr = 0; hr < ROWS; ++hr) {
                const float4 w = reinterpret_cast<const float4*>(hw + hr * n + col0)[q];
                ptx::ffma2(acc[hr][(q & 1) * 4 + 0], acc[hr][(q & 1) * 4 + 1], x[4 * q + 0], x[4 * q + 1], w.x, w.y);
                ptx::ffma2(acc[hr][(q & 1) * 4 + 2], acc[hr][(q & 1) * 4 + 3], x[4 * q + 2], x[4 * q + 3], w.z, w.w);
              }
            }
#pragma unroll
            for (int hr = 0; hr < ROWS; ++hr)
              hacc[hr] += ((acc[hr][0] + acc[hr][1]) + (acc[hr][2] + acc[hr][3])) + ((acc[hr][4] + acc[hr][5]) + (acc[hr][6] + acc[hr][7]));
          };
          if (head_rows == 1) head_dot(std::integral_constant<int, 1>{});
          else if (head_rows == 3) head_dot(std::integral_constant<int, 3>{});
          else head_dot(std::integral_constant<int, 4>{});
          if (split) { split_pairs(x, pk); return; }
#pragma unroll
          for (int j = 0; j < COLS / 2; ++j) pk[j] = ptx::pack_bf16(x[2 * j], x[2 * j + 1]);
        };
        // the slice becomes part of K block `kb` of the next layer's A operand (in place over the old H:
        // every MMA of this layer has retired once d_full fired)
        auto publish = [&](const uint32_t (&pk)[COLS / 2], int kb) {
          if (write_h) {
            // packed pairs go back over the first 16 of the 32 accumulator columns this warp just drained: K block kb of the
            // next layer's A operand never leaves tensor memory
            ptx::tmem_st16(t_row + kb * 64, pk);
            if (split) ptx::tmem_st16(t_row + kb * 64 + 16, pk_lo);      // the other 16 columns of the drained slice
            ptx::tmem_st_wait();
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              // (pair, peer CTA) relaxed: the tensor-memory stores completed at tcgen05.wait::st and were fenced above
              if (remote_arrive) ptx::mbar_arrive_cluster_relaxed(a_ready_addr + (uint32_t)(1 + kb) * 8u);
              else ptx::mbar_arrive_addr(a_ready_addr + (uint32_t)(1 + kb) * 8u);
            }
          }
        };
        // training: the same packed pairs of TWO K blocks go to HBM as tile images, AFTER both blocks have been published to the
        // tensor core (the next layer's MMAs are what the tile waits for; the copies ride under them).  Staged through shared
        // memory: a thread owns 64 bytes of a 128-byte image row, so direct global stores touch 32 lines per instruction; the
        // swizzled staging images leave as 16 KB bulk copies instead, one barrier and one bulk group per pair of blocks.
        auto save_pair = [&](const uint32_t (&pa)[COLS / 2], uint32_t ma, const uint32_t (&pb)[COLS / 2], uint32_t mb, int kb) {
          uint8_t* stg = &sm.w[0][0] + (size_t)SAVE_STAGES * STAGE_BYTES + (size_t)(save_count & 1u) * 2 * IMAGE_BYTES;
#pragma unroll
          for (int u = 0; u < COLS / 8; ++u) {
            *reinterpret_cast<uint4*>(stg + ptx::sw128_offset(row, grp * (COLS / 8) + u)) =
                make_uint4(pa[4 * u], pa[4 * u + 1], pa[4 * u + 2], pa[4 * u + 3]);
            *reinterpret_cast<uint4*>(stg + IMAGE_BYTES + ptx::sw128_offset(row, grp * (COLS / 8) + u)) =
                make_uint4(pb[4 * u], pb[4 * u + 1], pb[4 * u + 2], pb[4 * u + 3]);
          }
          // one mask word per thread and block, 128 bytes per warp (the dgrad chain reads these instead of the 64-byte activation row)
          *reinterpret_cast<uint32_t*>(tile_base + act_mask_offset(args.act_slots, save_slot + kb) + grp * 512 + row * 4) = ma;
          *reinterpret_cast<uint32_t*>(tile_base + act_mask_offset(args.act_slots, save_slot + kb + 1) + grp * 512 + row * 4) = mb;
          ptx::fence_proxy_async_smem();
          // the 32 rows of a lane quarter are 4 KB of each image, written by the TWO warps of that quarter (same SM sub-partition):
          // they meet at their own 64-thread barrier and one of them ships the quarter of both images - no CTA-wide epilogue barrier.
          // The pair of buffers the NEXT pair will use was handed to the copy engine one bulk group ago: it must have been read out by
          // the time both warps pass the barrier.
          const bool q_leader = grp == 0 && lane == 0;
          if (q_leader) ptx::bulk_wait_read<0>();
          asm volatile("bar.sync %0, 64;" ::"r"(5 + quarter) : "memory");
          if (q_leader) {
            const uint64_t once = ptx::l2_policy_evict_first();          // read once, by the backward kernels, much later
            ptx::bulk_s2g_hint(save_base + (size_t)kb * KBLOCK_BYTES + quarter * 4096, stg + quarter * 4096, 4096, once);
            ptx::bulk_s2g_hint(save_base + (size_t)(kb + 1) * KBLOCK_BYTES + quarter * 4096, stg + IMAGE_BYTES + quarter * 4096, 4096, once);
            ptx::bulk_commit();
          }
          ++save_count;
        };

        const int nblocks = n >> 6;
        uint32_t va[COLS], pk[COLS / 2];
        if (save_base == nullptr) {
          for (int kb = 0; kb < nblocks; ++kb) {
            if (kb == 0 || (SRF_MLP_SPLIT && kb == 2)) {            // blocks 2, 3 belong to the second column half
              const uint32_t b = buf * 2 + (SRF_MLP_SPLIT ? (uint32_t)(kb >> 1) : 0u);
              ptx::mbar_wait_addr(d_full_addr + b * 8u, (d_phase >> b) & 1);
              d_phase ^= 1u << b;
              ptx::tc_fence_after();
              if (warp == EPI_WARP0 && lane == 0) TRACE(1024 + l * 16 + (kb ? 9 : 0));
            }
            tmem_load<COLS>(t_row + kb * 64, va);
            ptx::tmem_ld_wait(va);
            if (warp == EPI_WARP0 && lane == 0) TRACE(1024 + l * 16 + 1 + 2 * kb);
            compute(va, pk, kb);       // (a compile-time specialisation of this loop for plain hidden layers measured 8 % SLOWER)
            publish(pk, kb);
            if (warp == EPI_WARP0 && lane == 0) TRACE(1024 + l * 16 + 2 + 2 * kb);
          }
        } else {
          // training: blocks in pairs (layer widths are multiples of 128) - publish both, then save both
          uint32_t pk2[COLS / 2];
          for (int kb = 0; kb < nblocks; kb += 2) {
            if (kb == 0 || SRF_MLP_SPLIT) {
              const uint32_t b = buf * 2 + (SRF_MLP_SPLIT ? (uint32_t)(kb >> 1) : 0u);
              ptx::mbar_wait_addr(d_full_addr + b * 8u, (d_phase >> b) & 1);
              d_phase ^= 1u << b;
              ptx::tc_fence_after();
              if (warp == EPI_WARP0 && lane == 0) TRACE(1024 + l * 16 + (kb ? 9 : 0));
            }
            tmem_load<COLS>(t_row + kb * 64, va);
            ptx::tmem_ld_wait(va);
            compute(va, pk, kb);
            const uint32_t m0 = mask_bits;
            publish(pk, kb);
            if (warp == EPI_WARP0 && lane == 0) TRACE(1024 + l * 16 + 2 + 2 * kb);
            tmem_load<COLS>(t_row + (kb + 1) * 64, va);
            ptx::tmem_ld_wait(va);
            compute(va, pk2, kb + 1);
            const uint32_t m1 = mask_bits;
            publish(pk2, kb + 1);
            if (warp == EPI_WARP0 && lane == 0) TRACE(1024 + l * 16 + 2 + 2 * (kb + 1));
            save_pair(pk, m0, pk2, m1, kb);
          }
        }
        ptx::tc_fence_before();
        if (head_rows > 0) {
          if (warp == EPI_WARP0 && lane == 0) TRACE(2010 + l);
          const int slot = head == 3 ? 1 : 0;
          if (grp != 0) {
#pragma unroll
            for (int hr = 0; hr < 4; ++hr) sm.part[slot][grp][row][hr] = hacc[hr];
          }
          epi_bar_sync();
          if (warp == EPI_WARP0 && lane == 0) TRACE(2020 + l);
          if (grp == 0 && valid) {
            const float* hb = hw + head_rows * n;
            float o[4];
#pragma unroll
            for (int hr = 0; hr < 4; ++hr) {
              float a = hacc[hr];
#pragma unroll
              for (int g = 1; g < GROUPS; ++g) a += sm.part[slot][g][row][hr];
              o[hr] = a + (hr < head_rows ? hb[hr] : 0.f);
            }
            if (head == 1 || head == 2) {
              float sg = o[0];
              if (args.noise != nullptr) sg += args.noise[m];
              args.sigma[m] = fmaxf(sg, 0.f);
            }
            if (head == 2 || head == 3) {            // selects, not o[b + c]: a runtime index would put o[] in local memory
              const float c0 = head == 2 ? o[1] : o[0], c1 = head == 2 ? o[2] : o[1], c2 = head == 2 ? o[3] : o[2];
              float* dst = args.rgb + m * 3;
              dst[0] = __fdividef(1.f, 1.f + __expf(-c0));
              dst[1] = __fdividef(1.f, 1.f + __expf(-c1));
              dst[2] = __fdividef(1.f, 1.f + __expf(-c2));
            }
          }
          if (warp == EPI_WARP0 && lane == 0) TRACE(2000 + l);
        }
      }
    }
  }
  // staged images have left shared memory: every lane quarter's shipping thread drains its own bulk groups
  if (args.save_acts != nullptr && warp >= EPI_WARP0 && warp < EPI_WARP0 + 4 && lane == 0) ptx::bulk_wait_all();
  ptx::tc_fence_before();
  __syncthreads();
  if constexpr (PAIR) ptx::cluster_sync_all();        // the peer may still be reading / being written by the pair's last MMAs
  if (warp == ISSUER1_WARP) {
    ptx::tc_fence_after();
    if constexpr (PAIR) ptx::tmem_dealloc_pair(tmem, 512);
    else ptx::tmem_dealloc(tmem, 512);
  }
}

}  // namespace srf

using namespace srf;

namespace {
int g_pair_mode = -1;          // -1: not decided yet (environment / build default), 0: one CTA per tile, 1: CTA pairs

int validate_program(const MlpProgram& prog, const char* where, bool rows_mode) {
  SRF_REQUIRE(prog.num_layers >= 1 && prog.num_layers <= MLP_MAX_LAYERS, where, "bad layer count");
  SRF_REQUIRE(prog.points_degree >= 0 && prog.points_degree <= 10 && prog.views_degree <= 4, where,
              "encoding degree out of range (points <= 10, views <= 4)");
  SRF_REQUIRE(prog.side_count <= MAX_SIDE, where, "side table too large");
  int used_any = 0;
  for (int l = 0; l < prog.num_layers; ++l) {
    const MlpLayer& L = prog.layers[l];
    SRF_REQUIRE(L.n == 256 || L.n == 128, where, "layer width must be 128 or 256");
    SRF_REQUIRE(L.num_kblocks >= 1 && L.num_kblocks <= MLP_MAX_KBLOCKS, where, "bad K-block count");
    SRF_REQUIRE((L.bias_offset & 3) == 0 && (L.head_offset & 3) == 0, where, "side-table offsets must be multiples of 4");
    int used = 0;
    for (int kb = 0; kb < L.num_kblocks; ++kb) {
      SRF_REQUIRE(L.kblock_region[kb] >= 0 && L.kblock_region[kb] <= 5 && L.kblock_ksteps[kb] >= 1 && L.kblock_ksteps[kb] <= 4,
                  where, "bad K-block descriptor");
      SRF_REQUIRE(!((used >> L.kblock_region[kb]) & 1), where, "a region may appear once per layer");
      used |= 1 << L.kblock_region[kb];
    }
    used_any |= used;
    // barrier generations: whatever an epilogue writes must be consumed in full by the following layer
    if (l == 0) SRF_REQUIRE(used & 1, where, "layer 0 must read region 0");
    if (l > 0) {
      const MlpLayer& P = prog.layers[l - 1];
      const int written = P.write_h ? (((1 << (P.n >> 6)) - 1) << 1) : 0;
      SRF_REQUIRE((used & 0x1E) == written, where, "a layer must read exactly the H blocks the previous epilogue wrote");
    }
    SRF_REQUIRE(!(L.write_h && l == prog.num_layers - 1), where, "last layer cannot write H");
  }
  SRF_REQUIRE(rows_mode ? prog.views_degree < 0 : true, where, "rows mode has no view encoding (views_degree -1: one block, -2: two blocks)");
  SRF_REQUIRE((prog.views_degree >= 0 || (rows_mode && prog.views_degree == -2)) == (((used_any >> 5) & 1) != 0), where,
              "region 5 must be used iff views_degree >= 0 (or -2 in rows mode)");
  return 0;
}

// flatten the layer program into the MMA issuer's K-block steps
MmaSchedule make_schedule(const MlpProgram& prog, bool pair) {
  MmaSchedule sc{};
  int ns = 0, last_e = -1, last_v = -1;
  uint32_t seen = 0;
  const bool split = prog.lo_offset != 0;
  // split-bf16: a K block is three steps (A_hi W_hi, A_lo W_hi, A_hi W_lo).  The lo halves of a hidden block sit 16 columns
  // behind the hi halves in tensor memory; the lo images of E / V occupy the last ring stages.
  const uint32_t smem_lo[2] = {(uint32_t)(offsetof(MlpSmem, w) + (size_t)E_LO_STAGE * STAGE_BYTES - offsetof(MlpSmem, a)) >> 4,
                               (uint32_t)(offsetof(MlpSmem, w) + (size_t)V_LO_STAGE * STAGE_BYTES - offsetof(MlpSmem, a)) >> 4};
  for (int l = 0; l < prog.num_layers; ++l) {
    const MlpLayer& L = prog.layers[l];
    const int halves = SRF_MLP_SPLIT ? L.n >> 7 : 1;
    const int images = SRF_MLP_SPLIT ? 1 : L.n >> 7;      // 128-row weight images per step
    // issue order inside a layer: chunks of SRF_MLP_CHUNK K blocks, both column halves per chunk - the first half completes
    // (and its epilogue starts) while the last chunk of the second half is still on the tensor core, and the early steps
    // only need the first K blocks of the previous layer's output
    for (int c0 = 0; c0 < L.num_kblocks; c0 += SRF_MLP_CHUNK)
    for (int h = 0; h < halves; ++h) {
      for (int kb = c0; kb < L.num_kblocks && kb < c0 + SRF_MLP_CHUNK; ++kb) {
        const int reg = L.kblock_region[kb];
        const int terms = split ? 3 : 1;
        for (int term = 0; term < terms; ++term, ++ns) {
          MmaStep& st = sc.steps[ns];
          const bool a_lo = term == 1, w_lo = term == 2;
          if (reg == 0) st.a_off = a_lo ? smem_lo[0] : 0u;
          else if (reg == 5) st.a_off = a_lo ? smem_lo[1] : (uint32_t)(V_REGION * KBLOCK_BYTES) >> 4;
          else st.a_off = (uint32_t)(reg - 1) * 64u + (a_lo ? 16u : 0u);
          st.idesc = ptx::make_idesc_bf16(pair ? 256 : 128, (uint32_t)(128 * images));      // a CTA pair issues M = 256 (128 rows per CTA)
          const bool wait = !((seen >> reg) & 1);
          seen |= 1u << reg;
          st.meta = (uint32_t)L.kblock_ksteps[kb] | (kb == 0 && term == 0 ? 8u : 0u) |
                    (kb == L.num_kblocks - 1 && term == terms - 1 ? 0x10u : 0u) | ((uint32_t)h << 7) |
                    (wait ? (uint32_t)(reg + 1) << 8 : 0u) | ((uint32_t)(l & 1) << 12) | (reg >= 1 && reg <= 4 ? 0x2000u : 0u) |
                    ((uint32_t)images << 16) | ((uint32_t)L.num_kblocks << 18);
          // blob order: [layer][128-row half][K block]
          st.w_off = (uint32_t)((L.weight_offset + (int64_t)(h * L.num_kblocks + kb) * IMAGE_BYTES + (w_lo ? prog.lo_offset : 0)) >> 4);
          if (reg == 0) last_e = ns;
          if (reg == 5) last_v = ns;
        }
      }
    }
    if (L.write_h) seen &= ~0x1Eu;          // the epilogue of this layer rewrites H: re-acquire its blocks
  }
  if (last_e >= 0) sc.steps[last_e].meta |= 0x20u;
  if (last_v >= 0) sc.steps[last_v].meta |= 0x40u;
  int waits[6] = {0, 0, 0, 0, 0, 0};
  for (int i = 0; i < ns; ++i) {
    const uint32_t wr = (sc.steps[i].meta >> 8) & 15u;
    if (wr) sc.steps[i].meta |= (uint32_t)(waits[wr - 1]++ & 1) << 14;
  }
  for (int i = 0; i < ns; ++i) {
    const uint32_t wr = (sc.steps[i].meta >> 8) & 15u;
    if (wr) sc.steps[i].meta |= (uint32_t)(waits[wr - 1] & 1) << 15;
  }
  sc.num_steps = ns;
  return sc;
}

int schedule_steps(const MlpProgram& prog) {
  int n = 0;
  for (int l = 0; l < prog.num_layers; ++l) n += prog.layers[l].num_kblocks * (SRF_MLP_SPLIT ? prog.layers[l].n >> 7 : 1);
  return n * (prog.lo_offset != 0 ? 3 : 1);
}

int launch_mlp(const MlpProgram& prog, const MlpArgs& a, long long max_total, void* stream, const char* where) {
  const size_t smem = sizeof(MlpSmem);
  SRF_REQUIRE(schedule_steps(prog) <= MAX_STEPS, where, "layer program needs more MMA steps than the schedule holds");
  SRF_REQUIRE(prog.lo_offset == 0 || (prog.lo_offset > 0 && (prog.lo_offset & 15) == 0 && a.save_acts == nullptr && a.rows == nullptr),
              where, "split-bf16 programs (lo_offset != 0) are inference-only: no saved activation tiles, no rows mode");
  const bool split = prog.lo_offset != 0;
  const long long tiles = (max_total + 127) / 128;
  // CTA pairs (cta_group::2) for the bf16 inference forward: srf_mlp_set_pairing(), else the environment (SRF_MLP_PAIR), else the
  // build default.  Same results bit for bit; measured throughput-neutral under the board's power cap
  if (g_pair_mode < 0) { const char* e = getenv("SRF_MLP_PAIR"); g_pair_mode = e != nullptr ? (atoi(e) != 0) : SRF_MLP_PAIR_DEFAULT; }
  const bool pair = g_pair_mode && !split && a.save_acts == nullptr && a.rows == nullptr && tiles >= 2;
  SRF_REQUIRE(a.rows == nullptr || (schedule_steps(prog) <= ROWS_RING && a.row_units <= 16), where,
              "rows mode keeps the weight images resident: at most 4 MMA steps per tile, rows of at most 256 bytes");
  const MmaSchedule sched = make_schedule(prog, pair);
  auto* kernel = pair ? nerf_mlp_fwd_kernel<false, true> : (split ? nerf_mlp_fwd_kernel<true, false> : nerf_mlp_fwd_kernel<false, false>);
  static unsigned long long configured[3] = {0ull, 0ull, 0ull};          // per kernel instantiation, one bit per device
  if (first_use_on_this_device(configured[pair ? 2 : (split ? 1 : 0)])) {
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail(where, cudaGetErrorString(e));
  }
  if (pair) {
    const long long pairs = (tiles + 1) / 2;
    const int clusters = pairs < sm_count() / 2 ? (int)pairs : sm_count() / 2;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(2 * clusters); cfg.blockDim = dim3(MLP_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = (cudaStream_t)stream;
    cudaLaunchAttribute attr;
    attr.id = cudaLaunchAttributeClusterDimension;
    attr.val.clusterDim.x = 2; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
    cfg.attrs = &attr; cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, prog, sched, a);
    if (e != cudaSuccess) return fail(where, cudaGetErrorString(e));
    return check_launch(where);
  }
  const int grid = tiles < sm_count() ? (int)tiles : sm_count();
  kernel<<<grid, MLP_THREADS, smem, (cudaStream_t)stream>>>(prog, sched, a);
  return check_launch(where);
}
}  // namespace

SRF_API int srf_nerf_mlp_fwd(const void* program, const void* weights, const float* side, const float* rays_o,
                             const float* rays_d, const float* z, const float* view_dirs, const float* noise,
                             int64_t num_rays, int num_samples, float* sigma, float* rgb, void* save_acts,
                             int act_slots, int e_slot, int v_slot, void* stream) {
  if (num_rays == 0) return 0;
  SRF_REQUIRE(program && weights && side && rays_o && rays_d && z && sigma && rgb, "srf_nerf_mlp_fwd", "null pointer");
  SRF_REQUIRE(save_acts == nullptr || (act_slots > 0 && e_slot >= 0 && e_slot < act_slots && v_slot < act_slots),
              "srf_nerf_mlp_fwd", "bad activation-save description");
  MlpProgram prog = *reinterpret_cast<const MlpProgram*>(program);
  if (validate_program(prog, "srf_nerf_mlp_fwd", false)) return 1;
  SRF_REQUIRE(prog.views_degree < 0 || view_dirs != nullptr, "srf_nerf_mlp_fwd", "view_dirs required by the program");
  MlpArgs a{};
  a.weights = reinterpret_cast<const uint8_t*>(weights);
  a.side = side; a.rays_o = rays_o; a.rays_d = rays_d; a.z = z; a.view_dirs = view_dirs; a.noise = noise;
  a.sigma = sigma; a.rgb = rgb;
  a.save_acts = reinterpret_cast<uint8_t*>(save_acts);
  a.act_slots = act_slots; a.e_slot = e_slot; a.v_slot = v_slot;
  a.total = (long long)num_rays * num_samples;
  a.S = num_samples;
  return launch_mlp(prog, a, a.total, stream, "srf_nerf_mlp_fwd");
}

SRF_API int srf_mlp_rows_fwd(const void* program, const void* weights, const float* side, const void* rows, int row_pitch,
                             const int* count, int64_t max_rows, float* rgb, void* save_acts, int act_slots, int e_slot, int v_slot,
                             void* stream) {
  if (max_rows == 0) return 0;
  SRF_REQUIRE(program && weights && side && rows && rgb, "srf_mlp_rows_fwd", "null pointer");
  MlpProgram prog = *reinterpret_cast<const MlpProgram*>(program);
  if (validate_program(prog, "srf_mlp_rows_fwd", true)) return 1;
  for (int l = 0; l < prog.num_layers; ++l)
    SRF_REQUIRE(prog.layers[l].head == 0 || prog.layers[l].head == 3, "srf_mlp_rows_fwd", "only the rgb head is supported in rows mode");
  MlpArgs a{};
  a.weights = reinterpret_cast<const uint8_t*>(weights);
  a.side = side; a.rows = reinterpret_cast<const uint4*>(rows); a.count = count; a.rgb = rgb;
  SRF_REQUIRE(row_pitch % 8 == 0 && row_pitch >= 8 && row_pitch <= 128 && (prog.views_degree == -2 || row_pitch <= 64), "srf_mlp_rows_fwd",
              "row_pitch must be a multiple of 8 in 8..128 (<= 64 for a one-block program)");
  SRF_REQUIRE(save_acts == nullptr || (act_slots > 0 && e_slot >= 0 && e_slot < act_slots && v_slot < act_slots), "srf_mlp_rows_fwd",
              "bad saved-tile slots");
  a.save_acts = reinterpret_cast<uint8_t*>(save_acts);
  a.act_slots = act_slots; a.e_slot = e_slot; a.v_slot = v_slot;
  a.row_units = row_pitch / 8;
  a.total = max_rows;
  a.S = 1;
  return launch_mlp(prog, a, max_rows, stream, "srf_mlp_rows_fwd");
}

SRF_API int srf_nerf_mlp_program_bytes(void) { return (int)sizeof(MlpProgram); }

SRF_API int srf_mlp_set_pairing(int mode) {
  const int before = g_pair_mode;
  g_pair_mode = mode < 0 ? -1 : (mode != 0);
  return before;
}

#if SRF_MLP_TRACE
extern "C" __attribute__((visibility("default"))) int srf_debug_mlp_trace(long long* host_out) {
  cudaDeviceSynchronize();
  return cudaMemcpyFromSymbol(host_out, g_mlp_trace, sizeof(long long) * 4096) == cudaSuccess ? 0 : 1;
}
#endif
