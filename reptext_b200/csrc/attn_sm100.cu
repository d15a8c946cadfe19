// Joint text+image flash attention for sm_100a: non-causal, no mask, head_dim 128, bf16 in / fp32 softmax.
// (diffusers dispatch_attention_fn -> F.scaled_dot_product_attention as reached from FluxTransformerBlock /
//  FluxSingleTransformerBlock, RepText/controlnet_flux.py:343-348)
//
// One CTA owns TWO 128-row query tiles (A, B) of one (batch, head) and walks the key/value sequence in
// 128-key tiles.  q and k arrive already RMS-normed and rotated (fused into the QKV GEMM epilogue), so this
// kernel reads q | k | v straight out of the [B, S, 3D] projection buffer with TMA - no concat / permute copies.
//
//   warp 0        TMA producer: Q once, then K and V rings (2 stages each), 128B-swizzled tiles
//   warp 1        tcgen05.mma issuer:  S = Q K^T  (SS, both operands K-major)
//                                      O += P V   (TS: P read from TMEM, V MN-major from smem)
//   warp 2        TMEM allocator (all 512 columns: S_A | S_B | O_A | O_B, 128 fp32 columns each)
//   warps 4-7     softmax warpgroup for tile A: thread r owns query row r (TMEM lane r)
//   warps 8-11    softmax warpgroup for tile B
//
// The tensor pipe ping-pongs between the two tiles: while warpgroup A runs exp2 on S_A[j+1] the MMA warp
// issues P_B V and Q_B K^T.  P (bf16) overwrites the first 64 columns of its own S buffer; the in-order
// execution of tcgen05.mma makes "P V(j) then Q K^T(j+1) into the same columns" safe.  The running max
// is only raised when it grows by more than 2^8 (lazy rescale), so O is almost never read back.  The reference
// exponent is kept INTEGER (ceil of the maximum, in log2 units): P = 2^(x - m) is then the same bf16 mantissa whatever
// m a row happens to carry, so the result does not depend on the ORDER or the split of the key blocks beyond fp32
// summation order - the sequence-parallel mode (keys arrive rank-major) agrees with the single-GPU run to ~1e-5
// where a fractional reference made the two differ by a bf16 rounding of every P (1e-2 after 57 blocks).
#include "dtype_utils.cuh"
#include "ptx_sm100.cuh"
#include "rt_internal.h"
#include "sp_sync.cuh"

namespace rt {
namespace {

constexpr int HD = 128;
constexpr int BQ = 128;   // rows per query tile
constexpr int BKV = 128;  // keys per tile
constexpr int kStages = 2;
constexpr int kThreads = 384;
constexpr int kTileBytes = BQ * HD * 2;       // 32 KB: two 16 KB sub-tiles of 64 columns (one swizzle row each)
constexpr int kSubBytes = kTileBytes / 2;
constexpr int kSmemTiles = 2 + 2 * kStages;   // Q_A Q_B | K ring | V ring
constexpr int kNumBars = 1 + 4 * kStages + 2 + 2 + 2 + 1;
constexpr int kSmemBytes = kSmemTiles * kTileBytes + kNumBars * 8 + 16 + 1024;
constexpr int kThreadsHalfRow = 640;                              // kHalfRow: 4 + 16 warps
constexpr int kXchOff = kSmemTiles * kTileBytes + 256;            // kHalfRow: exchange area behind the barriers
constexpr int kXchFloats = 3 * 512;                               // [2 parities + 1 for the row sums][tile][half][128 rows]
constexpr int kSmemBytesHalfRow = kXchOff + kXchFloats * 4 + 1024;
static_assert(kNumBars * 8 + 16 <= 256, "barrier block");

struct AttnParams {
  CUtensorMap tm;   // (col, row, batch) over the qkv buffer, box (64, 128, 1), SWIZZLE_128B
  CUtensorMap tmh;  // same tensor, box (64, 64, 1): half a K tile per CTA of a pair (v6)
  bf16* out;
  long long out_bs;
  int out_ld, out_col0;
  int q_col0, k_col0, v_col0;
  int S, heads, n_qpairs;
  float scale_log2;  // log2(e) / sqrt(head_dim)
  // sequence-parallel mode: output row r belongs to rank r / sp_rows and is stored straight into that rank's buffer
  int sp_rows;
  bf16* sp_out[RT_SP_MAX_RANKS];
  // sequence-parallel mode, keys in the unsharded order (0 = off): text rows per shard, text key blocks in total
  int sp_txt, sp_txt_blocks;
  // sequence-parallel phase barrier at the head of this kernel (sp_sync.cuh): the TMA warp waits for the peers' q | k | v
  // stores; world == 0: none
  SpSyncParams sync;
};

// Row coordinate(s) of key block j in the (rank-major) buffer.  Plain launches and sp_txt == 0: rows j*128 .. +127 as
// they lie.  Unsharded order: block j < sp_txt_blocks holds 128 text rows = two 64-row pieces of (possibly) different
// ranks' shards; the other blocks are 128 image rows inside one rank's shard.  `half`: 0 / 1 = rows 0-63 / 64-127.
__device__ __forceinline__ int kv_block_row(const AttnParams& P, int j, int half) {
  if (P.sp_txt == 0) return j * BKV + half * 64;
  if (j < P.sp_txt_blocks) {
    const int t = j * BKV + half * 64;  // text row in the unsharded order
    return (t / P.sp_txt) * P.sp_rows + t % P.sp_txt;
  }
  const int n_loc = P.sp_rows - P.sp_txt;
  const int i = (j - P.sp_txt_blocks) * BKV + half * 64;  // image row in the unsharded order
  return (i / n_loc) * P.sp_rows + P.sp_txt + i % n_loc;
}

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  ptx::tmem_ld_32x32b_x32(taddr, r);
  ptx::tmem_ld_wait();
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// 2^x for x <= ~8 on the FMA pipe (Cody-Waite split + cubic minimax, rel. error ~1e-4, far inside bf16's 2^-9):
// the MUFU unit (16 ex2 / clk / SM) is the co-bottleneck of this kernel, so a fixed share of the exponentials
// is computed here instead.
__device__ __forceinline__ float exp2_poly(float x) {
  x = fmaxf(x, -126.f);
  const float xi = __fadd_rd(x, 12582912.f);  // 1.5 * 2^23 + floor(x)
  const float f = x - (xi - 12582912.f);      // [0, 1)
  float p = fmaf(0.077119089663028717f, f, 0.227564394474029541f);
  p = fmaf(p, f, 0.695146143436431885f);
  p = fmaf(p, f, 1.0f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(xi) << 23));
}

// Packed form of the same idea for a PAIR of exponentials (fma.rn.f32x2 / add.f32x2, sm_100): round-to-nearest
// split x = n + f, f in [-0.5, 0.5], cubic minimax of 2^f (max rel. error 1.3e-4), n added into the exponent field.
// 10 instructions per pair against 2 MUFU.EX2 - the MUFU unit (16 / clk / SM) is the unit this kernel saturates.
__device__ __forceinline__ float2 exp2_poly2(float2 x) {
  x.x = fmaxf(x.x, -125.f);
  x.y = fmaxf(x.y, -125.f);
  const float2 magic = make_float2(12582912.f, 12582912.f);  // 1.5 * 2^23: the low mantissa bits become round(x)
  const float2 xi = __fadd2_rn(x, magic);
  const float2 n = __fadd2_rn(xi, make_float2(-12582912.f, -12582912.f));
  const float2 f = __ffma2_rn(n, make_float2(-1.f, -1.f), x);
  float2 p = __ffma2_rn(f, make_float2(0.0550440177f, 0.0550440177f), make_float2(0.24229379f, 0.24229379f));
  p = __ffma2_rn(p, f, make_float2(0.69325459f, 0.69325459f));
  p = __ffma2_rn(p, f, make_float2(0.99994999f, 0.99994999f));
  return make_float2(__int_as_float(__float_as_int(p.x) + (__float_as_int(xi.x) << 23)),
                     __int_as_float(__float_as_int(p.y) + (__float_as_int(xi.y) << 23)));
}

// kDebug: 0 = product; 1 / 2 / 3 are timing experiments (no exp2 / no row max / neither; wrong results).
// kPolyEvery: (scalar form) every kPolyEvery-th PAIR of exponentials goes to exp2_poly (0 = none).
// kPacked: softmax arithmetic on fp32 pairs, TMEM loads of S pipelined against the running max; kPolyMask8: bit
// (i % 8) set = pair i takes exp2_poly2 instead of MUFU.  kSplitP: P is handed to the MMA warp in two halves of
// 64 keys, so that P V can start while the second half is still being exponentiated.
// kHalfRow: TWO threads per query row (640 threads: warps 4-7 / 8-11 = tile A keys 0-63 / 64-127 of the block, warps
// 12-15 / 16-19 = tile B) - two softmax warps per tile on every scheduler instead of one, so that the fixed-latency
// dependencies of one warp (IPC 0.38 in the ncu sampling of the one-thread-per-row form) are covered by the other.  The
// two halves of a row exchange their maxima through shared memory (a 64-thread named barrier per tile and lane
// quadrant), keep partial row sums until the epilogue, and hand P over separately: the first half arrives on p_half,
// the second on p_full (needs kPacked and kSplitP).
// kTrace: CTA 0 records clock64() at every hand-off of every key block into g_attn_trace (rt_debug_attn_trace).
#ifdef RT_AB_VARIANTS
__device__ long long g_attn_trace[128 * 32];
#define RT_ATTN_TRACE_STORE(j, slot) g_attn_trace[(j) * 32 + (slot)] = clock64()
#else
#define RT_ATTN_TRACE_STORE(j, slot) (void)0
#endif
template <int kDebug, int kPolyEvery, bool kPacked = false, int kPolyMask8 = 0, bool kSplitP = false,
          bool kElect = false, bool kHalfRow = false, bool kTrace = false, bool kEarly = false, bool kPipe = false>
__global__ void __launch_bounds__(kHalfRow ? kThreadsHalfRow : kThreads, 1)
    attn_tc_kernel(const __grid_constant__ AttnParams P) {
  constexpr int kBackoff = (kDebug & 8) ? 20 : (kDebug & 16) ? 100 : 0;  // A/B: sleep between barrier polls
  auto trace = [&](int j, int slot) {
    if constexpr (kTrace) {
      if (blockIdx.x == 0 && (threadIdx.x & 31) == 0 && j < 128) RT_ATTN_TRACE_STORE(j, slot);
    }
  };
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_u32 = ptx::smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_u32 + 1023u) & ~1023u) - raw_u32);
  uint8_t* smem_q = smem;                                    // 2 tiles
  uint8_t* smem_k = smem + 2 * kTileBytes;                   // kStages tiles
  uint8_t* smem_v = smem + (2 + kStages) * kTileBytes;       // kStages tiles
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kSmemTiles * kTileBytes);
  uint64_t* q_full = bars;
  uint64_t* k_full = bars + 1;
  uint64_t* k_empty = k_full + kStages;
  uint64_t* v_full = k_empty + kStages;
  uint64_t* v_empty = v_full + kStages;
  uint64_t* s_full = v_empty + kStages;  // [2]
  uint64_t* p_full = s_full + 2;         // [2]
  uint64_t* p_half = p_full + 2;         // [2] first 64 keys of P written (kSplitP)
  uint64_t* o_full = p_half + 2;         // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 1);
  uint64_t* t_bar = bars + 20;  // kTrace: completion of PV_A a / PV_A b / PV_B a / PV_B b, watched by the idle warp 3

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qp = blockIdx.x % P.n_qpairs;
  const int bh = blockIdx.x / P.n_qpairs;
  const int h = bh % P.heads, b = bh / P.heads;
  const int q0 = qp * 2 * BQ;
  const int n_kv = (P.S + BKV - 1) / BKV;

  if (warp == 0 && lane == 0) ptx::prefetch_tmap(&P.tm);
  if (warp == 1 && lane == 0) {
    ptx::mbar_init(q_full, 1);
    for (int i = 0; i < kStages; ++i) {
      ptx::mbar_init(&k_full[i], 1);
      ptx::mbar_init(&k_empty[i], 1);
      ptx::mbar_init(&v_full[i], 1);
      ptx::mbar_init(&v_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&s_full[i], 1);
      ptx::mbar_init(&p_full[i], kElect ? 4 : 128);  // kElect: one lane per softmax warp arrives
      ptx::mbar_init(&p_half[i], kElect ? 4 : 128);
    }
    ptx::mbar_init(o_full, 1);
    if constexpr (kTrace)
      for (int i = 0; i < 4; ++i) ptx::mbar_init(&t_bar[i], 1);
    ptx::fence_barrier_init();
  }
  if (warp == 2) ptx::tmem_alloc<1>(tmem_slot, 512);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  ptx::grid_launch_dependents();  // programmatic dependent launch: nothing above reads the previous kernel's output
  ptx::grid_dependency_wait();

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
  if (kTrace && warp == 3) {
    // observer: tensor-pipe completion times (tcgen05.commit arrivals), in the order the MMA warp issues them
    for (int j = 0; j < n_kv; ++j) {
      ptx::mbar_wait<kBackoff>(&t_bar[0], j & 1); trace(j, 20);
      ptx::mbar_wait<kBackoff>(&t_bar[1], j & 1); trace(j, 21);
      if (j + 1 < n_kv) { ptx::mbar_wait<kBackoff>(&s_full[0], (j + 1) & 1); trace(j, 22); }
      ptx::mbar_wait<kBackoff>(&t_bar[2], j & 1); trace(j, 23);
      ptx::mbar_wait<kBackoff>(&t_bar[3], j & 1); trace(j, 24);
      if (j + 1 < n_kv) { ptx::mbar_wait<kBackoff>(&s_full[1], (j + 1) & 1); trace(j, 25); }
    }
  }
  if (warp == 0) spsync::sp_barrier_head(P.sync, lane, blockIdx.x == 0);  // the peers' q | k | v stores (sp_sync.cuh)
  if (warp == 0 && lane == 0) {
    // ===================== TMA producer =====================
    ptx::mbar_arrive_expect_tx(q_full, 2 * kTileBytes);
#pragma unroll
    for (int t = 0; t < 2; ++t)
#pragma unroll
      for (int sub = 0; sub < 2; ++sub)
        ptx::tma_load_3d(&P.tm, q_full, smem_q + t * kTileBytes + sub * kSubBytes, P.q_col0 + h * HD + sub * 64,
                         q0 + t * BQ, b);
    for (int j = 0; j < n_kv; ++j) {
      const int st = j % kStages, ph = (j / kStages) & 1;
      // a text block of the unsharded order is two 64-row pieces (box 64 x 64); every other block one 128-row box
      const bool split = P.sp_txt != 0 && j < P.sp_txt_blocks;
      const int r0 = kv_block_row(P, j, 0), r1 = kv_block_row(P, j, 1);
      auto load_tile = [&](uint64_t* bar, uint8_t* dst, int col0) {
#pragma unroll
        for (int sub = 0; sub < 2; ++sub) {
          if (split) {
            ptx::tma_load_3d(&P.tmh, bar, dst + sub * kSubBytes, col0 + sub * 64, r0, b);
            ptx::tma_load_3d(&P.tmh, bar, dst + sub * kSubBytes + 64 * 128, col0 + sub * 64, r1, b);
          } else {
            ptx::tma_load_3d(&P.tm, bar, dst + sub * kSubBytes, col0 + sub * 64, r0, b);
          }
        }
      };
      ptx::mbar_wait<kBackoff>(&k_empty[st], ph ^ 1);
      ptx::mbar_arrive_expect_tx(&k_full[st], kTileBytes);
      load_tile(&k_full[st], smem_k + st * kTileBytes, P.k_col0 + h * HD);
      ptx::mbar_wait<kBackoff>(&v_empty[st], ph ^ 1);
      ptx::mbar_arrive_expect_tx(&v_full[st], kTileBytes);
      load_tile(&v_full[st], smem_v + st * kTileBytes, P.v_col0 + h * HD);
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // The whole warp runs the loop (so that descriptor arithmetic stays on the uniform datapath); one elected
    // lane issues the tcgen05 instructions.
    constexpr uint32_t idesc_qk = ptx::make_idesc_bf16(BQ, BKV, 0, 0);  // A, B K-major
    constexpr uint32_t idesc_pv = ptx::make_idesc_bf16(BQ, HD, 0, 1);   // A from TMEM, B (= V) MN-major
    const uint64_t q_desc = ptx::make_smem_desc_sw128(ptx::smem_u32(smem_q), 0, 1024);
    const uint64_t k_desc = ptx::make_smem_desc_sw128(ptx::smem_u32(smem_k), 0, 1024);
    const uint64_t v_desc = ptx::make_smem_desc_sw128(ptx::smem_u32(smem_v), kSubBytes, 1024);
    constexpr uint32_t kTile16 = kTileBytes >> 4, kSub16 = kSubBytes >> 4;
    auto issue_qk = [&](int t, int st) {
      const uint64_t qa = q_desc + (uint64_t)(t * kTile16), ka = k_desc + (uint64_t)(st * kTile16);
#pragma unroll
      for (int kk = 0; kk < HD / 16; ++kk) {
        const uint32_t off = (kk >> 2) * kSub16 + (kk & 3) * 2;  // (addr >> 4) units
        ptx::mma_bf16_ss<1>(tmem + t * 128, qa + off, ka + off, idesc_qk, kk != 0 ? 1u : 0u);
      }
    };
    auto issue_pv = [&](int t, int st, uint32_t acc, int kk0 = 0, int kk1 = BKV / 16) {
      const uint64_t va = v_desc + (uint64_t)(st * kTile16);
#pragma unroll
      for (int kk = kk0; kk < kk1; ++kk) {
        // 16 keys = two 8-row core groups (SBO 1024 B); 128 head-dim columns = two 64-column sub-tiles (LBO 16 KB)
        ptx::mma_bf16_ts(tmem + 256 + t * 128, tmem + t * 128 + kk * 8, va + (uint64_t)(kk * 128), idesc_pv,
                         kk != 0 ? 1u : acc);
      }
    };
    ptx::mbar_wait<kBackoff>(q_full, 0);
    ptx::mbar_wait<kBackoff>(&k_full[0], 0);
    ptx::tc_fence_after();
    if (ptx::elect_one()) {
      issue_qk(0, 0);
      ptx::mma_commit(&s_full[0]);
      issue_qk(1, 0);
      ptx::mma_commit(&s_full[1]);
      ptx::mma_commit(&k_empty[0]);
    }
    __syncwarp();
    for (int j = 0; j < n_kv; ++j) {
      const int st = j % kStages, ph = (j / kStages) & 1;
      const int nst = (j + 1) % kStages, nph = ((j + 1) / kStages) & 1;
      const bool more = j + 1 < n_kv;
      const uint32_t acc = j > 0 ? 1u : 0u;
      ptx::mbar_wait<kBackoff>(&v_full[st], ph);
      trace(j, 0);
      if constexpr (kSplitP) {
        ptx::mbar_wait<kBackoff>(&p_half[0], j & 1);
        trace(j, 1);
        ptx::tc_fence_after();
        if (ptx::elect_one()) {
          issue_pv(0, st, acc, 0, 4);
          if constexpr (kTrace) ptx::mma_commit(&t_bar[0]);
        }
        __syncwarp();
      }
      ptx::mbar_wait<kBackoff>(&p_full[0], j & 1);
      trace(j, 2);
      if (more) ptx::mbar_wait<kBackoff>(&k_full[nst], nph);
      ptx::tc_fence_after();
      if (ptx::elect_one()) {
        if constexpr (kSplitP) issue_pv(0, st, 1u, 4, 8); else issue_pv(0, st, acc);
        if constexpr (kTrace) ptx::mma_commit(&t_bar[1]);
        if (more) {
          issue_qk(0, nst);
          ptx::mma_commit(&s_full[0]);
        }
      }
      __syncwarp();
      trace(j, 3);
      if constexpr (kSplitP) {
        ptx::mbar_wait<kBackoff>(&p_half[1], j & 1);
        trace(j, 4);
        ptx::tc_fence_after();
        if (ptx::elect_one()) {
          issue_pv(1, st, acc, 0, 4);
          if constexpr (kTrace) ptx::mma_commit(&t_bar[2]);
        }
        __syncwarp();
      }
      ptx::mbar_wait<kBackoff>(&p_full[1], j & 1);
      trace(j, 5);
      ptx::tc_fence_after();
      if (ptx::elect_one()) {
        if constexpr (kSplitP) issue_pv(1, st, 1u, 4, 8); else issue_pv(1, st, acc);
        if constexpr (kTrace) ptx::mma_commit(&t_bar[3]);
        ptx::mma_commit(&v_empty[st]);
        if (more) {
          issue_qk(1, nst);
          ptx::mma_commit(&s_full[1]);
          ptx::mma_commit(&k_empty[nst]);
        }
      }
      __syncwarp();
      trace(j, 6);
    }
    if (ptx::elect_one()) ptx::mma_commit(o_full);
    __syncwarp();
  }
  } else if constexpr (kHalfRow) {
    // ===================== softmax, two threads per row =====================
    static_assert(!kHalfRow || (kPacked && kSplitP && !kElect), "kHalfRow needs the packed, split-P, all-threads-arrive form");
    // register budget: 640 x 96 at launch = 4 x 32 x 64 (after the dec above) + 16 x 32 x 104; the inc can only draw on what the
    // CTA itself released, a larger request blocks for ever
    asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
    const int idx = warp - 4;
    const int quad = idx & 3;        // TMEM lane quadrant (= warp % 4)
    const int hf = (idx >> 2) & 1;   // keys hf * 64 .. hf * 64 + 63 of every block
    const int t = idx >> 3;          // 0: tile A, 1: tile B
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t s_addr = tmem + lane_off + t * 128;
    const uint32_t o_addr = tmem + lane_off + 256 + t * 128;
    const int bar_id = 1 + t * 4 + quad;  // named barrier of the two warps that share these 32 rows
    auto pair_sync = [&]() { asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory"); };
    float* xch = reinterpret_cast<float*>(smem + kXchOff);
    float* x_mine = xch + (t * 2 + hf) * 128 + quad * 32 + lane;
    const float* x_other = xch + (t * 2 + (hf ^ 1)) * 128 + quad * 32 + lane;
    const float c = P.scale_log2;
    float m_ref = -INFINITY, l = 0.f;
    for (int j = 0; j < n_kv; ++j) {
      ptx::mbar_wait<kBackoff>(&s_full[t], j & 1);
      if (quad == 0 && hf == 0) trace(j, 8 + t * 4);
      ptx::tc_fence_after();
      const int n_valid = P.S - j * BKV - hf * 64;  // valid keys of this half (<= 0: none) - only short on the last block
      uint32_t s0[32], s1[32];
      auto chunk_max = [&](uint32_t (&sv)[32], int col0) {
        if (n_valid < 64) {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (col0 + i >= n_valid) sv[i] = 0xff800000u;  // -inf
        }
        float a = -INFINITY, b2 = -INFINITY;
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          a = fmaxf(a, fmaxf(__uint_as_float(sv[i]), __uint_as_float(sv[i + 1])));
          b2 = fmaxf(b2, fmaxf(__uint_as_float(sv[i + 2]), __uint_as_float(sv[i + 3])));
        }
        return fmaxf(a, b2);
      };
      ptx::tmem_ld_32x32b_x32(s_addr + hf * 64, s0);
      ptx::tmem_ld_wait();
      ptx::tmem_ld_32x32b_x32(s_addr + hf * 64 + 32, s1);
      float mx = chunk_max(s0, 0);
      ptx::tmem_ld_wait();
      mx = fmaxf(mx, chunk_max(s1, 32));
      // the other half of the row: both threads now hold their scores in registers, so after this barrier the second
      // half may overwrite columns 32..63 (P) that the first half has just read as scores
      x_mine[(j & 1) * 512] = mx;
      if (quad == 0 && hf == 0) trace(j, 9 + t * 4);
      pair_sync();
      if (quad == 0) trace(j, 26 + t * 2 + hf);
      mx = fmaxf(mx, x_other[(j & 1) * 512]);
      const float mx_s = mx * c;
      if (j == 0) {
        m_ref = ceilf(mx_s);
      } else if (__any_sync(0xffffffffu, mx_s > m_ref + 8.f)) {  // same rows, same values: both halves decide alike
        const float m_new = ceilf(fmaxf(m_ref, mx_s));
        const float f = ptx::ex2_approx(m_ref - m_new);
        l *= f;
#pragma unroll 1
        for (int ch = 0; ch < 4; ++ch) {  // this half's 64 columns of O
          uint32_t r[16];
          ptx::tmem_ld_32x32b_x16(o_addr + hf * 64 + ch * 16, r);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * f);
          ptx::tmem_st_32x32b_x16(o_addr + hf * 64 + ch * 16, r);
        }
        m_ref = m_new;
      }
      const float2 c2 = make_float2(c, c), nm2 = make_float2(-m_ref, -m_ref);
      float2 lsum = make_float2(0.f, 0.f);
      auto exp_chunk = [&](const uint32_t (&sv)[32], int col) {
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float2 x = __ffma2_rn(make_float2(__uint_as_float(sv[2 * i]), __uint_as_float(sv[2 * i + 1])), c2, nm2);
          float2 e;
          if ((kPolyMask8 >> (i & 7)) & 1) {
            e = exp2_poly2(x);
          } else {
            e.x = ptx::ex2_approx(x.x);
            e.y = ptx::ex2_approx(x.y);
          }
          lsum = __fadd2_rn(lsum, e);
          pk[i] = ptx::pack_bf16x2(e.x, e.y);
        }
        ptx::tmem_st_32x32b_x16(s_addr + col, pk);
      };
      exp_chunk(s0, hf * 32);
      exp_chunk(s1, hf * 32 + 16);
      l += lsum.x + lsum.y;
      ptx::tmem_st_wait();
      ptx::tc_fence_before();
      ptx::mbar_arrive(hf == 0 ? &p_half[t] : &p_full[t]);
      if (quad == 0) trace(j, 10 + t * 4 + hf);
    }
    // ---- epilogue: as below, the two halves of a row share one staging tile and split the copy-out
    ptx::mbar_wait<kBackoff>(o_full, 0);
    ptx::tc_fence_after();
    x_mine[2 * 512] = l;
    pair_sync();
    const float inv = 1.f / (l + x_other[2 * 512]);
    constexpr int kPitch = HD * 2 + 16;
    uint8_t* stage = smem + (t * 4 + quad) * (32 * kPitch);  // Q / K tiles are dead (o_full)
#pragma unroll 1
    for (int ch = 0; ch < 2; ++ch) {
      float v[32];
      tmem_ld32(o_addr + hf * 64 + ch * 32, v);
      uint4* dst = reinterpret_cast<uint4*>(stage + lane * kPitch + (hf * 2 + ch) * 64);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        uint4 u;
        u.x = ptx::pack_bf16x2(v[8 * i + 0] * inv, v[8 * i + 1] * inv);
        u.y = ptx::pack_bf16x2(v[8 * i + 2] * inv, v[8 * i + 3] * inv);
        u.z = ptx::pack_bf16x2(v[8 * i + 4] * inv, v[8 * i + 5] * inv);
        u.w = ptx::pack_bf16x2(v[8 * i + 6] * inv, v[8 * i + 7] * inv);
        dst[i] = u;
      }
    }
    pair_sync();
    const int row0 = q0 + t * BQ + quad * 32;
    const int rr = lane >> 4, cc = lane & 15;
#pragma unroll 4
    for (int it = hf * 8; it < hf * 8 + 8; ++it) {
      const int r = it * 2 + rr;
      const int grow = row0 + r;
      if (grow < P.S) {
        bf16* orow;
        if (P.sp_rows > 0) {
          const int dest = grow / P.sp_rows;
          orow = P.sp_out[dest] + (long long)b * P.out_bs + (long long)(grow - dest * P.sp_rows) * P.out_ld +
                 P.out_col0 + h * HD;
        } else {
          orow = P.out + (long long)b * P.out_bs + (long long)grow * P.out_ld + P.out_col0 + h * HD;
        }
        *reinterpret_cast<uint4*>(orow + cc * 8) = *reinterpret_cast<const uint4*>(stage + r * kPitch + cc * 16);
      }
    }
  } else {
    // ===================== softmax warpgroups =====================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 216;");
    const int t = (warp - 4) >> 2;  // 0: tile A, 1: tile B
    const int quad = warp & 3;
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t s_addr = tmem + lane_off + t * 128;
    const uint32_t o_addr = tmem + lane_off + 256 + t * 128;
    const float c = P.scale_log2;
    float m_ref = -INFINITY, l = 0.f;
    for (int j = 0; j < n_kv; ++j) {
      ptx::mbar_wait<kBackoff>(&s_full[t], j & 1);
      if (quad == 0) trace(j, 8 + t * 4);
      ptx::tc_fence_after();
      const int n_valid = P.S - j * BKV;  // < 128 only on the last tile
      if constexpr (kPacked) {
        // ---- S -> registers, one chunk in flight while the previous one feeds the running max
        uint32_t s0[32], s1[32], s2[32], s3[32];
        auto chunk_max = [&](uint32_t (&sv)[32], int col0) {
          if (n_valid < BKV) {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (col0 + i >= n_valid) sv[i] = 0xff800000u;  // -inf
          }
          float a = -INFINITY, b2 = -INFINITY;
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            a = fmaxf(a, fmaxf(__uint_as_float(sv[i]), __uint_as_float(sv[i + 1])));
            b2 = fmaxf(b2, fmaxf(__uint_as_float(sv[i + 2]), __uint_as_float(sv[i + 3])));
          }
          return fmaxf(a, b2);
        };
        // kEarly: the first 32 keys are exponentiated against the CURRENT reference while the other 96 scores are
        // still on their way from tensor memory (the loads' latency, ~260 cycles, was on the tile's critical chain);
        // the reference is final once the block's maximum is known - in the rare blocks that raise it (and in block
        // 0) those 32 are simply computed again.  Same reference, same arithmetic, same results as the plain form.
        const float2 c2 = make_float2(c, c);
        uint32_t pk0[16];
        float2 ls0[4] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
        auto exp_regs = [&](const uint32_t (&sv)[32], const float2 nm2e) {
#pragma unroll
          for (int i = 0; i < 4; ++i) ls0[i] = make_float2(0.f, 0.f);
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            float2 x = __ffma2_rn(make_float2(__uint_as_float(sv[2 * i]), __uint_as_float(sv[2 * i + 1])), c2, nm2e);
            float2 e;
            if ((kPolyMask8 >> (i & 7)) & 1) {
              e = exp2_poly2(x);
            } else {
              e.x = ptx::ex2_approx(x.x);
              e.y = ptx::ex2_approx(x.y);
            }
            ls0[i & 3] = __fadd2_rn(ls0[i & 3], e);
            pk0[i] = ptx::pack_bf16x2(e.x, e.y);
          }
        };
        bool redo0 = true;
        float mx;
        if constexpr (kEarly) {
          ptx::tmem_ld_32x32b_x32(s_addr, s0);
          ptx::tmem_ld_wait();
          ptx::tmem_ld_32x32b_x32(s_addr + 32, s1);
          ptx::tmem_ld_32x32b_x32(s_addr + 64, s2);
          ptx::tmem_ld_32x32b_x32(s_addr + 96, s3);
          mx = chunk_max(s0, 0);
          if (j > 0) {
            exp_regs(s0, make_float2(-m_ref, -m_ref));
            redo0 = false;
          }
          ptx::tmem_ld_wait();
          mx = fmaxf(fmaxf(mx, chunk_max(s1, 32)), fmaxf(chunk_max(s2, 64), chunk_max(s3, 96)));
        } else {
          ptx::tmem_ld_32x32b_x32(s_addr, s0);
          ptx::tmem_ld_wait();
          ptx::tmem_ld_32x32b_x32(s_addr + 32, s1);
          mx = chunk_max(s0, 0);
          ptx::tmem_ld_wait();
          ptx::tmem_ld_32x32b_x32(s_addr + 64, s2);
          mx = fmaxf(mx, chunk_max(s1, 32));
          ptx::tmem_ld_wait();
          ptx::tmem_ld_32x32b_x32(s_addr + 96, s3);
          mx = fmaxf(mx, chunk_max(s2, 64));
          ptx::tmem_ld_wait();
          mx = fmaxf(mx, chunk_max(s3, 96));
        }
        const float mx_s = mx * c;
        if (quad == 0) trace(j, 9 + t * 4);
        if (j == 0) {
          m_ref = ceilf(mx_s);
        } else if (__any_sync(0xffffffffu, mx_s > m_ref + 8.f)) {
          redo0 = true;
          if (quad == 0) trace(j, 16 + t);
          const float m_new = ceilf(fmaxf(m_ref, mx_s));
          const float f = ptx::ex2_approx(m_ref - m_new);
          l *= f;
#pragma unroll 1
          for (int ch = 0; ch < 8; ++ch) {
            uint32_t r[16];
            ptx::tmem_ld_32x32b_x16(o_addr + ch * 16, r);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * f);
            ptx::tmem_st_32x32b_x16(o_addr + ch * 16, r);
          }
          m_ref = m_new;
        }
        // ---- P = 2^(S c - m) on pairs; bf16 P overwrites the first 64 columns of S
        const float2 nm2 = make_float2(-m_ref, -m_ref);
        if constexpr (kPipe) {
          // SOFTWARE-PIPELINED form (tools/softmax_rate_probe.cu, profiles/r2_softmax_pipelined_probe.txt): with one or
          // two warps on a scheduler the straight-line form below runs at 10-18 cycles per key - the MUFU pipe (8 cycles
          // per exponential instruction) and the FMA pipe take turns instead of running side by side, because nothing
          // fixes the distance between an exponential and its consumers.  Here a ROLLED loop walks the row in 16-key
          // chunks and every half-iteration interleaves three stages of three different chunks: the exponentials of
          // chunk ch, the scale-subtract of chunk ch + 1, the row sum / bf16 packing / store of chunk ch - 1 - every
          // consumer sits a whole half-iteration behind its producer.  The loop needs dynamic addressing, so the scores
          // come from tensor memory a second time (16 columns per load, issued one half-iteration ahead; the first two
          // chunks are still in registers from the maximum pass).  P chunk ch - 1 (8 columns) lands on score columns
          // that were read at least two half-iterations earlier.  Same arithmetic in the same order as the
          // straight-line form: bit-identical results.
          float2 lsum[4] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
          float2 xA[8], eA[8], xB[8], eB[8];
          uint32_t rawA[16], rawB[16];
          const uint32_t s_base = __shfl_sync(0xffffffffu, s_addr, 0);  // warp-uniform, and the compiler can see it
#pragma unroll
          for (int i = 0; i < 8; ++i)
            xA[i] = __ffma2_rn(make_float2(__uint_as_float(s0[2 * i]), __uint_as_float(s0[2 * i + 1])), c2, nm2);
#pragma unroll
          for (int i = 0; i < 16; ++i) rawB[i] = s0[16 + i];
          // one half-iteration: exps of chunk ch (xc -> ec) | scale-subtract of chunk ch + 1 (rawn -> xn) | sum / pack /
          // store of chunk ch - 1 (ep).  kFirst: no chunk ch - 1; kLast: no chunk ch + 1.  The raw set `rawf` - its last
          // reader ran a half-iteration ago - receives chunk pf (ch + 2; past the end a harmless reload of chunk 7).
          auto half = [&](auto first_tag, auto last_tag, auto masked_tag, float2 (&xc)[8], float2 (&ec)[8],
                          uint32_t (&rawn)[16], float2 (&xn)[8], float2 (&ep)[8], uint32_t (&rawf)[16], int ch) {
            constexpr bool kFirst = decltype(first_tag)::value, kLast = decltype(last_tag)::value;
            constexpr bool kMasked = decltype(masked_tag)::value;
            if constexpr (!kLast) {
              ptx::tmem_ld_wait();
              const int pf = ch + 2 < 8 ? ch + 2 : 7;
              ptx::tmem_ld_32x32b_x16(s_base + pf * 16, rawf);
              if constexpr (kMasked) {
#pragma unroll
                for (int i = 0; i < 16; ++i)
                  if ((ch + 1) * 16 + i >= n_valid) rawn[i] = 0xff800000u;  // -inf
              }
            }
            uint32_t pk[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              constexpr int kM = kPolyMask8;
              const bool poly = (kM >> (i & 7)) & 1;
              if (!poly) ec[i].x = ptx::ex2_approx(xc[i].x);
              if constexpr (!kLast)
                xn[i] = __ffma2_rn(make_float2(__uint_as_float(rawn[2 * i]), __uint_as_float(rawn[2 * i + 1])), c2, nm2);
              if constexpr (!kFirst) lsum[i & 3] = __fadd2_rn(lsum[i & 3], ep[i]);
              if (!poly) ec[i].y = ptx::ex2_approx(xc[i].y);
              else ec[i] = exp2_poly2(xc[i]);
              if constexpr (!kFirst) pk[i] = ptx::pack_bf16x2(ep[i].x, ep[i].y);
            }
            if constexpr (!kFirst) ptx::tmem_st_32x32b_x8(s_base + (ch - 1) * 8, pk);
          };
          auto run = [&](auto masked_tag) {
            using T = std::true_type;
            using F = std::false_type;
            half(T{}, F{}, masked_tag, xA, eA, rawB, xB, eB, rawA, 0);
#pragma unroll 1
            for (int ch = 1; ch < 7; ch += 2) {
              half(F{}, F{}, masked_tag, xB, eB, rawA, xA, eA, rawB, ch);
              half(F{}, F{}, masked_tag, xA, eA, rawB, xB, eB, rawA, ch + 1);
              if constexpr (kSplitP) {
                if (ch == 3) {  // P chunks 0..3 (keys 0-63) are on their way: P V may start on them
                  ptx::tmem_st_wait();
                  ptx::tc_fence_before();
                  ptx::mbar_arrive(&p_half[t]);
                  if (quad == 0) trace(j, 10 + t * 4);
                }
              }
            }
            half(F{}, T{}, masked_tag, xB, eB, rawA, xA, eA, rawB, 7);
          };
          if (n_valid < BKV) run(std::true_type{});
          else run(std::false_type{});
          {  // drain: chunk 7
            uint32_t pk[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              lsum[i & 3] = __fadd2_rn(lsum[i & 3], eB[i]);
              pk[i] = ptx::pack_bf16x2(eB[i].x, eB[i].y);
            }
            ptx::tmem_st_32x32b_x8(s_addr + 7 * 8, pk);
          }
          {
            const float2 a = __fadd2_rn(lsum[0], lsum[1]), b2 = __fadd2_rn(lsum[2], lsum[3]);
            l += (a.x + a.y) + (b2.x + b2.y);
          }
          ptx::tmem_st_wait();
          ptx::tc_fence_before();
          ptx::mbar_arrive(&p_full[t]);
          if (quad == 0) trace(j, 11 + t * 4);
          continue;
        }
        // four independent partial sums: one warp per tile and scheduler is latency-bound, a single FADD2 chain of 64
        // links per block is one of the latencies
        float2 lsum[4] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
        auto exp_chunk = [&](const uint32_t (&sv)[32], int col) {
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            float2 x = __ffma2_rn(make_float2(__uint_as_float(sv[2 * i]), __uint_as_float(sv[2 * i + 1])), c2, nm2);
            float2 e;
            if ((kPolyMask8 >> (i & 7)) & 1) {
              e = exp2_poly2(x);
            } else {
              e.x = ptx::ex2_approx(x.x);
              e.y = ptx::ex2_approx(x.y);
            }
            lsum[i & 3] = __fadd2_rn(lsum[i & 3], e);
            pk[i] = ptx::pack_bf16x2(e.x, e.y);
          }
          ptx::tmem_st_32x32b_x16(s_addr + col, pk);
        };
        if constexpr (kEarly) {
          if (redo0) exp_regs(s0, nm2);  // warp-uniform: block 0, or the reference has just been raised
#pragma unroll
          for (int i = 0; i < 4; ++i) lsum[i] = ls0[i];
          ptx::tmem_st_32x32b_x16(s_addr, pk0);
        } else {
          exp_chunk(s0, 0);
        }
        exp_chunk(s1, 16);
        if constexpr (kSplitP) {
          ptx::tmem_st_wait();
          ptx::tc_fence_before();
          if constexpr (kElect) {
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&p_half[t]);
          } else {
            ptx::mbar_arrive(&p_half[t]);
          }
          if (quad == 0) trace(j, 10 + t * 4);
        }
        exp_chunk(s2, 32);
        exp_chunk(s3, 48);
        {
          const float2 a = __fadd2_rn(lsum[0], lsum[1]), b2 = __fadd2_rn(lsum[2], lsum[3]);
          l += (a.x + a.y) + (b2.x + b2.y);
        }
        ptx::tmem_st_wait();
        ptx::tc_fence_before();
        if constexpr (kElect) {
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(&p_full[t]);
        } else {
          ptx::mbar_arrive(&p_full[t]);
        }
        if (quad == 0) trace(j, 11 + t * 4);
        continue;
      }
      // the whole score row (128 fp32) in registers: ONE TMEM round trip per tile
      uint32_t s0[32], s1[32], s2[32], s3[32];
      if (kDebug & 4) {  // timing experiment: no TMEM read of S at all
#pragma unroll
        for (int i = 0; i < 32; ++i) s0[i] = s1[i] = s2[i] = s3[i] = 0x3f000000u + (uint32_t)(i + j);
      } else {
        ptx::tmem_ld_32x32b_x32(s_addr, s0);
        ptx::tmem_ld_32x32b_x32(s_addr + 32, s1);
        ptx::tmem_ld_32x32b_x32(s_addr + 64, s2);
        ptx::tmem_ld_32x32b_x32(s_addr + 96, s3);
        ptx::tmem_ld_wait();
      }
      if (n_valid < BKV) {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          if (i >= n_valid) s0[i] = 0xff800000u;  // -inf
          if (32 + i >= n_valid) s1[i] = 0xff800000u;
          if (64 + i >= n_valid) s2[i] = 0xff800000u;
          if (96 + i >= n_valid) s3[i] = 0xff800000u;
        }
      }
      float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
      if (!(kDebug & 2))
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        mx0 = fmaxf(mx0, __uint_as_float(s0[i]));
        mx1 = fmaxf(mx1, __uint_as_float(s1[i]));
        mx2 = fmaxf(mx2, __uint_as_float(s2[i]));
        mx3 = fmaxf(mx3, __uint_as_float(s3[i]));
      }
      float mx_s = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * c;
      if (kDebug & 2) mx_s = 0.f;
      if (j == 0) {
        m_ref = ceilf(mx_s);
      } else if (__any_sync(0xffffffffu, mx_s > m_ref + 8.f)) {
        const float m_new = ceilf(fmaxf(m_ref, mx_s));
        const float f = ptx::ex2_approx(m_ref - m_new);
        l *= f;
#pragma unroll 1
        for (int ch = 0; ch < 8; ++ch) {
          uint32_t r[16];
          ptx::tmem_ld_32x32b_x16(o_addr + ch * 16, r);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * f);
          ptx::tmem_st_32x32b_x16(o_addr + ch * 16, r);
        }
        m_ref = m_new;
      }
      float l0 = 0.f, l1 = 0.f;
#define RT_SOFTMAX_CHUNK(SV, COL)                                                            \
      {                                                                                      \
        uint32_t pk[16];                                                                     \
        _Pragma("unroll") for (int i = 0; i < 16; ++i) {                                     \
          float p0 = fmaf(__uint_as_float(SV[2 * i]), c, -m_ref);                            \
          float p1 = fmaf(__uint_as_float(SV[2 * i + 1]), c, -m_ref);                        \
          if (!(kDebug & 1)) {                                                               \
            if (kPolyEvery > 0 && (i % (kPolyEvery > 0 ? kPolyEvery : 1)) == kPolyEvery - 1) { \
              p0 = exp2_poly(p0); p1 = exp2_poly(p1);                                        \
            } else {                                                                         \
              p0 = ptx::ex2_approx(p0); p1 = ptx::ex2_approx(p1);                            \
            }                                                                                \
          }                                                                                  \
          l0 += p0;                                                                          \
          l1 += p1;                                                                          \
          pk[i] = ptx::pack_bf16x2(p0, p1);                                                  \
        }                                                                                    \
        ptx::tmem_st_32x32b_x16(s_addr + COL, pk);                                           \
      }
      RT_SOFTMAX_CHUNK(s0, 0)
      RT_SOFTMAX_CHUNK(s1, 16)
      RT_SOFTMAX_CHUNK(s2, 32)
      RT_SOFTMAX_CHUNK(s3, 48)
#undef RT_SOFTMAX_CHUNK
      l += l0 + l1;
      ptx::tmem_st_wait();
      ptx::tc_fence_before();
      ptx::mbar_arrive(&p_full[t]);
    }
    // ---- epilogue: O / l -> bf16 -> shared (row-wise) -> global (2 rows x 256 B per warp instruction).
    // TMEM gives each thread one ROW; storing it directly would write 32 rows x 16 B per instruction - 32 lines for
    // the LSU and, when the destination is a peer GPU (sequence-parallel mode), 16-byte NVLink packets.  All MMAs
    // have completed (o_full), so the Q / K tiles are dead: each softmax warp stages its 32 x 128 block there.
    ptx::mbar_wait<kBackoff>(o_full, 0);
    ptx::tc_fence_after();
    const float inv = 1.f / l;
    constexpr int kPitch = HD * 2 + 16;  // 272 B: conflict-free row-wise writes and transposed reads
    uint8_t* stage = smem + (warp - 4) * (32 * kPitch);
#pragma unroll 1
    for (int ch = 0; ch < 4; ++ch) {
      float v[32];
      tmem_ld32(o_addr + ch * 32, v);
      uint4* dst = reinterpret_cast<uint4*>(stage + lane * kPitch + ch * 64);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        uint4 u;
        u.x = ptx::pack_bf16x2(v[8 * i + 0] * inv, v[8 * i + 1] * inv);
        u.y = ptx::pack_bf16x2(v[8 * i + 2] * inv, v[8 * i + 3] * inv);
        u.z = ptx::pack_bf16x2(v[8 * i + 4] * inv, v[8 * i + 5] * inv);
        u.w = ptx::pack_bf16x2(v[8 * i + 6] * inv, v[8 * i + 7] * inv);
        dst[i] = u;
      }
    }
    __syncwarp();
    const int row0 = q0 + t * BQ + quad * 32;  // first row of this warp
    const int rr = lane >> 4, cc = lane & 15;
#pragma unroll 4
    for (int it = 0; it < 16; ++it) {
      const int r = it * 2 + rr;
      const int grow = row0 + r;
      if (grow < P.S) {
        bf16* orow;
        if (P.sp_rows > 0) {
          const int dest = grow / P.sp_rows;
          orow = P.sp_out[dest] + (long long)b * P.out_bs + (long long)(grow - dest * P.sp_rows) * P.out_ld +
                 P.out_col0 + h * HD;
        } else {
          orow = P.out + (long long)b * P.out_bs + (long long)grow * P.out_ld + P.out_col0 + h * HD;
        }
        *reinterpret_cast<uint4*>(orow + cc * 8) = *reinterpret_cast<const uint4*>(stage + r * kPitch + cc * 16);
      }
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) ptx::tmem_dealloc<1>(tmem, 512);
}


#ifdef RT_AB_VARIANTS
#include "attn_decoupled_sm100.cuh"
#include "attn_rr_sm100.cuh"
#include "attn_variants_sm100.cuh"
#endif

}  // namespace

#ifdef RT_AB_VARIANTS
// Every kernel other than the product (A/B build only; tools/attn_sweep.py, tools/attn_trace.py).
static void launch_attention_variant(const AttnParams& P, const AttnArgs& a, cudaStream_t stream, int variant) {
  using KernelFn = void (*)(const AttnParams);
  // variant 0 = product configuration.  The rest are kept for A/B timing (tools/attn_sweep.py):
  //   1-3 timing experiments of the scalar form (wrong results); 4 scalar form, 5 scalar + 25 % polynomial;
  //   6.. packed form: (poly mask, split P) = (0,0) (0,1) (25%,0) (25%,1) (37.5%,1) (50%,1) (12.5%,1)
  static const KernelFn table[] = {
      attn_tc_kernel<0, 0, true, 0x88, true>,
      attn_tc_kernel<8, 0, true, 0x88, true>, attn_tc_kernel<16, 0, true, 0x88, true>,  // 1, 2: barrier polls with 20 / 100 ns sleeps
      attn_tc_kernel<3, 0>,
      attn_tc_kernel<0, 0>, attn_tc_kernel<0, 4>,
      attn_tc_kernel<0, 0, true, 0x00, false>, attn_tc_kernel<0, 0, true, 0x00, true>,
      attn_tc_kernel<0, 0, true, 0x88, false>, attn_tc_kernel<0, 0, true, 0x88, true>,
      attn_tc_kernel<0, 0, true, 0x92, true>, attn_tc_kernel<0, 0, true, 0xAA, true>,
      attn_tc_kernel<0, 0, true, 0x80, true>,
      attn_tc_kernel<4, 0>, attn_tc_kernel<7, 0>,  // 13: no TMEM read of S; 14: no read, no max, no exp2
      attn_tc_kernel<0, 0, true, 0x88, true, true>, attn_tc_kernel<0, 0, true, 0x88, false, true>,  // 15, 16: elected arrive
      attn_tc_kernel<0, 0, true, 0x00, false, true>,
      attn_tc_kernel<0, 0, true, 0x88, true, false, false, true>,   // 18: product + hand-off trace
      attn_tc_kernel<0, 0, true, 0x88, true, false, false, false, true>};  // 19: early first chunk
  constexpr int kNumVariants = sizeof(table) / sizeof(table[0]);
  // v4 kernels (64-key blocks, double-buffered S): variant 20 + i.  (poly mask, debug) =
  //   (25%,0) (0,0) (37.5%,0) (50%,0) (12.5%,0) | timing experiments: (0, no S read) (0, no exp2) (0, neither)
  static const KernelFn table4[] = {attn_tc_kernel_v4<0x88, 0>, attn_tc_kernel_v4<0x00, 0>, attn_tc_kernel_v4<0x92, 0>,
                                    attn_tc_kernel_v4<0xAA, 0>, attn_tc_kernel_v4<0x80, 0>, attn_tc_kernel_v4<0, 1>,
                                    attn_tc_kernel_v4<0, 2>,    attn_tc_kernel_v4<0, 3>};
  constexpr int kNumVariants4 = sizeof(table4) / sizeof(table4[0]);
  static PerDeviceOnce attr_set;
  if (attr_set.first()) {
    for (int i = 0; i < kNumVariants; ++i)
      RT_CHECK_CUDA(cudaFuncSetAttribute(table[i], cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    for (int i = 0; i < kNumVariants4; ++i)
      RT_CHECK_CUDA(cudaFuncSetAttribute(table4[i], cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes4));
  }
  // v5 kernels (one query tile per CTA, double-buffered 128-key score tile): variant 30 + i
  static const KernelFn table5[] = {attn_tc_kernel_v5<0x88, 0>, attn_tc_kernel_v5<0x00, 0>, attn_tc_kernel_v5<0x92, 0>,
                                    attn_tc_kernel_v5<0xAA, 0>, attn_tc_kernel_v5<0x80, 0>, attn_tc_kernel_v5<0, 1>,
                                    attn_tc_kernel_v5<0, 2>,    attn_tc_kernel_v5<0, 3>};
  constexpr int kNumVariants5 = sizeof(table5) / sizeof(table5[0]);
  static PerDeviceOnce attr5_set;
  if (attr5_set.first()) {
    for (int i = 0; i < kNumVariants5; ++i)
      RT_CHECK_CUDA(cudaFuncSetAttribute(table5[i], cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes5));
  }
  // v6 kernels (CTA pair, cta_group::2; one query tile per CTA, K / V halves split across the pair): variant 40 + i
  static const KernelFn table6[] = {attn_tc_kernel_v6<0x88, 0>, attn_tc_kernel_v6<0x00, 0>, attn_tc_kernel_v6<0x92, 0>,
                                    attn_tc_kernel_v6<0xAA, 0>, attn_tc_kernel_v6<0x80, 0>, attn_tc_kernel_v6<0, 1>,
                                    attn_tc_kernel_v6<0, 2>,    attn_tc_kernel_v6<0, 3>,
                                    attn_tc_kernel_v6<0x88, 0, true>, attn_tc_kernel_v6<0x00, 0, true>,  // 48, 49: two MMA issuers
                                    attn_tc_kernel_v6<0xAA, 0, true>, attn_tc_kernel_v6<0, 3, true>,
                                    attn_tc_kernel_v6<0x88, 0, false, 3>, attn_tc_kernel_v6<0x00, 0, false, 3>,  // 52.. three score buffers
                                    attn_tc_kernel_v6<0xAA, 0, false, 3>, attn_tc_kernel_v6<0, 3, false, 3>};
  constexpr int kNumVariants6 = sizeof(table6) / sizeof(table6[0]);
  static PerDeviceOnce attr6_set;
  if (attr6_set.first()) {
    for (int i = 0; i < kNumVariants6; ++i)
      RT_CHECK_CUDA(cudaFuncSetAttribute(table6[i], cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes6));
  }
  // half-row kernels (two softmax threads per query row, 640 threads): variant 60 + i; poly share 25 / 0 / 50 / 37.5 %
  static const KernelFn tableH[] = {attn_tc_kernel<0, 0, true, 0x88, true, false, true>,
                                    attn_tc_kernel<0, 0, true, 0x00, true, false, true>,
                                    attn_tc_kernel<0, 0, true, 0xAA, true, false, true>,
                                    attn_tc_kernel<0, 0, true, 0x92, true, false, true>,
                                    attn_tc_kernel<0, 0, true, 0x88, true, false, true, true>};  // 64: 60 + trace
  constexpr int kNumVariantsH = sizeof(tableH) / sizeof(tableH[0]);
  static PerDeviceOnce attrH_set;
  if (attrH_set.first()) {
    for (int i = 0; i < kNumVariantsH; ++i)
      RT_CHECK_CUDA(cudaFuncSetAttribute(tableH[i], cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytesHalfRow));
  }
  if (variant >= 110) {
    // the product's structure with the SOFTWARE-PIPELINED exponential loop (kPipe): 110 = 25 % polynomial, 111 = none,
    // 112 = 37.5 %, 113 = 50 %, 114 = 25 % + hand-off trace
    static const KernelFn tableP[] = {attn_tc_kernel<0, 0, true, 0x88, true, false, false, false, false, true>,
                                      attn_tc_kernel<0, 0, true, 0x00, true, false, false, false, false, true>,
                                      attn_tc_kernel<0, 0, true, 0x92, true, false, false, false, false, true>,
                                      attn_tc_kernel<0, 0, true, 0xAA, true, false, false, false, false, true>,
                                      attn_tc_kernel<0, 0, true, 0x88, true, false, false, true, false, true>};
    constexpr int kNumVariantsP = sizeof(tableP) / sizeof(tableP[0]);
    RT_REQUIRE(variant - 110 < kNumVariantsP, "attention: unknown variant");
    static PerDeviceOnce attrP_set;
    if (attrP_set.first()) {
      for (int i = 0; i < kNumVariantsP; ++i)
        RT_CHECK_CUDA(cudaFuncSetAttribute(tableP[i], cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)((long long)P.n_qpairs * a.heads * a.batch));
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = kSmemBytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = get_option("no_pdl") ? 0 : 1;
    RT_CHECK_CUDA(cudaLaunchKernelEx(&cfg, tableP[variant - 110], P));
    count_launch();
    return;
  }
  if (variant >= 90) {
    // round-robin decoupled kernel (one thread per row, two softmax warpgroups per query tile take the key blocks in
    // turn): variant 90 + i; polynomial share 25 / 0 / 37.5 / 50 %; 94: 25 % + hand-off trace
    // 95-99: the ONE-pass form of the same (block j exponentiated against block j - 1's reference), same order
    static const KernelFn table9[] = {attn_tc_kernel_v9<0x88>, attn_tc_kernel_v9<0x00>, attn_tc_kernel_v9<0x92>,
                                      attn_tc_kernel_v9<0xAA>, attn_tc_kernel_v9<0x88, true>,
                                      attn_tc_kernel_v9<0x88, false, true>, attn_tc_kernel_v9<0x00, false, true>,
                                      attn_tc_kernel_v9<0x92, false, true>, attn_tc_kernel_v9<0xAA, false, true>,
                                      attn_tc_kernel_v9<0x88, true, true>};
    constexpr int kNumVariants9 = sizeof(table9) / sizeof(table9[0]);
    RT_REQUIRE(variant - 90 < kNumVariants9, "attention: unknown variant");
    RT_REQUIRE(P.sp_txt == 0, "attention: variant 9x walks the keys in buffer order only");
    static PerDeviceOnce attr9_set;
    if (attr9_set.first()) {
      for (int i = 0; i < kNumVariants9; ++i)
        RT_CHECK_CUDA(cudaFuncSetAttribute(table9[i], cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes9));
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)((long long)P.n_qpairs * a.heads * a.batch));
    cfg.blockDim = dim3(kThreads9);
    cfg.dynamicSmemBytes = kSmemBytes9;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = get_option("no_pdl") ? 0 : 1;
    RT_CHECK_CUDA(cudaLaunchKernelEx(&cfg, table9[variant - 90], P));
    count_launch();
    return;
  }
  if (variant >= 80) {
    // decoupled kernel (P through shared memory, two threads per row): variant 80 + i; polynomial share 25 / 0 / 37.5 /
    // 50 %; 84: 25 % + hand-off trace
    static const KernelFn table8[] = {attn_tc_kernel_v8<0x88>, attn_tc_kernel_v8<0x00>, attn_tc_kernel_v8<0x92>,
                                      attn_tc_kernel_v8<0xAA>, attn_tc_kernel_v8<0x88, true>};
    constexpr int kNumVariants8 = sizeof(table8) / sizeof(table8[0]);
    RT_REQUIRE(variant - 80 < kNumVariants8, "attention: unknown variant");
    static PerDeviceOnce attr8_set;
    if (attr8_set.first()) {
      for (int i = 0; i < kNumVariants8; ++i)
        RT_CHECK_CUDA(cudaFuncSetAttribute(table8[i], cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes8));
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)((long long)P.n_qpairs * a.heads * a.batch));
    cfg.blockDim = dim3(kThreads8);
    cfg.dynamicSmemBytes = kSmemBytes8;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = get_option("no_pdl") ? 0 : 1;
    RT_CHECK_CUDA(cudaLaunchKernelEx(&cfg, table8[variant - 80], P));
    count_launch();
    return;
  }
  if (variant >= 70) {
    // pair kernel (CTA pair, cta_group::2, two query tiles per CTA): variant 70 + i; polynomial share 25 / 0 / 37.5 / 50 %
    static const KernelFn table7[] = {attn_tc_pair_kernel<0x88>, attn_tc_pair_kernel<0x00>, attn_tc_pair_kernel<0x92>,
                                      attn_tc_pair_kernel<0xAA>};
    constexpr int kNumVariants7 = sizeof(table7) / sizeof(table7[0]);
    RT_REQUIRE(variant - 70 < kNumVariants7, "attention: unknown variant");
    static PerDeviceOnce attr7_set;
    if (attr7_set.first()) {
      for (int i = 0; i < kNumVariants7; ++i)
        RT_CHECK_CUDA(cudaFuncSetAttribute(table7[i], cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes7));
    }
    const long long clusters = (long long)((a.S + 4 * BQ - 1) / (4 * BQ)) * a.heads * a.batch;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(2 * clusters));
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = kSmemBytes7;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = get_option("no_pdl") ? 1 : 2;
    RT_CHECK_CUDA(cudaLaunchKernelEx(&cfg, table7[variant - 70], P));
    count_launch();
    return;
  }
  if (variant >= 60) {
    RT_REQUIRE(variant - 60 < kNumVariantsH, "attention: unknown variant");
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)((long long)P.n_qpairs * a.heads * a.batch));
    cfg.blockDim = dim3(kThreadsHalfRow);
    cfg.dynamicSmemBytes = kSmemBytesHalfRow;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = get_option("no_pdl") ? 0 : 1;
    RT_CHECK_CUDA(cudaLaunchKernelEx(&cfg, tableH[variant - 60], P));
    count_launch();
    return;
  }
  if (variant >= 40) {
    RT_REQUIRE(variant - 40 < kNumVariants6, "attention: unknown variant");
    const long long pairs = (long long)P.n_qpairs * a.heads * a.batch;
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)(2 * pairs));
    cfg.blockDim = dim3(kThreads5);
    cfg.dynamicSmemBytes = kSmemBytes6;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    RT_CHECK_CUDA(cudaLaunchKernelEx(&cfg, table6[variant - 40], P));
    count_launch();
    return;
  }
  if (variant >= 30) {
    RT_REQUIRE(variant - 30 < kNumVariants5, "attention: unknown variant");
    const long long grid5 = (long long)((a.S + BQ - 1) / BQ) * a.heads * a.batch;
    table5[variant - 30]<<<(unsigned)grid5, kThreads5, kSmemBytes5, stream>>>(P);
    RT_POST_LAUNCH();
    return;
  }
  const long long grid = (long long)P.n_qpairs * a.heads * a.batch;
  if (variant >= 20) {
    RT_REQUIRE(variant - 20 < kNumVariants4, "attention: unknown variant");
    table4[variant - 20]<<<(unsigned)grid, kThreads, kSmemBytes4, stream>>>(P);
  } else {
    RT_REQUIRE(variant >= 0 && variant < kNumVariants, "attention: unknown variant");
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(kThreads);
    cfg.dynamicSmemBytes = kSmemBytes;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = get_option("no_pdl") ? 0 : 1;
    RT_CHECK_CUDA(cudaLaunchKernelEx(&cfg, table[variant], P));
    count_launch();
    return;
  }
  RT_POST_LAUNCH();
}
#endif

bool attention_tc_supported(const AttnArgs& a, std::string* why) {
  auto fail = [&](const char* m) { if (why) *why = m; return false; };
  if (a.dtype != RT_BF16) return fail("dtype is not bf16");
  if (a.hd != HD) return fail("head_dim is not 128");
  if (a.ld % 8 || a.batch_stride % 8 || a.q_col0 % 8 || a.k_col0 % 8 || a.v_col0 % 8) return fail("qkv alignment");
  if (a.out_ld % 8 || a.out_col0 % 8 || a.out_batch_stride % 8) return fail("out alignment");
  if (reinterpret_cast<uintptr_t>(a.qkv) & 15) return fail("pointer alignment");
  if (a.sp_rows > 0) {
    const int ndest = (a.S + a.sp_rows - 1) / a.sp_rows;
    if (ndest > RT_SP_MAX_RANKS) return fail("sp_rows: more destinations than RT_SP_MAX_RANKS");
    for (int d = 0; d < ndest; ++d)
      if (!a.sp_out[d] || (reinterpret_cast<uintptr_t>(a.sp_out[d]) & 15)) return fail("sp_out pointer missing or misaligned");
  } else if (!a.out || (reinterpret_cast<uintptr_t>(a.out) & 15)) {
    return fail("pointer alignment");
  }
  if (a.S < 1) return fail("empty sequence");
  return true;
}

void launch_attention_tc(const AttnArgs& a, cudaStream_t stream, int variant, const SpSyncParams* sync) {
  std::string why;
  if (!attention_tc_supported(a, &why)) throw Error(RT_ERR_UNSUPPORTED, "tcgen05 attention: " + why);
  if (a.batch == 0) return;
  AttnParams P{};
  if (sync) {
    RT_REQUIRE(variant == 0, "attention: in-kernel phase synchronisation exists in the product kernel only");
    P.sync = *sync;
  }
  const int cols = a.ld;  // the map spans whole rows of the projection buffer; q / k / v are column offsets
  uint64_t dims[3] = {(uint64_t)cols, (uint64_t)a.S, (uint64_t)a.batch};
  uint64_t strides[2] = {(uint64_t)a.ld * 2, (uint64_t)a.batch_stride * 2};
  uint32_t box[3] = {64, 128, 1};
  encode_tmap_bf16(&P.tm, a.qkv, 3, dims, strides, box);
  uint32_t box_half[3] = {64, 64, 1};
  encode_tmap_bf16(&P.tmh, a.qkv, 3, dims, strides, box_half);
  P.out = reinterpret_cast<bf16*>(a.out);
  P.out_bs = a.out_batch_stride; P.out_ld = a.out_ld; P.out_col0 = a.out_col0;
  P.q_col0 = a.q_col0; P.k_col0 = a.k_col0; P.v_col0 = a.v_col0;
  P.S = a.S; P.heads = a.heads;
  P.sp_rows = a.sp_rows > 0 ? a.sp_rows : 0;
  if (P.sp_rows > 0 && a.sp_txt_rows > 0 && a.sp_txt_rows < a.sp_rows && a.S % a.sp_rows == 0) {
    const int world = a.S / a.sp_rows, n_loc = a.sp_rows - a.sp_txt_rows;
    if (a.sp_txt_rows % 64 == 0 && (world * a.sp_txt_rows) % BKV == 0 && n_loc % BKV == 0) {
      P.sp_txt = a.sp_txt_rows;
      P.sp_txt_blocks = world * a.sp_txt_rows / BKV;
    }
  }
  for (int i = 0; i < RT_SP_MAX_RANKS; ++i) P.sp_out[i] = reinterpret_cast<bf16*>(a.sp_out[i]);
  P.n_qpairs = (a.S + 2 * BQ - 1) / (2 * BQ);
  P.scale_log2 = 1.4426950408889634f / sqrtf((float)a.hd);
  using KernelFn = void (*)(const AttnParams);
  if (variant != 0) {
#ifdef RT_AB_VARIANTS
    launch_attention_variant(P, a, stream, variant);
    return;
#else
    throw Error(RT_ERR_UNSUPPORTED,
                "attention: this library ships the product kernel only; the A/B variants live in the build made by "
                "`python -m reptext_b200.build --ab` (csrc/librt_reptext_ab.so, load it with RT_LIB=...)");
#endif
  }
  const KernelFn product = attn_tc_kernel<0, 0, true, 0x88, true>;
  static PerDeviceOnce attr_set;
  if (attr_set.first())
    RT_CHECK_CUDA(cudaFuncSetAttribute(product, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes));
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)((long long)P.n_qpairs * a.heads * a.batch));
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = kSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = get_option("no_pdl") ? 0 : 1;
  RT_CHECK_CUDA(cudaLaunchKernelEx(&cfg, product, P));
  count_launch();
}

}  // namespace rt

#ifdef RT_AB_VARIANTS
extern "C" __attribute__((visibility("default"))) int rt_debug_attn_trace(long long* host_out, int n) {
  return rt::guarded([&] {
    RT_REQUIRE(host_out && n > 0 && n <= 128 * 32, "debug_attn_trace: bad argument");
    RT_CHECK_CUDA(cudaMemcpyFromSymbol(host_out, rt::g_attn_trace, (size_t)n * sizeof(long long)));
  });
}
#endif
