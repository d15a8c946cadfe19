// Persistent, warp-specialised bf16 GEMM for sm_100a: TMA -> swizzled smem -> tcgen05.mma -> TMEM ->
// fused epilogue -> global.  One launch covers up to two "problems" (text rows and image rows of the
// joint sequence, each with its own weights) and up to four column segments per problem (q | k | v |
// mlp, each with its own weight matrix, bias, epilogue mode and destination), so that the QKV(+MLP)
// projections of a Flux block are ONE kernel without concatenated weight copies.
//
//   warp 0      TMA producer (one elected lane)            warp 2   TMEM allocator
//   warp 1      tcgen05.mma issuer (one lane)              warps 4-7 epilogue (TMEM lane quadrant = warp % 4)
//
// Pipelines: smem ring full/empty (TMA <-> MMA), TMEM accumulator double buffer full/empty
// (MMA <-> epilogue), static persistent tile scheduler (tile = blockIdx.x + i * gridDim.x).
// kCtaGroup == 2 pairs two CTAs of a cluster on a 256 x BN tile (tcgen05 cta_group::2): each CTA
// loads its own 128 rows of A and HALF of the W tile, the leader issues the MMAs for both.
//
// Epilogues (reference semantics: diffusers FluxTransformerBlock / FluxSingleTransformerBlock,
// reached from RepText/controlnet_flux.py:343-348, and controlnet_flux.py:385-396 for SCALE_MASK):
//   BIAS, GELU(tanh), QK-RMSNorm + RoPE per 128-wide head, gate * x + residual (+ ControlNet
//   residual), (x) * conditioning_scale * regional_mask (+ accumulate).
#include <mutex>
#include <unordered_map>
#include <vector>

#include "dtype_utils.cuh"
#include "ptx_sm100.cuh"
#include "rt_internal.h"
#include "sp_sync.cuh"

namespace rt {

namespace {

constexpr int BM = 128;       // rows per CTA tile (= TMEM lanes)
constexpr int BK = 64;        // k-block: 64 bf16 = 128 B = one swizzle row
constexpr int UMMA_K = 16;
constexpr int kThreads = 256;      // 4 epilogue warps; the 8-warp form (kEpiWarps = 8) launches 384
constexpr int kEpiThreads = 128;
constexpr int kMaxSeg = 4;

struct alignas(64) TcSegment {
  CUtensorMap tmW;  // 2D (K, rows), box (64, BN / cta_group)
  const bf16* bias;
  bf16* out;
  const bf16* norm_w;
  long long out_bs;
  int n_begin, n_end, mode, out_ld, out_col0, scatter, out_f32;
};

struct alignas(64) TcProblem {
  CUtensorMap tmA;  // 3D (K, rows, batch), box (64, 128, 1)
  TcSegment seg[kMaxSeg];
  const float* gate;
  const bf16* extra;
  const bf16* mask;
  long long e_bs;
  int a_row0, a_bcast, m_rows, out_row0, K, nseg, gate_ld, e_ld, e_row0, accumulate;
  float scale;
  int tiles_m, tiles_n, tile_base, num_tiles;
  // 3x3 convolution as an implicit GEMM (VAE): A is an NHWC image, tmA is 4-D (C, W, H, B) with a box of
  // (64 channels, conv_bw, 128 / conv_bw, 1) pixels; k-block kb = tap * conv_kcb + channel block, the tap shifts the
  // box by (dx, dy) and TMA's out-of-bounds zero fill is the padding.  conv_w == 0: plain GEMM.
  int conv_w, conv_bw, conv_kcb;
  // Tile order.  band == 0: row tiles fastest (a wave = every row tile x a few column tiles: right when A stays in L2 and
  // W is streamed once).  band > 0: the row tiles are cut into bands of `band` tiles and a band is swept over ALL column
  // tiles before the next one starts: a wave then reads `band` row blocks of A and all of W - the order for long-K problems whose A operand (141 MB for the single blocks' proj_out) does not survive in L2
  // from one wave to the next and would otherwise be re-read once per wave.
  int band;
};

struct alignas(64) TcParams {
  TcProblem prob[2];
  const float2* rope;
  int nprob, batch, total_tiles, head_dim;
  // sequence-parallel scatter (rt_gemm_segment::scatter): destination buffers, columns per destination, row offset
  bf16* sp_out[RT_SP_MAX_RANKS];
  int sp_cols, sp_row0;
  SpSyncParams sync;  // sequence-parallel phase synchronisation inside this kernel (sp_sync.cuh); world == 0: none
  int bn;     // kDynN kernels: the tile width of THIS launch (a multiple of 32, <= 256); the last column tile may be partial
  int debug;  // option "gemm_debug": 1 = no epilogue, 2 = every k-block loads k = 0 (timing experiments, wrong
              // results); 4 = direct row-per-thread epilogue stores instead of the staged ones, 8 = L2 eviction
              // hints on the TMA loads (A/B, same results)
};

template <int BN, int kCtaGroup, int kEpiWarps = 4>
struct Cfg {
  static constexpr int kBRows = BN / kCtaGroup;             // W rows each CTA loads
  static constexpr int kABytes = BM * BK * 2;               // 16 KB
  static constexpr int kBBytes = kBRows * BK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  // 192 KB of ring with four epilogue warps; eight warps need twice the staging tiles: 176 KB of ring
  static constexpr int kRingBytes = (kEpiWarps == 8 ? 176 : 192) * 1024;
  static constexpr int kStages = kRingBytes / kStageBytes > 8 ? 8 : kRingBytes / kStageBytes;
  static constexpr int kAccStages = 2;
  static constexpr int kTmemCols = (kAccStages * BN <= 32) ? 32 : (kAccStages * BN <= 64) ? 64
                                   : (kAccStages * BN <= 128) ? 128 : (kAccStages * BN <= 256) ? 256 : 512;
  static constexpr int kBarBytes = ((2 * kStages + 2 * kAccStages) * 8 + 16 + 127) / 128 * 128;
  static constexpr int kEpiStageOff = kStages * kStageBytes + kBarBytes;  // per-warp epilogue staging tiles
  static constexpr int kSmemBytes = kEpiStageOff + kEpiWarps * 32 * 144 + 1024;  // + alignment slack
};

struct TileCoord {
  int p, b, m0, n0, seg;
};

__device__ __forceinline__ TileCoord decode_tile(const TcParams& P, int t, int bn, int rows_per_tile) {
  TileCoord c;
  c.p = (P.nprob > 1 && t >= P.prob[1].tile_base) ? 1 : 0;
  const TcProblem& pr = P.prob[c.p];
  int tl = t - pr.tile_base;
  int mt, nt;
  if (pr.band > 0) {
    const int per_batch = pr.tiles_m * pr.tiles_n;
    c.b = tl / per_batch;
    const int r = tl - c.b * per_batch;
    // column tiles fastest inside a band: the few tiles of the NEXT band that a wave picks up early (a band is rarely
    // exactly one wave) then share one row block of A instead of touching several
    const int full = pr.band * pr.tiles_n;             // tiles of a full band
    const int bi = r / full, rem = r - bi * full;
    const int mi = rem / pr.tiles_n;
    nt = rem - mi * pr.tiles_n;
    mt = bi * pr.band + mi;
  } else {
    mt = tl % pr.tiles_m;
    const int rest = tl / pr.tiles_m;
    c.b = rest % P.batch;
    nt = rest / P.batch;
  }
  c.m0 = mt * rows_per_tile;
  c.n0 = nt * bn;
  c.seg = 0;
#pragma unroll
  for (int s = 1; s < kMaxSeg; ++s)
    if (s < pr.nseg && c.n0 >= pr.seg[s].n_begin) c.seg = s;
  return c;
}

// ---- epilogue helpers: 32 consecutive columns of one row ------------------------------------------
__device__ __forceinline__ void load_bf16x32(const bf16* p, float (&v)[32]) {
  const uint4* q = reinterpret_cast<const uint4*>(p);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    uint4 t = __ldg(q + i);
    const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      v[i * 8 + 2 * j] = ptx::bf16_lo(w[j]);
      v[i * 8 + 2 * j + 1] = ptx::bf16_hi(w[j]);
    }
  }
}
__device__ __forceinline__ void load_bf16x32_rw(const bf16* p, float (&v)[32]) {  // data this kernel also writes
  const uint4* q = reinterpret_cast<const uint4*>(p);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    uint4 t = q[i];
    const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      v[i * 8 + 2 * j] = ptx::bf16_lo(w[j]);
      v[i * 8 + 2 * j + 1] = ptx::bf16_hi(w[j]);
    }
  }
}
__device__ __forceinline__ void store_bf16x32(bf16* p, const float (&v)[32]) {
  uint4* q = reinterpret_cast<uint4*>(p);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    uint4 t;
    t.x = ptx::pack_bf16x2(v[i * 8 + 0], v[i * 8 + 1]);
    t.y = ptx::pack_bf16x2(v[i * 8 + 2], v[i * 8 + 3]);
    t.z = ptx::pack_bf16x2(v[i * 8 + 4], v[i * 8 + 5]);
    t.w = ptx::pack_bf16x2(v[i * 8 + 6], v[i * 8 + 7]);
    q[i] = t;
  }
}
__device__ __forceinline__ void tmem_load_f32x32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  ptx::tmem_ld_32x32b_x32(taddr, r);
  ptx::tmem_ld_wait();
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ float gelu_tanh_fast(float x) {
  const float k0 = 0.7978845608028654f, k1 = 0.044715f;
  float inner = k0 * (x + k1 * x * x * x);
  return 0.5f * x * (1.0f + ptx::tanh_approx(inner));
}

// part / nparts: the eight-warp epilogue gives every TMEM lane quadrant TWO warps, each takes half of the tile's columns
// (whole heads for the per-head mode, whole 64-column store groups otherwise)
template <int BN>
__device__ __forceinline__ void epilogue_tile(const TcParams& P, const TcProblem& pr, const TcSegment& sg,
                                              uint32_t tacc, int b, int m, int n0, int part = 0, int nparts = 1) {
  constexpr int kChunks = BN / 32, kHeads = BN / 128;
  const int c_per = (kChunks + nparts - 1) / nparts, c_begin = part * c_per, c_end = min(kChunks, c_begin + c_per);
  const int h_per = (kHeads + nparts - 1) / nparts, h_begin = part * h_per, h_end = min(kHeads, h_begin + h_per);
  const bool row_ok = m < pr.m_rows;
  const int nl0 = n0 - sg.n_begin;  // column within the segment
  bf16* orow = sg.out + (long long)b * sg.out_bs + (long long)(pr.out_row0 + m) * sg.out_ld + sg.out_col0 + nl0;
  const bf16* bias = sg.bias ? sg.bias + nl0 : nullptr;
  // Sequence-parallel scatter: the 32-column chunk starting at segment column `col` belongs to rank col / sp_cols
  // and is stored straight into that rank's buffer (a peer-mapped pointer: the store crosses NVLink).
  auto out_chunk = [&](int col) -> bf16* {
    if (!sg.scatter) return orow + (col - nl0);
    const int dest = col / P.sp_cols;
    return P.sp_out[dest] + (long long)b * sg.out_bs + (long long)(P.sp_row0 + pr.out_row0 + m) * sg.out_ld +
           sg.out_col0 + (col - dest * P.sp_cols);
  };

  if (sg.out_f32) {
    // fp32 destination (rt_gemm_segment::out_f32): (acc + bias) [* scale], no bf16 rounding; a thread owns one row and
    // writes it as 16-byte stores (128 contiguous bytes per 32 columns)
    float* frow = reinterpret_cast<float*>(sg.out) + (long long)b * sg.out_bs + (long long)(pr.out_row0 + m) * sg.out_ld +
                  sg.out_col0 + nl0;
    const float sc = sg.mode == EPI_SCALE_MASK ? pr.scale : 1.f;
#pragma unroll 1
    for (int c = c_begin; c < c_end; ++c) {
      float v[32];
      tmem_load_f32x32(tacc + c * 32, v);
      if (bias) {
        float bv[32];
        load_bf16x32(bias + c * 32, bv);
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] += bv[i];
      }
      if (row_ok) {
        float4* dst = reinterpret_cast<float4*>(frow + c * 32);
#pragma unroll
        for (int i = 0; i < 8; ++i) dst[i] = make_float4(v[4 * i] * sc, v[4 * i + 1] * sc, v[4 * i + 2] * sc, v[4 * i + 3] * sc);
      }
    }
    return;
  }
  if (sg.mode == EPI_QKNORM_ROPE) {
    // one head = 128 columns; two passes over TMEM (reads are cheap) instead of 128 live registers
    const float2* rp = P.rope ? P.rope + (long long)(pr.out_row0 + m) * 64 : nullptr;
#pragma unroll 1
    for (int hc = h_begin; hc < h_end; ++hc) {
      bf16* ohead = out_chunk(nl0 + hc * 128);
      float ss = 0.f;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        float v[32];
        tmem_load_f32x32(tacc + hc * 128 + c * 32, v);
        if (bias) {
          float bv[32];
          load_bf16x32(bias + hc * 128 + c * 32, bv);
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] += bv[i];
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) ss += v[i] * v[i];
      }
      const float rs = rsqrtf(ss * (1.f / 128.f) + 1e-6f);
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        float v[32];
        tmem_load_f32x32(tacc + hc * 128 + c * 32, v);
        if (bias) {
          float bv[32];
          load_bf16x32(bias + hc * 128 + c * 32, bv);
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] += bv[i];
        }
        float wv[32];
        load_bf16x32(sg.norm_w + c * 32, wv);
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = v[i] * rs * wv[i];
        if (rp && row_ok) {
          const float4* r4 = reinterpret_cast<const float4*>(rp + c * 16);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float4 cs = __ldg(r4 + i);  // (cos0, sin0, cos1, sin1)
            float x0 = v[4 * i], x1 = v[4 * i + 1], x2 = v[4 * i + 2], x3 = v[4 * i + 3];
            v[4 * i] = x0 * cs.x - x1 * cs.y;
            v[4 * i + 1] = x1 * cs.x + x0 * cs.y;
            v[4 * i + 2] = x2 * cs.z - x3 * cs.w;
            v[4 * i + 3] = x3 * cs.z + x2 * cs.w;
          }
        }
        if (row_ok) store_bf16x32(ohead + c * 32, v);
      }
    }
    return;
  }

  const float* gate = (sg.mode == EPI_GATE_RESID && pr.gate) ? pr.gate + (long long)b * pr.gate_ld + n0 : nullptr;
  const bf16* extra = nullptr;
  if (sg.mode == EPI_GATE_RESID && pr.extra && m >= pr.e_row0)
    extra = pr.extra + (long long)b * pr.e_bs + (long long)(m - pr.e_row0) * pr.e_ld + n0;
  float mk = 1.f;
  if (sg.mode == EPI_SCALE_MASK) {
    mk = pr.scale;
    if (pr.mask && row_ok) mk *= __bfloat162float(pr.mask[m]);
  }
#pragma unroll 1
  for (int c = c_begin; c < c_end; ++c) {
    float v[32];
    tmem_load_f32x32(tacc + c * 32, v);
    if (bias) {
      float bv[32];
      load_bf16x32(bias + c * 32, bv);
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] += bv[i];
    }
    if (sg.mode == EPI_GELU) {
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = gelu_tanh_fast(v[i]);
    } else if (sg.mode == EPI_GATE_RESID) {
      if (gate) {
        const float4* g4 = reinterpret_cast<const float4*>(gate + c * 32);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float4 g = __ldg(g4 + i);
          v[4 * i] *= g.x; v[4 * i + 1] *= g.y; v[4 * i + 2] *= g.z; v[4 * i + 3] *= g.w;
        }
      }
      if (row_ok) {
        float r[32];
        load_bf16x32_rw(orow + c * 32, r);
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] += r[i];
        if (extra) {
          load_bf16x32(extra + c * 32, r);
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] += r[i];
        }
      }
    } else if (sg.mode == EPI_SCALE_MASK) {
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] *= mk;
      if (pr.accumulate && row_ok) {
        float r[32];
        load_bf16x32_rw(orow + c * 32, r);
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] += r[i];
      }
    }
    if (row_ok) store_bf16x32(sg.mode == EPI_BIAS ? out_chunk(nl0 + c * 32) : orow + c * 32, v);
  }
}


// ---- staged epilogue ------------------------------------------------------------------------------------
// TMEM hands every thread ONE ROW of the accumulator, so a direct store writes 32 rows x 16 B per warp instruction:
// 32 separate lines for the LSU and - when the destination is a peer GPU's memory (sequence-parallel scatter), where
// no L2 merges partial lines - 16-byte NVLink packets.  Instead each epilogue warp owns a 32-row x 64-column bf16
// staging tile in shared memory (row pitch 144 B: conflict-free for the row-wise writes and for the transposed
// reads), fills it row-wise and copies it out 4 rows x 128 B per instruction.  Residual operands (in-place
// gate * x + residual, accumulate) take the same road in the other direction first; values are still rounded once.
constexpr int kStagePitch = 144;
constexpr int kStageWarpBytes = 32 * kStagePitch;

__device__ __forceinline__ void stage_put32(uint8_t* stage, int lane, int half, const float (&v)[32]) {
  uint4* q = reinterpret_cast<uint4*>(stage + lane * kStagePitch + half * 64);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    uint4 t;
    t.x = ptx::pack_bf16x2(v[i * 8 + 0], v[i * 8 + 1]);
    t.y = ptx::pack_bf16x2(v[i * 8 + 2], v[i * 8 + 3]);
    t.z = ptx::pack_bf16x2(v[i * 8 + 4], v[i * 8 + 5]);
    t.w = ptx::pack_bf16x2(v[i * 8 + 6], v[i * 8 + 7]);
    q[i] = t;
  }
}
__device__ __forceinline__ void stage_get32(const uint8_t* stage, int lane, int half, float (&v)[32]) {
  const uint4* q = reinterpret_cast<const uint4*>(stage + lane * kStagePitch + half * 64);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint4 t = q[i];
    const uint32_t w[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      v[i * 8 + 2 * j] = ptx::bf16_lo(w[j]);
      v[i * 8 + 2 * j + 1] = ptx::bf16_hi(w[j]);
    }
  }
}
// g: (first row of this warp, first column of the 64-column group); 8 passes of 4 rows x 128 B
// chunks: 16-byte chunks of the 64-column group that exist (8, or 4 when a kDynN tile ends on half a group)
__device__ __forceinline__ void stage_store(const uint8_t* stage, int lane, bf16* g, long long ld, int rows_valid,
                                            int chunks = 8) {
  const int rr = lane >> 3, ch = lane & 7;
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int r = it * 4 + rr;
    if (r < rows_valid && ch < chunks)
      *reinterpret_cast<uint4*>(g + (long long)r * ld + ch * 8) =
          *reinterpret_cast<const uint4*>(stage + r * kStagePitch + ch * 16);
  }
}
__device__ __forceinline__ void stage_load(uint8_t* stage, int lane, const bf16* g, long long ld, int rows_valid,
                                           int chunks = 8) {
  const int rr = lane >> 3, ch = lane & 7;
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int r = it * 4 + rr;
    if (r < rows_valid && ch < chunks)
      *reinterpret_cast<uint4*>(stage + r * kStagePitch + ch * 16) =
          *reinterpret_cast<const uint4*>(g + (long long)r * ld + ch * 8);
  }
}

// the same 4 rows x 128 B pattern, global -> registers now and registers -> staging tile later: the next group's residual
// travels while the current group is being stored
__device__ __forceinline__ void stage_fetch(uint4 (&r)[8], int lane, const bf16* g, long long ld, int rows_valid, int chunks) {
  const int rr = lane >> 3, ch = lane & 7;
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int row = it * 4 + rr;
    if (row < rows_valid && ch < chunks) r[it] = *reinterpret_cast<const uint4*>(g + (long long)row * ld + ch * 8);
  }
}
__device__ __forceinline__ void stage_commit(uint8_t* stage, int lane, const uint4 (&r)[8], int rows_valid, int chunks) {
  const int rr = lane >> 3, ch = lane & 7;
#pragma unroll
  for (int it = 0; it < 8; ++it) {
    const int row = it * 4 + rr;
    if (row < rows_valid && ch < chunks) *reinterpret_cast<uint4*>(stage + row * kStagePitch + ch * 16) = r[it];
  }
}
__device__ __forceinline__ void prefetch_l1(const void* p) {
  asm volatile("prefetch.global.L1 [%0];" ::"l"(p));
}

// What a tile's epilogue needs besides the accumulator - the residual rows of its first 64-column group, the bias and
// gate lines of its columns - does not depend on the tile's MMAs: the epilogue warps fetch it BEFORE they wait for the
// accumulator (staging tile / L1), so that the epilogue of a CTA's last tile, which nothing hides, starts with its
// operands on the SM instead of with an L2 round trip per step of a dependent chain.  Returns true when the first
// group's residual is staged.  Same values, same arithmetic as the unprefetched form.
template <int BN, bool kDynN>
__device__ __forceinline__ bool epilogue_prefetch(const TcParams& P, const TcProblem& pr, const TcSegment& sg, int b, int m0w,
                                                  int lane, int n0, uint8_t* stage, int ncols, int part, int nparts) {
  const int rows_valid = min(max(pr.m_rows - m0w, 0), 32);
  if (rows_valid == 0 || sg.mode == EPI_QKNORM_ROPE || sg.out_f32 || sg.scatter) return false;  // warp-uniform
  const int n_groups = kDynN ? (ncols + 63) / 64 : BN / 64;
  const int g_per = (n_groups + nparts - 1) / nparts, g_begin = part * g_per, g_end = min(n_groups, g_begin + g_per);
  if (g_begin >= g_end) return false;
  const int nl0 = n0 - sg.n_begin;
  const int c0 = g_begin * 64, c1 = min(kDynN ? ncols : BN, g_end * 64);  // this warp's columns of the tile
  if (sg.bias && c0 + lane * 64 < c1) prefetch_l1(sg.bias + nl0 + c0 + lane * 64);
  if (sg.mode == EPI_GATE_RESID && pr.gate && c0 + lane * 32 < c1)
    prefetch_l1(pr.gate + (long long)b * pr.gate_ld + n0 + c0 + lane * 32);
  const bool resid = sg.mode == EPI_GATE_RESID || (sg.mode == EPI_SCALE_MASK && pr.accumulate);
  if (!resid) return false;
  const bf16* gp = sg.out + (long long)b * sg.out_bs + (long long)(pr.out_row0 + m0w) * sg.out_ld + sg.out_col0 + nl0 + c0;
  const int chunks = kDynN ? min(2, (ncols - c0) / 32) * 4 : 8;
  stage_load(stage, lane, gp, sg.out_ld, rows_valid, chunks);
  __syncwarp();
  return true;
}

// kDynN: the tile is `ncols` (a multiple of 32) columns wide instead of BN - the launch's run-time tile width, or what
// is left of the segment in its last column tile (single-segment BIAS / GELU / GATE_RESID / SCALE_MASK launches only).
template <int BN, bool kDynN = false>
__device__ __forceinline__ void epilogue_tile_staged(const TcParams& P, const TcProblem& pr, const TcSegment& sg,
                                                     uint32_t tacc, int b, int m0w, int lane, int n0, uint8_t* stage,
                                                     int ncols = BN, int part = 0, int nparts = 1,
                                                     bool first_staged = false) {
  const int m = m0w + lane;
  const bool row_ok = m < pr.m_rows;
  const int rows_valid = min(max(pr.m_rows - m0w, 0), 32);
  if (rows_valid == 0) return;  // warp-uniform
  const int nl0 = n0 - sg.n_begin;  // column within the segment
  const bf16* bias = sg.bias ? sg.bias + nl0 : nullptr;
  // (first row of this warp, segment column `col`); with scatter the 64-column group starting at `col` lives in
  // rank (col / sp_cols)'s buffer - a peer-mapped pointer, the store crosses NVLink
  auto gptr = [&](int col) -> bf16* {
    const long long row = pr.out_row0 + m0w;
    if (!sg.scatter) return sg.out + (long long)b * sg.out_bs + row * sg.out_ld + sg.out_col0 + col;
    const int dest = col / P.sp_cols;
    return P.sp_out[dest] + (long long)b * sg.out_bs + (P.sp_row0 + row) * sg.out_ld + sg.out_col0 +
           (col - dest * P.sp_cols);
  };

  if (sg.mode == EPI_QKNORM_ROPE) {
    // one head = 128 columns; two passes over TMEM (reads are cheap) instead of 128 live registers
    const float2* rp = P.rope ? P.rope + (long long)(pr.out_row0 + m) * 64 : nullptr;
    constexpr int kHeads = BN / 128;
    const int h_per = (kHeads + nparts - 1) / nparts, h_begin = part * h_per, h_end = min(kHeads, h_begin + h_per);
#pragma unroll 1
    for (int hc = h_begin; hc < h_end; ++hc) {
      float ss = 0.f;
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        float v[32];
        tmem_load_f32x32(tacc + hc * 128 + c * 32, v);
        if (bias) {
          float bv[32];
          load_bf16x32(bias + hc * 128 + c * 32, bv);
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] += bv[i];
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) ss += v[i] * v[i];
      }
      const float rs = rsqrtf(ss * (1.f / 128.f) + 1e-6f);
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        float v[32];
        tmem_load_f32x32(tacc + hc * 128 + c * 32, v);
        if (bias) {
          float bv[32];
          load_bf16x32(bias + hc * 128 + c * 32, bv);
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] += bv[i];
        }
        float wv[32];
        load_bf16x32(sg.norm_w + c * 32, wv);
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = v[i] * rs * wv[i];
        if (rp && row_ok) {
          const float4* r4 = reinterpret_cast<const float4*>(rp + c * 16);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float4 cs = __ldg(r4 + i);  // (cos0, sin0, cos1, sin1)
            float x0 = v[4 * i], x1 = v[4 * i + 1], x2 = v[4 * i + 2], x3 = v[4 * i + 3];
            v[4 * i] = x0 * cs.x - x1 * cs.y;
            v[4 * i + 1] = x1 * cs.x + x0 * cs.y;
            v[4 * i + 2] = x2 * cs.z - x3 * cs.w;
            v[4 * i + 3] = x3 * cs.z + x2 * cs.w;
          }
        }
        stage_put32(stage, lane, c & 1, v);
        if (c & 1) {
          __syncwarp();
          stage_store(stage, lane, gptr(nl0 + hc * 128 + (c >> 1) * 64), sg.out_ld, rows_valid);
          __syncwarp();
        }
      }
    }
    return;
  }

  const float* gate = (sg.mode == EPI_GATE_RESID && pr.gate) ? pr.gate + (long long)b * pr.gate_ld + n0 : nullptr;
  const bf16* extra = nullptr;
  if (sg.mode == EPI_GATE_RESID && pr.extra && m >= pr.e_row0 && row_ok)
    extra = pr.extra + (long long)b * pr.e_bs + (long long)(m - pr.e_row0) * pr.e_ld + n0;
  float mk = 1.f;
  if (sg.mode == EPI_SCALE_MASK) {
    mk = pr.scale;
    if (pr.mask && row_ok) mk *= __bfloat162float(pr.mask[m]);
  }
  const bool resid = sg.mode == EPI_GATE_RESID || (sg.mode == EPI_SCALE_MASK && pr.accumulate);
  const int n_groups = kDynN ? (ncols + 63) / 64 : BN / 64;
  const int g_per = (n_groups + nparts - 1) / nparts, g_begin = part * g_per, g_end = min(n_groups, g_begin + g_per);
  bool staged = first_staged;  // this group's residual is already in the staging tile (epilogue_prefetch / the previous group)
#pragma unroll 1
  for (int g = g_begin; g < g_end; ++g) {
    bf16* gp = gptr(nl0 + g * 64);
    const int n_half = kDynN ? min(2, (ncols - g * 64) / 32) : 2;
    const int chunks = kDynN ? n_half * 4 : 8;
    if (resid && !staged) {
      stage_load(stage, lane, gp, sg.out_ld, rows_valid, chunks);
      __syncwarp();
    }
#pragma unroll 1
    for (int half = 0; half < n_half; ++half) {
      const int c = g * 2 + half;
      float v[32];
      tmem_load_f32x32(tacc + c * 32, v);
      if (bias) {
        float bv[32];
        load_bf16x32(bias + c * 32, bv);
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] += bv[i];
      }
      if (sg.mode == EPI_GELU) {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = gelu_tanh_fast(v[i]);
      } else if (sg.mode == EPI_GATE_RESID) {
        if (gate) {
          const float4* g4 = reinterpret_cast<const float4*>(gate + c * 32);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float4 gv = __ldg(g4 + i);
            v[4 * i] *= gv.x; v[4 * i + 1] *= gv.y; v[4 * i + 2] *= gv.z; v[4 * i + 3] *= gv.w;
          }
        }
      } else if (sg.mode == EPI_SCALE_MASK) {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] *= mk;
      }
      if (resid) {
        float r[32];
        stage_get32(stage, lane, half, r);
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] += r[i];
        if (extra) {
          load_bf16x32(extra + c * 32, r);
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] += r[i];
        }
      }
      stage_put32(stage, lane, half, v);
    }
    __syncwarp();
    // the next group's residual leaves global memory now and lands in the staging tile once this group has left it
    const bool next = resid && g + 1 < g_end;
    uint4 nxt[8];
    const int n_chunks = kDynN ? min(2, (ncols - (g + 1) * 64) / 32) * 4 : 8;
    if (next) stage_fetch(nxt, lane, gptr(nl0 + (g + 1) * 64), sg.out_ld, rows_valid, n_chunks);
    stage_store(stage, lane, gp, sg.out_ld, rows_valid, chunks);
    __syncwarp();
    staged = false;
    if (next) {
      stage_commit(stage, lane, nxt, rows_valid, n_chunks);
      __syncwarp();
      staged = true;
    }
  }
}

// -------------------------------------------------------------------------------------------------
// kDynN (instantiated for BN = 256 only): the shared-memory ring and the accumulator stride keep the BN = 256 layout,
// but a tile is P.bn <= 256 columns wide (TMA box, expected bytes, instruction descriptor and epilogue follow P.bn).
// The host takes this form when a narrower tile fills the last wave better (launch_gemm_tc: sequence-parallel shards
// of 1216 rows x N = 3072 are 120 tiles of 256 columns for 148 SMs, but 140 tiles of 224).  An output element's
// arithmetic does not depend on the tile width: results are bit-identical to the BN = 256 kernel.
// kEpiWarps = 8: two epilogue warps per TMEM lane quadrant (warps 4-7 take the first half of a tile's columns, warps
// 8-11 the second).  The epilogue of a CTA's LAST tile is not hidden behind the next tile's MMAs: with four warps it is
// 3-6 % of the large launches' time and 12 % of (4608, 3072, 3072) (option gemm_debug = 1 removes it).
template <int BN, int kCtaGroup, bool kDynN = false, int kEpiWarps = 4>
__global__ void __launch_bounds__(128 + 32 * kEpiWarps, 1) gemm_tc_kernel(const __grid_constant__ TcParams P) {
  using C = Cfg<BN, kCtaGroup, kEpiWarps>;
  static_assert(!kDynN || BN == 256, "the run-time tile width lives in the BN = 256 layout");
  const int bn = kDynN ? P.bn : BN;
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B tiles need 1024-byte alignment
  const uint32_t raw_u32 = ptx::smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_u32 + 1023u) & ~1023u) - raw_u32);
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + C::kStages * C::kABytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::kStages * C::kStageBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + C::kStages;
  uint64_t* tfull_bar = bars + 2 * C::kStages;
  uint64_t* tempty_bar = bars + 2 * C::kStages + C::kAccStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * C::kStages + 2 * C::kAccStages);

  // warp / CTA-rank values go through a shuffle so that the compiler can PROVE them warp-uniform: the producer and
  // MMA-issuer loops below are then compiled onto the uniform datapath (descriptors in uniform registers) instead of
  // per-instruction R2UR + waterfall loops (115 SASS instructions per k-block for 4 MMAs, measured with ncu: the
  // issuing thread was busy 82 % of the time and the tensor pipe starved behind it).
  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = (kCtaGroup == 2) ? (uint32_t)__shfl_sync(0xffffffffu, (int)ptx::cluster_ctarank(), 0) : 0u;
  const bool leader = cta_rank == 0;
  const int cluster_id = blockIdx.x / kCtaGroup;
  const int num_clusters = gridDim.x / kCtaGroup;
  constexpr int kRowsPerTile = BM * kCtaGroup;

  if (warp == 0 && lane == 0) {
    for (int p = 0; p < P.nprob; ++p) {
      ptx::prefetch_tmap(&P.prob[p].tmA);
      for (int s = 0; s < P.prob[p].nseg; ++s) ptx::prefetch_tmap(&P.prob[p].seg[s].tmW);
    }
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < C::kStages; ++i) {
      // ONE arrival: the leader's arrive.expect_tx covering the bytes of BOTH CTAs.  The peer's TMA loads
      // complete_tx directly on the leader's barrier (.cta_group::2), so the peer never arrives: a remote
      // mbarrier.arrive.release.cluster costs MEMBAR.ALL.GPU + ERRBAR per stage and serialises its TMA queue.
      ptx::mbar_init(&full_bar[i], 1);
      ptx::mbar_init(&empty_bar[i], 1);         // one tcgen05.commit
    }
    for (int i = 0; i < C::kAccStages; ++i) {
      ptx::mbar_init(&tfull_bar[i], 1);                         // one tcgen05.commit
      ptx::mbar_init(&tempty_bar[i], kCtaGroup * 32 * kEpiWarps);  // every epilogue thread of the pair
    }
    ptx::fence_barrier_init();
  }
  if (warp == 2) ptx::tmem_alloc<kCtaGroup>(tmem_slot, C::kTmemCols);
  ptx::tc_fence_before();
  if constexpr (kCtaGroup == 2) ptx::cluster_sync(); else __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // everything above is independent of the previous kernel's output (programmatic dependent launch)
  ptx::grid_launch_dependents();
  ptx::grid_dependency_wait();

  if (warp == 0) {
    // ===================== TMA producer (whole warp runs the loop, one elected lane issues) =====================
    // L2 eviction hints (activations evict_last, weights evict_first) are OFF by default: measured with ncu over the
    // 188 GEMM launches of a step they RAISE DRAM reads from 52 GB to 74 GB - a weight tile is shared by the 16-18
    // row tiles of its wave, and evict_first drops it before the neighbours have fetched it.
    // (bits 32 / 64: the opposite polarity for the long-K problems, whose W is re-read once per band of row tiles -
    //  W evict_last with A normal / A evict_first; A/B, same results)
    const uint64_t pol_a = (P.debug & 8) ? ptx::kL2EvictLast : (P.debug & 64) ? ptx::kL2EvictFirst : ptx::kL2EvictNormal;
    const uint64_t pol_w = (P.debug & 8) ? ptx::kL2EvictFirst : (P.debug & (32 | 64)) ? ptx::kL2EvictLast : ptx::kL2EvictNormal;
    // sequence-parallel: A holds rows the peers' attention epilogues stored - the phase barrier runs here (sp_sync.cuh)
    spsync::sp_barrier_head(P.sync, lane, blockIdx.x == 0);
    int stage = 0, phase = 0;
    // bytes one CTA's two loads of a stage deliver (out-of-bounds box rows are zero-filled AND counted)
    const uint32_t stage_tx = kDynN ? (uint32_t)(C::kABytes + (bn / kCtaGroup) * BK * 2) : (uint32_t)C::kStageBytes;
    for (int t = cluster_id; t < P.total_tiles; t += num_clusters) {
      const TileCoord tc = decode_tile(P, t, bn, kRowsPerTile);
      const TcProblem& pr = P.prob[tc.p];
      const TcSegment& sg = pr.seg[tc.seg];
      const int a_row = pr.a_row0 + tc.m0 + (int)cta_rank * BM;
      const int a_b = pr.a_bcast ? 0 : tc.b;
      const int w_row = (tc.n0 - sg.n_begin) + (int)cta_rank * (kDynN ? bn / kCtaGroup : C::kBRows);
      const int nkb = (pr.K + BK - 1) / BK;
      for (int kb = 0; kb < nkb; ++kb) {
        ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
        void* sa = smem_a + stage * C::kABytes;
        void* sb = smem_b + stage * C::kBBytes;
        const int k0 = (P.debug & 2) ? 0 : kb * BK;
        if (ptx::elect_one()) {
          if (pr.conv_w > 0) {
            // implicit 3x3 convolution: this CTA's 128 pixels start at (x, y) of image a_b
            const int m_cta = tc.m0 + (int)cta_rank * BM;
            const int y = m_cta / pr.conv_w, x = m_cta - y * pr.conv_w;
            const int tap = kb / pr.conv_kcb, cb = kb - tap * pr.conv_kcb;
            const int dy = tap / 3 - 1, dx = tap - (tap / 3) * 3 - 1;
            if constexpr (kCtaGroup == 1) {
              ptx::mbar_arrive_expect_tx(&full_bar[stage], stage_tx);
              ptx::tma_load_4d(&pr.tmA, &full_bar[stage], sa, cb * BK, x + dx, y + dy, a_b);
              ptx::tma_load_2d(&sg.tmW, &full_bar[stage], sb, k0, w_row);
            } else {
              if (leader) ptx::mbar_arrive_expect_tx(&full_bar[stage], 2 * stage_tx);
              ptx::tma_load_4d_2sm(&pr.tmA, &full_bar[stage], sa, cb * BK, x + dx, y + dy, a_b);
              ptx::tma_load_2d_2sm(&sg.tmW, &full_bar[stage], sb, k0, w_row);
            }
          } else if constexpr (kCtaGroup == 1) {
            ptx::mbar_arrive_expect_tx(&full_bar[stage], stage_tx);
            ptx::tma_load_3d_hint(&pr.tmA, &full_bar[stage], sa, k0, a_row, a_b, pol_a);
            ptx::tma_load_2d_hint(&sg.tmW, &full_bar[stage], sb, k0, w_row, pol_w);
          } else {
            if (leader) ptx::mbar_arrive_expect_tx(&full_bar[stage], 2 * stage_tx);
            ptx::tma_load_3d_2sm_hint(&pr.tmA, &full_bar[stage], sa, k0, a_row, a_b, pol_a);
            ptx::tma_load_2d_2sm_hint(&sg.tmW, &full_bar[stage], sb, k0, w_row, pol_w);
          }
        }
        __syncwarp();
        if (++stage == C::kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1 && leader) {
    // ===================== MMA issuer (leader CTA only; whole warp runs the loop, one elected lane issues) =====
    const uint32_t idesc = ptx::make_idesc_bf16(BM * kCtaGroup, kDynN ? bn : BN, 0, 0);
    // stage 0 descriptors; a stage advances the 14-bit (addr >> 4) field (the ring is < 256 KB: no carry out of it)
    const uint64_t adesc0 = ptx::make_smem_desc_sw128(ptx::smem_u32(smem_a), 0, 1024);
    const uint64_t bdesc0 = ptx::make_smem_desc_sw128(ptx::smem_u32(smem_b), 0, 1024);
    int stage = 0, phase = 0, iter = 0;
    for (int t = cluster_id; t < P.total_tiles; t += num_clusters, ++iter) {
      const TileCoord tc = decode_tile(P, t, bn, kRowsPerTile);
      const int nkb = (P.prob[tc.p].K + BK - 1) / BK;
      const int as = iter & 1, aphase = (iter >> 1) & 1;
      ptx::mbar_wait(&tempty_bar[as], aphase ^ 1);
      ptx::tc_fence_after();
      const uint32_t d_tmem = tmem_base + as * BN;
      for (int kb = 0; kb < nkb; ++kb) {
        ptx::mbar_wait(&full_bar[stage], phase);
        ptx::tc_fence_after();
        const uint64_t adesc = adesc0 + (uint64_t)(stage * (C::kABytes >> 4));
        const uint64_t bdesc = bdesc0 + (uint64_t)(stage * (C::kBBytes >> 4));
        if (ptx::elect_one()) {
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            // advance 16 bf16 = 32 B along K inside the 128-byte swizzle row: +2 in the (addr >> 4) field
            ptx::mma_bf16_ss<kCtaGroup>(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          if constexpr (kCtaGroup == 1) {
            ptx::mma_commit(&empty_bar[stage]);
            if (kb == nkb - 1) ptx::mma_commit(&tfull_bar[as]);
          } else {
            ptx::mma_commit_2sm(&empty_bar[stage], 3);
            if (kb == nkb - 1) ptx::mma_commit_2sm(&tfull_bar[as], 3);
          }
        }
        __syncwarp();
        if (++stage == C::kStages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    const int quad = warp & 3;
    constexpr int kParts = kEpiWarps / 4;
    const int part = (warp - 4) >> 2;
    uint8_t* stage_w = smem + C::kEpiStageOff + (warp - 4) * kStageWarpBytes;
    int iter = 0;
    for (int t = cluster_id; t < P.total_tiles; t += num_clusters, ++iter) {
      const TileCoord tc = decode_tile(P, t, bn, kRowsPerTile);
      const TcProblem& pr = P.prob[tc.p];
      const int as = iter & 1, aphase = (iter >> 1) & 1;
      const int m0w = tc.m0 + (int)cta_rank * BM + quad * 32;
      const TcSegment& sgt = pr.seg[tc.seg];
      const int ncols_t = kDynN ? min(bn, sgt.n_end - tc.n0) : BN;
      // operands of the epilogue that do not depend on the accumulator, fetched while the tile's MMAs still run
      const bool direct = (P.debug & 4) || sgt.out_f32;
      const bool first_staged = (P.debug & 1) || direct || (P.debug & 16)
                                    ? false
                                    : epilogue_prefetch<BN, kDynN>(P, pr, sgt, tc.b, m0w, lane, tc.n0, stage_w, ncols_t, part, kParts);
      ptx::mbar_wait(&tfull_bar[as], aphase);
      ptx::tc_fence_after();
      const uint32_t tacc = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + as * BN;
      if (P.debug & 1) {
        // timing experiment: accumulators are dropped
      } else if (direct) {
        epilogue_tile<BN>(P, pr, sgt, tacc, tc.b, m0w + lane, tc.n0, part, kParts);
      } else {
        epilogue_tile_staged<BN, kDynN>(P, pr, sgt, tacc, tc.b, m0w, lane, tc.n0, stage_w, ncols_t, part, kParts, first_staged);
      }
      ptx::tc_fence_before();
      // The hand-off's payload is TMEM (tcgen05.ld completed by tcgen05.wait::ld, ordered by the fence above), so the
      // remote arrive needs no release: .release.cluster costs MEMBAR.ALL.GPU + ERRBAR per thread and tile, i.e. every
      // epilogue thread would wait for its global stores to be acknowledged before freeing the accumulator.
      if constexpr (kCtaGroup == 1) ptx::mbar_arrive(&tempty_bar[as]);
      else ptx::mbar_arrive_cluster_relaxed(&tempty_bar[as], 0);
    }
  }

  ptx::tc_fence_before();
  if constexpr (kCtaGroup == 2) ptx::cluster_sync(); else __syncthreads();
  if (warp == 2) ptx::tmem_dealloc<kCtaGroup>(tmem_base, C::kTmemCols);
}

// -------------------------------------------------------------------------------------------------
// host side
// -------------------------------------------------------------------------------------------------
using EncodeFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                              const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                              CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeFn get_encode_fn() {
  static EncodeFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    if (e == cudaSuccess && q == cudaDriverEntryPointSuccess) fn = reinterpret_cast<EncodeFn>(p);
  });
  if (!fn) throw Error(RT_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
  return fn;
}

struct TmapKey {
  const void* base;
  uint64_t d[4], s[3];
  uint32_t box[4];
  int rank;
  bool operator==(const TmapKey& o) const {
    return base == o.base && rank == o.rank && d[0] == o.d[0] && d[1] == o.d[1] && d[2] == o.d[2] && d[3] == o.d[3] &&
           s[0] == o.s[0] && s[1] == o.s[1] && s[2] == o.s[2] && box[0] == o.box[0] && box[1] == o.box[1] &&
           box[2] == o.box[2] && box[3] == o.box[3];
  }
};
struct TmapKeyHash {
  size_t operator()(const TmapKey& k) const {
    size_t h = reinterpret_cast<size_t>(k.base);
    auto mix = [&h](uint64_t v) { h ^= v + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2); };
    mix(k.d[0]); mix(k.d[1]); mix(k.d[2]); mix(k.d[3]); mix(k.s[0]); mix(k.s[1]); mix(k.s[2]); mix(k.box[1]);
    mix(k.box[2]); mix(k.box[3]); mix(k.rank);
    return h;
  }
};

}  // namespace

void encode_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_b,
                      const uint32_t* box) {
  static std::unordered_map<TmapKey, CUtensorMap, TmapKeyHash> cache;
  static std::mutex mu;
  TmapKey key{};
  key.base = base;
  key.rank = rank;
  for (int i = 0; i < rank; ++i) { key.d[i] = dims[i]; key.box[i] = box[i]; }
  for (int i = 0; i + 1 < rank; ++i) key.s[i] = strides_b[i];
  {
    std::lock_guard<std::mutex> lk(mu);
    auto it = cache.find(key);
    if (it != cache.end()) { *out = it->second; return; }
  }
  RT_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA base must be 16-byte aligned");
  cuuint64_t gd[4], gs[3];
  cuuint32_t bx[4], es[4] = {1, 1, 1, 1};
  for (int i = 0; i < rank; ++i) { gd[i] = dims[i]; bx[i] = box[i]; }
  for (int i = 0; i + 1 < rank; ++i) {
    RT_REQUIRE(strides_b[i] % 16 == 0, "TMA strides must be multiples of 16 bytes");
    gs[i] = strides_b[i];
  }
  CUresult r = get_encode_fn()(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gd, gs,
                               bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) throw Error(RT_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r));
  std::lock_guard<std::mutex> lk(mu);
  if (cache.size() > 8192) cache.clear();
  cache.emplace(key, *out);
}

static int pick_bn(const GemmLaunch& L) {
  for (int bn : {256, 128, 64}) {
    bool ok = true;
    for (int p = 0; p < L.nprob && ok; ++p)
      for (int s = 0; s < L.prob[p].nseg && ok; ++s) {
        const GemmSegment& S = L.prob[p].seg[s];
        if (S.n_begin % bn || S.n_end % bn) ok = false;
        if (S.mode == EPI_QKNORM_ROPE && bn < 128) ok = false;
      }
    if (ok) return bn;
  }
  return 0;
}

bool gemm_tc_supported(const GemmLaunch& L, std::string* why) {
  auto fail = [&](const char* m) { if (why) *why = m; return false; };
  if (L.dtype != RT_BF16) return fail("dtype is not bf16");
  if (L.nprob < 1 || L.nprob > 2) return fail("nprob");
  if (pick_bn(L) == 0) return fail("segment boundaries are not multiples of 64 (or 128 for qk-norm)");
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  for (int p = 0; p < L.nprob; ++p) {
    const GemmProblem& P = L.prob[p];
    if (P.K % 8 || P.a_ld % 8 || P.a_batch_stride % 8) return fail("K / a_ld / a_batch_stride not multiples of 8");
    if (P.conv_w > 0) {
      const int kcb = (P.conv_c + 63) / 64;
      if (P.conv_h <= 0 || P.conv_c <= 0 || P.conv_c % 8) return fail("conv: bad image shape");
      if (!(P.conv_w % 128 == 0 || 128 % P.conv_w == 0) || P.conv_w < 8) return fail("conv: width must divide or be a multiple of 128");
      if (P.conv_w < 128 && P.conv_h % (128 / P.conv_w)) return fail("conv: height must be a multiple of 128 / width");
      if (P.K != 9 * kcb * 64) return fail("conv: K must be 9 * ceil(C / 64) * 64 (tap-major, zero-padded weights)");
      if (P.m_rows != P.conv_h * P.conv_w || P.a_row0 != 0) return fail("conv: m_rows must be H * W");
      if (P.a_batch_stride == 0) return fail("conv: broadcast A is not supported");
    }
    if (!al16(P.A)) return fail("A not 16-byte aligned");
    if (P.nseg < 1 || P.nseg > kMaxSeg) return fail("nseg");
    if (P.gate && (!al16(P.gate) || P.gate_ld % 4)) return fail("gate alignment");
    if (P.extra && (!al16(P.extra) || P.extra_ld % 8 || P.extra_batch_stride % 8)) return fail("extra alignment");
    int expect = 0;
    for (int s = 0; s < P.nseg; ++s) {
      const GemmSegment& S = P.seg[s];
      if (S.n_begin != expect || S.n_end <= S.n_begin) return fail("segments must tile [0, N) in order");
      expect = S.n_end;
      if (!al16(S.W) || (!S.scatter && (!S.out || !al16(S.out))) || (S.bias && !al16(S.bias)))
        return fail("W / out / bias alignment");
      if (S.out_ld % 8 || S.out_col0 % 8 || S.out_batch_stride % 8) return fail("out ld / col0 / stride alignment");
      if (S.out_f32) {
        if (S.scatter || (S.mode != EPI_BIAS && S.mode != EPI_SCALE_MASK)) return fail("out_f32 needs a BIAS or SCALE_MASK segment");
        if (S.mode == EPI_SCALE_MASK && (P.mask || P.accumulate)) return fail("out_f32: no mask / accumulate");
      }
      if (S.scatter) {
        if (S.mode != EPI_BIAS && S.mode != EPI_QKNORM_ROPE) return fail("scatter needs a BIAS or QKNORM_ROPE segment");
        if (L.sp_cols <= 0 || L.sp_cols % 128) return fail("sp_cols must be a positive multiple of 128");
        const int ndest = (S.n_end - S.n_begin + L.sp_cols - 1) / L.sp_cols;
        if (ndest > RT_SP_MAX_RANKS) return fail("scatter: more destinations than RT_SP_MAX_RANKS");
        for (int d = 0; d < ndest; ++d)
          if (!L.sp_out[d] || !al16(L.sp_out[d])) return fail("scatter: sp_out pointer missing or misaligned");
      }
      if (S.mode == EPI_QKNORM_ROPE) {
        if (L.head_dim != 128) return fail("fused qk-norm needs head_dim 128");
        if (!S.norm_w || !al16(S.norm_w)) return fail("norm_w");
        if (L.rope && !al16(L.rope)) return fail("rope alignment");
      }
    }
  }
  return true;
}

template <int BN, int CG, bool kDynN = false, int kEpiWarps = 4>
static void launch_cfg(const TcParams& P, int num_sms, cudaStream_t stream) {
  using C = Cfg<BN, CG, kEpiWarps>;
  static PerDeviceOnce attr_set;
  if (attr_set.first()) {
    RT_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<BN, CG, kDynN, kEpiWarps>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       C::kSmemBytes));
  }
  int clusters = num_sms / CG;
  if (clusters > P.total_tiles) clusters = P.total_tiles;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(clusters * CG);
  cfg.blockDim = dim3(128 + 32 * kEpiWarps);
  cfg.dynamicSmemBytes = C::kSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = get_option("no_pdl") ? 1 : 2;
  RT_CHECK_CUDA(cudaLaunchKernelEx(&cfg, gemm_tc_kernel<BN, CG, kDynN, kEpiWarps>, P));
  count_launch();
}

// Run-time tile width (kDynN kernels).  A persistent launch takes ceil(tiles / slots) tile times; a tile's time grows
// with its width as (bn + 64) - measured, profiles/r2_gemm_dyn_tile_width.txt: a 128 x N x 16 tcgen05.mma takes
// ~23 + 0.41 N cycles, and every tile pays the same A traffic - so 256-wide tiles that leave the last wave mostly empty
// lose to narrower ones that fill it (1216 rows x N = 3072 on CTA pairs: 60 tiles of 256 for 74 pairs, or 70 tiles of
// 224: measured 36.1 -> 30.4 us at K = 3072, 102.9 -> 90.2 us at K = 15360), while a narrower tile that only adds waves
// loses.  Only
// launches whose every problem is ONE segment with a column-local epilogue qualify (no per-head RMSNorm, no scatter to
// 128-column destinations, no fp32 output), and only a clear win (>= 5 %) switches.  Returns 256 to keep the static form.
static int pick_dyn_bn(const GemmLaunch& L, int cg, int num_sms) {
  const int opt = get_option("gemm_dyn_bn");  // 0 = auto, -1 = never, n = force tiles of n columns where eligible
  if (opt < 0 || (get_option("gemm_debug") & 4)) return 256;
  for (int p = 0; p < L.nprob; ++p) {
    const GemmProblem& G = L.prob[p];
    if (G.nseg != 1 || G.conv_w > 0) return 256;
    const GemmSegment& S = G.seg[0];
    if (S.scatter || S.out_f32 || S.mode == EPI_QKNORM_ROPE || (S.n_end - S.n_begin) % 32) return 256;
  }
  const int slots = num_sms / cg;
  auto cost = [&](int bn) {
    long long t = 0;
    for (int p = 0; p < L.nprob; ++p)
      t += (long long)((L.prob[p].m_rows + BM * cg - 1) / (BM * cg)) * ((gemm_total_n(L.prob[p]) + bn - 1) / bn) * L.batch;
    return (double)((t + slots - 1) / slots) * (bn + 64);
  };
  if (opt > 0) return (opt % 32 == 0 && opt >= 64 * cg && opt <= 256) ? opt : 256;
  int best = 256;
  double best_cost = cost(256) * 0.95;
  for (int bn : {224, 192, 160}) {
    const double c = cost(bn);
    if (c < best_cost) { best = bn; best_cost = c; }
  }
  return best;
}

void launch_gemm_tc(const GemmLaunch& L, cudaStream_t stream, int force_cta_group, const SpSyncParams* sync) {
  std::string why;
  if (!gemm_tc_supported(L, &why)) throw Error(RT_ERR_UNSUPPORTED, "tcgen05 GEMM: " + why);
  if (L.batch == 0) return;
  int bn = pick_bn(L);
  int cg = force_cta_group;
  if (cg == 0) {
    // auto: pair two SMs on 256-row tiles (each CTA then stages only half of the W tile: less shared-memory
    // traffic per MMA, measured +5-8 % on the step's shapes) unless that pads some problem with extra rows
    cg = 2;
    for (int p = 0; p < L.nprob; ++p) {
      const int m = L.prob[p].m_rows;
      if ((m + 255) / 256 * 256 != (m + 127) / 128 * 128) cg = 1;
    }
  }
  RT_REQUIRE(cg == 1 || cg == 2, "cta_group must be 1 or 2");
  if (cg == 2 && bn < 128) cg = 1;  // each CTA must hold at least 64 rows of W

  const int num_sms = device_sm_count();
  // Small launches (the prompt encoders at 512 tokens: 32-160 tiles of 256 x 256 for 74 CTA pairs): when 128-wide
  // single-CTA tiles fill clearly more of the machine, take them - a 128-wide tile runs at ~85 % of the 256-wide rate
  // (profiles/r1_gemm_after_uniform_issue.txt), so the fill has to win by more than that.  Same arithmetic per element.
  if (force_cta_group == 0 && bn == 256) {
    auto tiles = [&](int rows_per_tile, int cols) {
      long long t = 0;
      for (int p = 0; p < L.nprob; ++p)
        t += (long long)((L.prob[p].m_rows + rows_per_tile - 1) / rows_per_tile) * (gemm_total_n(L.prob[p]) / cols) * L.batch;
      return t;
    };
    auto fill = [](long long t, int slots) { return t <= 0 ? 0.0 : (double)t / (double)(((t + slots - 1) / slots) * slots); };
    bool conv = false;
    for (int p = 0; p < L.nprob; ++p) conv |= L.prob[p].conv_w > 0;
    const int slots = num_sms / cg;
    const long long t_now = tiles(BM * cg, 256);
    if (!conv && t_now < 2 * slots) {
      const double eff_now = fill(t_now, slots), eff_alt = 0.85 * fill(tiles(BM, 128), num_sms);
      if (eff_alt > 1.10 * eff_now) { bn = 128; cg = 1; }
    }
  }

  // narrower run-time tiles when they fill the last wave better (bit-identical results; see pick_dyn_bn)
  const int dyn_bn = bn == 256 ? pick_dyn_bn(L, cg, num_sms) : 256;
  const bool dyn = dyn_bn != 256;
  if (dyn) bn = dyn_bn;

  TcParams P{};
  if (sync) P.sync = *sync;
  P.bn = bn;
  P.nprob = L.nprob;
  P.batch = L.batch;
  P.rope = reinterpret_cast<const float2*>(L.rope);
  P.head_dim = L.head_dim;
  P.sp_cols = L.sp_cols; P.sp_row0 = L.sp_row0;
  P.debug = get_option("gemm_debug");
  for (int i = 0; i < RT_SP_MAX_RANKS; ++i) P.sp_out[i] = reinterpret_cast<bf16*>(L.sp_out[i]);
  int tile_base = 0;
  for (int p = 0; p < L.nprob; ++p) {
    const GemmProblem& G = L.prob[p];
    TcProblem& T = P.prob[p];
    const bool bcast = G.a_batch_stride == 0;
    if (G.conv_w > 0) {
      const uint32_t bw = G.conv_w >= 128 ? 128u : (uint32_t)G.conv_w;
      uint64_t dims[4] = {(uint64_t)G.conv_c, (uint64_t)G.conv_w, (uint64_t)G.conv_h, (uint64_t)L.batch};
      uint64_t strides[3] = {(uint64_t)G.a_ld * 2, (uint64_t)G.conv_w * G.a_ld * 2, (uint64_t)G.a_batch_stride * 2};
      uint32_t box[4] = {BK, bw, 128u / bw, 1};
      encode_tmap_bf16(&T.tmA, G.A, 4, dims, strides, box);
      T.conv_w = G.conv_w; T.conv_bw = (int)bw; T.conv_kcb = (G.conv_c + 63) / 64;
    } else {
      uint64_t dims[3] = {(uint64_t)G.K, (uint64_t)G.a_rows_total, (uint64_t)(bcast ? 1 : L.batch)};
      uint64_t strides[2] = {(uint64_t)G.a_ld * 2, (uint64_t)(bcast ? (long long)G.a_rows_total * G.a_ld : G.a_batch_stride) * 2};
      uint32_t box[3] = {BK, BM, 1};
      encode_tmap_bf16(&T.tmA, G.A, 3, dims, strides, box);
    }
    T.a_row0 = G.a_row0; T.a_bcast = bcast; T.m_rows = G.m_rows; T.out_row0 = G.out_row0; T.K = G.K;
    T.nseg = G.nseg;
    T.gate = G.gate; T.gate_ld = G.gate_ld;
    T.extra = reinterpret_cast<const bf16*>(G.extra); T.e_bs = G.extra_batch_stride; T.e_ld = G.extra_ld;
    T.e_row0 = G.extra_row0;
    T.mask = reinterpret_cast<const bf16*>(G.mask); T.scale = G.scale; T.accumulate = G.accumulate;
    for (int s = 0; s < G.nseg; ++s) {
      const GemmSegment& S = G.seg[s];
      TcSegment& D = T.seg[s];
      uint64_t wd[2] = {(uint64_t)G.K, (uint64_t)(S.n_end - S.n_begin)};
      uint64_t ws[1] = {(uint64_t)G.K * 2};
      uint32_t wb[2] = {BK, (uint32_t)(bn / cg)};
      encode_tmap_bf16(&D.tmW, S.W, 2, wd, ws, wb);
      D.bias = reinterpret_cast<const bf16*>(S.bias);
      D.out = reinterpret_cast<bf16*>(S.out);
      D.norm_w = reinterpret_cast<const bf16*>(S.norm_w);
      D.out_bs = S.out_batch_stride; D.n_begin = S.n_begin; D.n_end = S.n_end; D.mode = S.mode;
      D.out_ld = S.out_ld; D.out_col0 = S.out_col0; D.scatter = S.scatter; D.out_f32 = S.out_f32;
    }
    const int rows_per_tile = BM * cg;
    T.tiles_m = (G.m_rows + rows_per_tile - 1) / rows_per_tile;
    T.tiles_n = (gemm_total_n(G) + bn - 1) / bn;  // exact unless dyn (pick_bn)
    T.tile_base = tile_base;
    T.num_tiles = T.tiles_m * T.tiles_n * L.batch;
    // option "gemm_band": 0 = auto, -1 = never, n > 0 = bands of n row tiles for every problem (tests, A/B)
    const int band_opt = get_option("gemm_band");
    const int slots = num_sms / cg;  // tiles in flight
    T.band = 0;
    if (band_opt > 0) {
      T.band = band_opt;
    } else if (band_opt == 0 && G.conv_w == 0) {
      // A is re-read once per wave when it cannot stay in L2 (126 MB, shared with W and the output): band it when it is
      // large, there is more than one wave, and a band of at least two row tiles fills the machine
      const double a_bytes = (double)G.m_rows * G.K * 2.0 * (bcast ? 1 : L.batch);
      const int band = slots / T.tiles_n;
      if (a_bytes > 64e6 && T.num_tiles > slots && band >= 2 && band < T.tiles_m) T.band = band;
    }
    tile_base += T.num_tiles;
  }
  P.total_tiles = tile_base;
  if (P.total_tiles == 0) return;

  // option "gemm_epi_warps": 0 = auto (eight epilogue warps on the CTA-pair kernels), 4 = four everywhere (A/B)
  const bool epi8 = cg == 2 && get_option("gemm_epi_warps") != 4;
  if (dyn) {
    if (cg == 1) launch_cfg<256, 1, true>(P, num_sms, stream);
    else if (epi8) launch_cfg<256, 2, true, 8>(P, num_sms, stream);
    else launch_cfg<256, 2, true>(P, num_sms, stream);
    return;
  }
  if (epi8 && bn == 256) { launch_cfg<256, 2, false, 8>(P, num_sms, stream); return; }
  if (epi8 && bn == 128) { launch_cfg<128, 2, false, 8>(P, num_sms, stream); return; }
#define RT_GEMM_CASE(BN_, CG_) \
  if (bn == BN_ && cg == CG_) { launch_cfg<BN_, CG_>(P, num_sms, stream); return; }
  RT_GEMM_CASE(256, 1)
  RT_GEMM_CASE(128, 1)
  RT_GEMM_CASE(64, 1)
  RT_GEMM_CASE(256, 2)
  RT_GEMM_CASE(128, 2)
#undef RT_GEMM_CASE
  throw Error(RT_ERR_INTERNAL, "no GEMM instantiation");
}

}  // namespace rt
