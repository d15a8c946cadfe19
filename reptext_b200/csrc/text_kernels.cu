// Kernels of the prompt-encoder path (the T5-XXL encoder and the CLIP-L text model the RepText
// pipelines call once per image, RepText/pipeline_flux_controlnet.py:232-347).  The projections and MLPs run on the
// tcgen05 GEMM (rt_gemm); here are the HBM- / latency-bound pieces around it:
//   * row norms: T5LayerNorm (RMS, no mean, no bias) and CLIP's LayerNorm (affine),
//   * attention for head_dim 64 with the two things FLUX's joint attention does not have: an additive relative-position
//     bias that depends on (key - query) only (T5) and a causal mask (CLIP); S <= 512 tokens, <3 % of the encoder's
//     FLOPs: warp-level bf16 tensor-core MMAs with an fp32 online softmax (a CUDA-core form is kept for A/B),
//   * the gated activation (T5: gelu(wi_0 x) * wi_1 x, the GELU already applied by the GEMM epilogue; CLIP: quick-GELU),
//   * the token (+ position) embedding gather.
#include <mma.h>

#include <cmath>

#include "dtype_utils.cuh"
#include "ptx_sm100.cuh"
#include "rt_internal.h"

namespace rt {
namespace {

// ---- row norm: out = (x - mean?) * rsqrt(var + eps) * w (+ b); one CTA of 128 threads per row, D <= 4096 in registers
template <bool kCenter>
__global__ void __launch_bounds__(128) norm_rows_kernel(const bf16* __restrict__ x, long long x_ld, bf16* __restrict__ out,
                                                        long long out_ld, int D, const bf16* __restrict__ w,
                                                        const bf16* __restrict__ b, float eps) {
  constexpr int kMaxVec = 4;                       // 128 threads * 4 vectors * 8 elements = 4096
  __shared__ float red[4];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const bf16* xr = x + (long long)blockIdx.x * x_ld;
  bf16* orow = out + (long long)blockIdx.x * out_ld;
  const int vecs = D / 8;
  float t[kMaxVec][8];
  float s = 0.f;
#pragma unroll
  for (int u = 0; u < kMaxVec; ++u) {
    const int v = threadIdx.x + u * 128;
    if (v < vecs) {
      ldvec(xr + v * 8, t[u]);
#pragma unroll
      for (int j = 0; j < 8; ++j) s += kCenter ? t[u][j] : t[u][j] * t[u][j];
    }
  }
  auto block_sum = [&](float v) {
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    return red[0] + red[1] + red[2] + red[3];
  };
  float mean = 0.f, var;
  if (kCenter) {
    mean = block_sum(s) / (float)D;
    float q = 0.f;
#pragma unroll
    for (int u = 0; u < kMaxVec; ++u) {
      const int v = threadIdx.x + u * 128;
      if (v < vecs) {
#pragma unroll
        for (int j = 0; j < 8; ++j) { const float d = t[u][j] - mean; q += d * d; }
      }
    }
    var = block_sum(q) / (float)D;
  } else {
    var = block_sum(s) / (float)D;
  }
  const float rstd = rsqrtf(var + eps);
#pragma unroll
  for (int u = 0; u < kMaxVec; ++u) {
    const int v = threadIdx.x + u * 128;
    if (v < vecs) {
      float wv[8], bv[8], o[8];
      ldvec(w + v * 8, wv);
      if (b) ldvec(b + v * 8, bv);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = (t[u][j] - mean) * rstd * wv[j] + (b ? bv[j] : 0.f);
      stvec(orow + v * 8, o);
    }
  }
}

// ---- attention, head_dim 64: block = (32 query rows, head, batch), 8 warps x 4 rows, K / V tiles of 32 keys in smem.
// score = scale * q.k + rel_bias[h][key - query + S - 1] (if rel_bias) ; keys > query masked (if causal).
struct TextAttnArgs {
  const bf16* qkv;
  long long batch_stride;
  int ld, q_col0, k_col0, v_col0;
  bf16* out;
  long long out_batch_stride;
  int out_ld, out_col0;
  int S, heads;
  float scale;
  const float* rel_bias;   // [heads, 2 S - 1] or null
  int causal;
};

__global__ void __launch_bounds__(256) text_attn_kernel(TextAttnArgs a) {
  constexpr int HD = 64, KT = 32, RPW = 4, DPL = HD / 32;
  extern __shared__ float smem[];
  float* Ks = smem;                      // [KT][HD + 1]
  float* Vs = Ks + KT * (HD + 1);        // [KT][HD]
  float* Qs = Vs + KT * HD;              // [32][HD]
  float* Bs = Qs + 32 * HD;              // [2 S - 1] (rel_bias row of this head)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int h = blockIdx.y, b = blockIdx.z;
  const int q0 = blockIdx.x * 32;
  const bf16* base = a.qkv + (long long)b * a.batch_stride;
  for (int e = threadIdx.x; e < 32 * HD; e += 256) {
    const int r = e / HD, d = e % HD;
    Qs[e] = (q0 + r < a.S) ? to_f(base[(long long)(q0 + r) * a.ld + a.q_col0 + h * HD + d]) * a.scale : 0.f;
  }
  if (a.rel_bias)
    for (int e = threadIdx.x; e < 2 * a.S - 1; e += 256) Bs[e] = a.rel_bias[(long long)h * (2 * a.S - 1) + e];
  float m[RPW], l[RPW], o[RPW][DPL];
#pragma unroll
  for (int r = 0; r < RPW; ++r) {
    m[r] = -INFINITY;
    l[r] = 0.f;
#pragma unroll
    for (int i = 0; i < DPL; ++i) o[r][i] = 0.f;
  }
  const int k_end = a.causal ? min(a.S, q0 + 32) : a.S;     // keys beyond the block's last query are all masked
  for (int k0 = 0; k0 < k_end; k0 += KT) {
    __syncthreads();
    for (int e = threadIdx.x; e < KT * HD; e += 256) {
      const int j = e / HD, d = e % HD;
      const bool ok = k0 + j < a.S;
      const long long off = (long long)(k0 + j) * a.ld + h * HD + d;
      Ks[j * (HD + 1) + d] = ok ? to_f(base[off + a.k_col0]) : 0.f;
      Vs[j * HD + d] = ok ? to_f(base[off + a.v_col0]) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
      const int row = q0 + warp * RPW + r;
      const int key = k0 + lane;
      const float* q = Qs + (warp * RPW + r) * HD;
      float s = 0.f;
#pragma unroll 8
      for (int d = 0; d < HD; ++d) s += q[d] * Ks[lane * (HD + 1) + d];
      if (a.rel_bias && key < a.S && row < a.S) s += Bs[key - row + a.S - 1];
      if (key >= a.S || (a.causal && key > row)) s = -INFINITY;
      // key 0 is never masked, so the running maximum is finite from the first tile on
      const float mx = fmaxf(m[r], warp_max(s));
      const float p = __expf(s - mx);
      const float corr = __expf(m[r] - mx);
      l[r] = l[r] * corr + warp_sum(p);
      m[r] = mx;
#pragma unroll
      for (int i = 0; i < DPL; ++i) o[r][i] *= corr;
#pragma unroll 8
      for (int j = 0; j < KT; ++j) {
        const float pj = __shfl_sync(0xffffffffu, p, j);
#pragma unroll
        for (int i = 0; i < DPL; ++i) o[r][i] += pj * Vs[j * HD + lane + 32 * i];
      }
    }
  }
  bf16* ob = a.out + (long long)b * a.out_batch_stride;
#pragma unroll
  for (int r = 0; r < RPW; ++r) {
    const int row = q0 + warp * RPW + r;
    if (row >= a.S) continue;
    const float inv = 1.f / l[r];
#pragma unroll
    for (int i = 0; i < DPL; ++i)
      ob[(long long)row * a.out_ld + a.out_col0 + h * HD + lane + 32 * i] = __float2bfloat16_rn(o[r][i] * inv);
  }
}

// ---- the same attention on the tensor cores (warp-level bf16 MMA, fp32 accumulate): block = (64 query rows, head,
// batch), 4 warps x 16 rows; K / V tiles of 64 keys in shared memory as bf16.  Per key tile a warp computes its
// 16 x 64 scores with 16 MMAs into shared memory, two threads per row run the online softmax there (scale, relative bias,
// causal mask, running max / sum, rescale of the output rows) and write P as bf16, and 16 more MMAs add P V to the
// fp32 output tile kept in shared memory (the fragment layout of this API is opaque, so the per-row rescale goes
// through memory).  S <= 512 here: 8 query blocks x 64 heads = 512 CTAs per T5 block.
constexpr int kTA_QT = 64, kTA_KT = 64, kTA_HD = 64;
constexpr int kTA_LDH = kTA_HD + 8;   // bf16 row stride (elements): 144 bytes, keeps 16 x 16 tile loads off one bank
constexpr int kTA_LDF = kTA_KT + 8;   // fp32 row stride (elements)
constexpr size_t kTA_SmemFixed = (size_t)4 * kTA_QT * kTA_LDH * 2 + (size_t)2 * kTA_QT * kTA_LDF * 4;

__global__ void __launch_bounds__(128) text_attn_mma_kernel(TextAttnArgs a) {
  using namespace nvcuda;
  constexpr int QT = kTA_QT, KT = kTA_KT, HD = kTA_HD, LDH = kTA_LDH, LDF = kTA_LDF;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  bf16* Qs = reinterpret_cast<bf16*>(smem_raw);          // [QT][LDH]
  bf16* Ks = Qs + QT * LDH;                              // [KT][LDH]
  bf16* Vs = Ks + KT * LDH;                              // [KT][LDH]
  bf16* Ps = Vs + KT * LDH;                              // [QT][LDH]
  float* Ss = reinterpret_cast<float*>(Ps + QT * LDH);   // [QT][LDF]
  float* Os = Ss + QT * LDF;                             // [QT][LDF]
  float* Bs = Os + QT * LDF;                             // [2 S - 1]
  const int tid = threadIdx.x, warp = tid >> 5;
  const int h = blockIdx.y, b = blockIdx.z;
  const int q0 = blockIdx.x * QT;
  const bf16* base = a.qkv + (long long)b * a.batch_stride;
  const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);
  for (int e = tid; e < QT * (HD / 8); e += 128) {
    const int r = e / (HD / 8), v = e % (HD / 8);
    uint4 t = zero4;
    if (q0 + r < a.S) t = *reinterpret_cast<const uint4*>(base + (long long)(q0 + r) * a.ld + a.q_col0 + h * HD + v * 8);
    *reinterpret_cast<uint4*>(Qs + r * LDH + v * 8) = t;
  }
  for (int e = tid; e < QT * LDF; e += 128) Os[e] = 0.f;
  if (a.rel_bias)
    for (int e = tid; e < 2 * a.S - 1; e += 128) Bs[e] = a.rel_bias[(long long)h * (2 * a.S - 1) + e];
  // softmax ownership: thread t -> row t / 2 (a row of ITS warp's 16), columns (t % 2) * 32 ..
  const int row = tid >> 1, half = tid & 1;
  const int qrow = q0 + row;
  float m_run = -INFINITY, l_run = 0.f;
  const int k_end = a.causal ? min(a.S, q0 + QT) : a.S;
  for (int k0 = 0; k0 < k_end; k0 += KT) {
    __syncthreads();     // the previous tile's MMAs are done with Ks / Vs (and Qs / Os / Bs are written, first time round)
    for (int e = tid; e < KT * (HD / 8); e += 128) {
      const int r = e / (HD / 8), v = e % (HD / 8);
      uint4 kk = zero4, vv = zero4;
      if (k0 + r < a.S) {
        const bf16* src = base + (long long)(k0 + r) * a.ld + h * HD + v * 8;
        kk = *reinterpret_cast<const uint4*>(src + a.k_col0);
        vv = *reinterpret_cast<const uint4*>(src + a.v_col0);
      }
      *reinterpret_cast<uint4*>(Ks + r * LDH + v * 8) = kk;
      *reinterpret_cast<uint4*>(Vs + r * LDH + v * 8) = vv;
    }
    __syncthreads();
    // S = Q K^T for this warp's 16 rows
#pragma unroll
    for (int n = 0; n < KT / 16; ++n) {
      wmma::fragment<wmma::accumulator, 16, 16, 16, float> acc;
      wmma::fill_fragment(acc, 0.f);
#pragma unroll
      for (int k = 0; k < HD / 16; ++k) {
        wmma::fragment<wmma::matrix_a, 16, 16, 16, bf16, wmma::row_major> fa;
        wmma::fragment<wmma::matrix_b, 16, 16, 16, bf16, wmma::col_major> fb;
        wmma::load_matrix_sync(fa, Qs + warp * 16 * LDH + k * 16, LDH);
        wmma::load_matrix_sync(fb, Ks + n * 16 * LDH + k * 16, LDH);   // (d, key) at [key * LDH + d]
        wmma::mma_sync(acc, fa, fb, acc);
      }
      wmma::store_matrix_sync(Ss + warp * 16 * LDF + n * 16, acc, LDF, wmma::mem_row_major);
    }
    __syncwarp();
    // online softmax on 32 scores of one row per thread
    {
      float* srow = Ss + row * LDF + half * 32;
      float sc[32];
      float tmax = -INFINITY;
#pragma unroll
      for (int c = 0; c < 32; ++c) {
        const int key = k0 + half * 32 + c;
        float s = srow[c] * a.scale;
        if (a.rel_bias && key < a.S && qrow < a.S) s += Bs[key - qrow + a.S - 1];
        if (key >= a.S || (a.causal && key > qrow)) s = -INFINITY;
        sc[c] = s;
        tmax = fmaxf(tmax, s);
      }
      tmax = fmaxf(tmax, __shfl_xor_sync(0xffffffffu, tmax, 1));
      const float m_new = fmaxf(m_run, tmax);        // finite: key 0 of the first tile is never masked
      const float corr = __expf(m_run - m_new);
      float psum = 0.f;
      bf16* prow = Ps + row * LDH + half * 32;
#pragma unroll
      for (int c = 0; c < 32; c += 2) {
        const float p0 = __expf(sc[c] - m_new), p1 = __expf(sc[c + 1] - m_new);
        const __nv_bfloat162 pb = __floats2bfloat162_rn(p0, p1);
        // the sum uses the ROUNDED probabilities, the ones the P V product sees
        psum += __bfloat162float(pb.x) + __bfloat162float(pb.y);
        *reinterpret_cast<__nv_bfloat162*>(prow + c) = pb;
      }
      psum += __shfl_xor_sync(0xffffffffu, psum, 1);
      l_run = l_run * corr + psum;
      m_run = m_new;
      float* orow = Os + row * LDF + half * 32;
#pragma unroll
      for (int c = 0; c < 32; ++c) orow[c] *= corr;
    }
    __syncwarp();
    // O += P V for this warp's 16 rows
#pragma unroll
    for (int n = 0; n < HD / 16; ++n) {
      wmma::fragment<wmma::accumulator, 16, 16, 16, float> acc;
      wmma::load_matrix_sync(acc, Os + warp * 16 * LDF + n * 16, LDF, wmma::mem_row_major);
#pragma unroll
      for (int k = 0; k < KT / 16; ++k) {
        wmma::fragment<wmma::matrix_a, 16, 16, 16, bf16, wmma::row_major> fa;
        wmma::fragment<wmma::matrix_b, 16, 16, 16, bf16, wmma::row_major> fb;
        wmma::load_matrix_sync(fa, Ps + warp * 16 * LDH + k * 16, LDH);
        wmma::load_matrix_sync(fb, Vs + k * 16 * LDH + n * 16, LDH);   // (key, d) at [key * LDH + d]
        wmma::mma_sync(acc, fa, fb, acc);
      }
      wmma::store_matrix_sync(Os + warp * 16 * LDF + n * 16, acc, LDF, wmma::mem_row_major);
    }
    __syncwarp();
  }
  if (qrow < a.S) {
    const float inv = 1.f / l_run;
    const float* orow = Os + row * LDF + half * 32;
    bf16* dst = a.out + (long long)b * a.out_batch_stride + (long long)qrow * a.out_ld + a.out_col0 + h * HD + half * 32;
#pragma unroll
    for (int c = 0; c < 32; c += 8) {
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = orow[c + j] * inv;
      stvec(dst + c, o);
    }
  }
}

// ---- the same attention on the 5th-generation tensor cores: tcgen05.mma with the accumulators in tensor memory, Q / K /
// V tiles fetched by TMA straight from the fused projection buffer (128B swizzle), the structure of the hot path's
// attn_tc_kernel cut down to what 512 tokens need.  CTA = (128 query rows, head, batch); key blocks of 128:
//   warp 0  TMA producer (Q once, K / V through a ring of two stages)
//   warp 1  MMA issuer: S = Q K^T (SS, 4 k-steps of 16) -> tensor memory; O += P V (TS: P read from tensor memory, V
//           MN-major, 8 k-steps) -> tensor memory; also allocates / frees the 256 columns (S 128, O 64)
//   warps 2-5  softmax, one thread per query row: the 128 scores of the block in registers, log2-domain scale +
//           relative-position bias (T5: a [2 S - 1] row of this head, staged in shared memory, indexed by key - query) or
//           causal mask (CLIP), online maximum with an exact rescale of O in tensor memory (skipped while the maximum
//           stands), P = 2^(x - m) as bf16 back over the score columns, row sum in fp32
// One CTA per SM and no ping-pong: the tensor pipe waits while the softmax runs - this path is < 3 % of an encoder's
// FLOPs and 24 x 256 CTAs per prompt; the point is the datapath, not the last 2x.
constexpr int kTT_Threads = 192;
constexpr int kTT_TileBytes = 128 * 64 * 2;   // 128 rows x 64 head-dim columns, bf16: 16 KB
constexpr int kTT_BarOff = 5 * kTT_TileBytes;
constexpr int kTT_BiasOff = kTT_BarOff + 128;

struct TextTcParams {
  CUtensorMap tm;   // 3-D (columns of the projection buffer, rows, batch), box (64, 128, 1)
  TextAttnArgs a;
  float scale_log2;
};

__global__ void __launch_bounds__(kTT_Threads, 2) text_attn_tc_kernel(const __grid_constant__ TextTcParams P) {
  extern __shared__ uint8_t tt_smem_raw[];
  const uint32_t raw_u32 = ptx::smem_u32(tt_smem_raw);
  uint8_t* smem = tt_smem_raw + (((raw_u32 + 1023u) & ~1023u) - raw_u32);
  uint8_t* smem_q = smem;
  uint8_t* smem_k = smem + kTT_TileBytes;        // 2 stages
  uint8_t* smem_v = smem + 3 * kTT_TileBytes;    // 2 stages
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kTT_BarOff);
  uint64_t* q_full = bars;
  uint64_t* k_full = bars + 1;     // [2]
  uint64_t* v_full = bars + 3;     // [2]
  uint64_t* kv_empty = bars + 5;   // [2] P V of the block has completed: K and V of the stage are free
  uint64_t* s_full = bars + 7;
  uint64_t* p_full = bars + 8;     // 128 arrivals
  uint64_t* o_done = bars + 9;     // O += P V of the block has completed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);
  float* Bs = reinterpret_cast<float*>(smem + kTT_BiasOff);
  const TextAttnArgs& a = P.a;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.y, b = blockIdx.z, q0 = blockIdx.x * 128;
  const int k_end = a.causal ? min(a.S, q0 + 128) : a.S;
  const int n_kv = (k_end + 127) / 128;

  if (warp == 0 && lane == 0) ptx::prefetch_tmap(&P.tm);
  if (warp == 1 && lane == 0) {
    ptx::mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&k_full[i], 1);
      ptx::mbar_init(&v_full[i], 1);
      ptx::mbar_init(&kv_empty[i], 1);
    }
    ptx::mbar_init(s_full, 1);
    ptx::mbar_init(p_full, 128);
    ptx::mbar_init(o_done, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc<1>(tmem_slot, 256);
  if (a.rel_bias) {   // this head's bias row, already in the log2 domain
    const float* src = a.rel_bias + (long long)h * (2 * a.S - 1);
    for (int e = threadIdx.x; e < 2 * a.S - 1; e += kTT_Threads) Bs[e] = src[e] * 1.4426950408889634f;
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      ptx::mbar_arrive_expect_tx(q_full, kTT_TileBytes);
      ptx::tma_load_3d(&P.tm, q_full, smem_q, a.q_col0 + h * 64, q0, b);
      for (int j = 0; j < n_kv; ++j) {
        const int st = j & 1, ph = (j >> 1) & 1;
        ptx::mbar_wait(&kv_empty[st], ph ^ 1);
        ptx::mbar_arrive_expect_tx(&k_full[st], kTT_TileBytes);
        ptx::tma_load_3d(&P.tm, &k_full[st], smem_k + st * kTT_TileBytes, a.k_col0 + h * 64, j * 128, b);
        ptx::mbar_arrive_expect_tx(&v_full[st], kTT_TileBytes);
        ptx::tma_load_3d(&P.tm, &v_full[st], smem_v + st * kTT_TileBytes, a.v_col0 + h * 64, j * 128, b);
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc_qk = ptx::make_idesc_bf16(128, 128, 0, 0);  // A, B K-major
    constexpr uint32_t idesc_pv = ptx::make_idesc_bf16(128, 64, 0, 1);   // A (= P) from tensor memory, B (= V) MN-major
    const uint64_t q_desc = ptx::make_smem_desc_sw128(ptx::smem_u32(smem_q), 0, 1024);
    const uint64_t k_desc = ptx::make_smem_desc_sw128(ptx::smem_u32(smem_k), 0, 1024);
    const uint64_t v_desc = ptx::make_smem_desc_sw128(ptx::smem_u32(smem_v), kTT_TileBytes, 1024);
    constexpr uint32_t kTile16 = kTT_TileBytes >> 4;
    ptx::mbar_wait(q_full, 0);
    for (int j = 0; j < n_kv; ++j) {
      const int st = j & 1, ph = (j >> 1) & 1;
      ptx::mbar_wait(&k_full[st], ph);
      ptx::tc_fence_after();
      if (ptx::elect_one()) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)   // 16 of the 64 head-dim columns = 32 B inside the 128-byte swizzle row
          ptx::mma_bf16_ss<1>(tmem, q_desc + 2 * kk, k_desc + (uint64_t)(st * kTile16) + 2 * kk, idesc_qk, kk != 0 ? 1u : 0u);
        ptx::mma_commit(s_full);
      }
      __syncwarp();
      ptx::mbar_wait(p_full, j & 1);
      ptx::mbar_wait(&v_full[st], ph);
      ptx::tc_fence_after();
      if (ptx::elect_one()) {
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)   // 16 keys = 8 packed columns of P = two 8-row groups of V (2048 B)
          ptx::mma_bf16_ts(tmem + 128, tmem + kk * 8, v_desc + (uint64_t)(st * kTile16) + (uint64_t)(kk * 128), idesc_pv,
                           (j > 0 || kk > 0) ? 1u : 0u);
        ptx::mma_commit(&kv_empty[st]);
        ptx::mma_commit(o_done);
      }
      __syncwarp();
    }
  } else {
    const int quad = warp & 3;   // tensor-memory lane quadrant of this warp (warps 2, 3, 4, 5 -> 2, 3, 0, 1)
    const int row = quad * 32 + lane, qrow = q0 + row;
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t s_addr = tmem + lane_off, o_addr = tmem + lane_off + 128;
    const int bias_base = a.S - 1 - min(qrow, a.S - 1);   // Bs[key + bias_base] = bias(key - query)
    float m = -INFINITY, l = 0.f;
    for (int j = 0; j < n_kv; ++j) {
      ptx::mbar_wait(s_full, j & 1);
      ptx::tc_fence_after();
      uint32_t s0[32], s1[32], s2[32], s3[32];
      ptx::tmem_ld_32x32b_x32(s_addr, s0);
      ptx::tmem_ld_32x32b_x32(s_addr + 32, s1);
      ptx::tmem_ld_32x32b_x32(s_addr + 64, s2);
      ptx::tmem_ld_32x32b_x32(s_addr + 96, s3);
      ptx::tmem_ld_wait();
      // scores -> log2-domain logits (bias, masks), running maximum on four independent chains (one softmax warp per
      // scheduler: every dependent chain is exposed latency)
      float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
      const bool full = j * 128 + 128 <= a.S && !(a.causal && j * 128 + 127 > q0);   // no key of this block is masked
      auto prep = [&](uint32_t (&sv)[32], int c0) {
        const float* bs = Bs + j * 128 + c0 + bias_base;
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          float x = __uint_as_float(sv[i]) * P.scale_log2;
          if (full) {
            if (a.rel_bias) x += bs[i];
          } else {
            const int key = j * 128 + c0 + i;
            if (a.rel_bias && key < a.S) x += bs[i];
            if (key >= a.S || (a.causal && key > qrow)) x = -INFINITY;
          }
          sv[i] = __float_as_uint(x);
          mx4[i & 3] = fmaxf(mx4[i & 3], x);
        }
      };
      prep(s0, 0);
      prep(s1, 32);
      prep(s2, 64);
      prep(s3, 96);
      const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
      const float m_new = fmaxf(m, mx);     // finite: every row sees at least one unmasked key in every block it visits
      const float corr = ptx::ex2_approx(m - m_new);
      if (j > 0 && __any_sync(0xffffffffu, m_new != m)) {   // the maximum moved for some row of this warp: rescale O
        ptx::mbar_wait(o_done, (j - 1) & 1);
        ptx::tc_fence_after();
#pragma unroll 1
        for (int ch = 0; ch < 4; ++ch) {
          uint32_t r[16];
          ptx::tmem_ld_32x32b_x16(o_addr + ch * 16, r);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * corr);
          ptx::tmem_st_32x32b_x16(o_addr + ch * 16, r);
        }
      }
      float2 ps2[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
      const float2 nm2 = make_float2(-m_new, -m_new);
      auto expo = [&](const uint32_t (&sv)[32], int col) {   // P = 2^(x - m) -> bf16 pairs over the score columns
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float2 x = __fadd2_rn(make_float2(__uint_as_float(sv[2 * i]), __uint_as_float(sv[2 * i + 1])), nm2);
          float2 e;
          e.x = ptx::ex2_approx(x.x);
          e.y = ptx::ex2_approx(x.y);
          ps2[i & 1] = __fadd2_rn(ps2[i & 1], e);
          pk[i] = ptx::pack_bf16x2(e.x, e.y);
        }
        ptx::tmem_st_32x32b_x16(s_addr + col, pk);
      };
      expo(s0, 0);
      expo(s1, 16);
      expo(s2, 32);
      expo(s3, 48);
      l = l * corr + ((ps2[0].x + ps2[0].y) + (ps2[1].x + ps2[1].y));
      m = m_new;
      ptx::tmem_st_wait();
      ptx::tc_fence_before();
      ptx::mbar_arrive(p_full);
    }
    ptx::mbar_wait(o_done, (n_kv - 1) & 1);
    ptx::tc_fence_after();
    const float inv = 1.f / l;
    bf16* dst = a.out + (long long)b * a.out_batch_stride + (long long)qrow * a.out_ld + a.out_col0 + h * 64;
#pragma unroll 1
    for (int ch = 0; ch < 2; ++ch) {
      uint32_t r[32];
      ptx::tmem_ld_32x32b_x32(o_addr + ch * 32, r);
      ptx::tmem_ld_wait();
      if (qrow < a.S) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          uint4 u;
          u.x = ptx::pack_bf16x2(__uint_as_float(r[8 * i + 0]) * inv, __uint_as_float(r[8 * i + 1]) * inv);
          u.y = ptx::pack_bf16x2(__uint_as_float(r[8 * i + 2]) * inv, __uint_as_float(r[8 * i + 3]) * inv);
          u.z = ptx::pack_bf16x2(__uint_as_float(r[8 * i + 4]) * inv, __uint_as_float(r[8 * i + 5]) * inv);
          u.w = ptx::pack_bf16x2(__uint_as_float(r[8 * i + 6]) * inv, __uint_as_float(r[8 * i + 7]) * inv);
          *reinterpret_cast<uint4*>(dst + ch * 32 + i * 8) = u;
        }
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc<1>(tmem, 256);
}

// ---- gated activation: kind 0: out = in[:, :F] * in[:, F:2F]; kind 1: out = quick_gelu(in[:, :F]) = x * sigmoid(1.702 x)
__global__ void __launch_bounds__(256) glu_act_kernel(const bf16* __restrict__ in, long long in_ld, bf16* __restrict__ out,
                                                      long long out_ld, int F, int kind, long long total_vecs) {
  const int vecs = F / 8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total_vecs;
       i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % vecs);
    const long long r = i / vecs;
    float x[8], y[8], o[8];
    ldvec(in + r * in_ld + v * 8, x);
    if (kind == 0) {
      ldvec(in + r * in_ld + F + v * 8, y);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = x[j] * y[j];
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = x[j] / (1.f + __expf(-1.702f * x[j]));
    }
    stvec(out + r * out_ld + v * 8, o);
  }
}

// ---- out[i, :] = table[ids[i], :] (+ pos_table[i % S, :]); ids int64; out-of-range ids raise a flag instead of reading
__global__ void __launch_bounds__(128) embedding_kernel(const bf16* __restrict__ table, long long vocab, int D,
                                                        const long long* __restrict__ ids, const bf16* __restrict__ pos,
                                                        int S, bf16* __restrict__ out, int* __restrict__ bad) {
  const long long i = blockIdx.x;
  const long long id = ids[i];
  if (id < 0 || id >= vocab) {
    if (threadIdx.x == 0) atomicExch(bad, 1);
    return;
  }
  const bf16* src = table + id * D;
  const bf16* ps = pos ? pos + (long long)(i % S) * D : nullptr;
  for (int v = threadIdx.x; v < D / 8; v += blockDim.x) {
    float t[8];
    ldvec(src + v * 8, t);
    if (ps) {
      float p[8];
      ldvec(ps + v * 8, p);
#pragma unroll
      for (int j = 0; j < 8; ++j) t[j] += p[j];
    }
    stvec(out + i * D + v * 8, t);
  }
}

int sm_count_t() { return rt::device_sm_count(); }

}  // namespace
}  // namespace rt

using namespace rt;

extern "C" {

int rt_norm_rows(const void* x, int64_t x_ld, void* out, int64_t out_ld, int64_t rows, int D, const void* weight,
                 const void* bias, float eps, int subtract_mean, void* stream) {
  return guarded([&] {
    RT_REQUIRE(x && out && weight && rows >= 1, "norm_rows: null argument");
    RT_REQUIRE(D >= 8 && D % 8 == 0 && D <= 4096 && x_ld % 8 == 0 && out_ld % 8 == 0, "norm_rows: D must be a multiple of 8, <= 4096");
    cudaStream_t s = (cudaStream_t)stream;
    ProfScope ps(PROF_ELEM, 2.0 * rows * (double)D * 2, s);
    if (subtract_mean)
      norm_rows_kernel<true><<<(unsigned)rows, 128, 0, s>>>((const bf16*)x, x_ld, (bf16*)out, out_ld, D, (const bf16*)weight,
                                                            (const bf16*)bias, eps);
    else
      norm_rows_kernel<false><<<(unsigned)rows, 128, 0, s>>>((const bf16*)x, x_ld, (bf16*)out, out_ld, D, (const bf16*)weight,
                                                             (const bf16*)bias, eps);
    RT_POST_LAUNCH();
  });
}

int rt_text_attention(const void* qkv, int64_t batch_stride, int ld, int q_col0, int k_col0, int v_col0, void* out,
                      int64_t out_batch_stride, int out_ld, int out_col0, int batch, int S, int heads, int hd,
                      float scale, const float* rel_bias, int causal, void* stream) {
  return guarded([&] {
    RT_REQUIRE(qkv && out && batch >= 1 && S >= 1 && heads >= 1, "text_attention: bad argument");
    RT_REQUIRE(hd == 64, "text_attention: head_dim must be 64 (T5-XXL, CLIP-L)");
    RT_REQUIRE(S <= 4096, "text_attention: S <= 4096");
    TextAttnArgs a{};
    a.qkv = (const bf16*)qkv; a.batch_stride = batch_stride; a.ld = ld;
    a.q_col0 = q_col0; a.k_col0 = k_col0; a.v_col0 = v_col0;
    a.out = (bf16*)out; a.out_batch_stride = out_batch_stride; a.out_ld = out_ld; a.out_col0 = out_col0;
    a.S = S; a.heads = heads; a.scale = scale; a.rel_bias = rel_bias; a.causal = causal;
    cudaStream_t s = (cudaStream_t)stream;
    ProfScope ps(PROF_ATTN_TEXT, 4.0 * batch * heads * (double)S * S * hd * (causal ? 0.5 : 1.0), s);
    const bool aligned = ld % 8 == 0 && q_col0 % 8 == 0 && k_col0 % 8 == 0 && v_col0 % 8 == 0 && out_ld % 8 == 0 &&
                         out_col0 % 8 == 0 && batch_stride % 8 == 0 && out_batch_stride % 8 == 0 &&
                         (reinterpret_cast<uintptr_t>(qkv) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0;
    if (!get_option("text_attn_simt") && !get_option("text_attn_mma") && aligned) {
      // tcgen05 / TMEM / TMA form
      TextTcParams T{};
      T.a = a;
      T.scale_log2 = scale * 1.4426950408889634f;
      uint64_t dims[3] = {(uint64_t)ld, (uint64_t)S, (uint64_t)batch};
      uint64_t strides[2] = {(uint64_t)ld * 2, (uint64_t)batch_stride * 2};
      uint32_t box[3] = {64, 128, 1};
      encode_tmap_bf16(&T.tm, qkv, 3, dims, strides, box);
      const size_t smem = (size_t)kTT_BiasOff + (size_t)(rel_bias ? 2 * S - 1 : 0) * sizeof(float) + 1024 + 16;
      RT_REQUIRE(smem <= 227 * 1024, "text_attention: the relative-bias row does not fit into shared memory");
      static PerDeviceOnce attr;
      if (attr.first())
        RT_CHECK_CUDA(cudaFuncSetAttribute(text_attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      dim3 grid((S + 127) / 128, heads, batch);
      text_attn_tc_kernel<<<grid, kTT_Threads, smem, s>>>(T);
    } else if (!get_option("text_attn_simt")) {
      RT_REQUIRE(aligned, "text_attention: pointers / strides / column offsets must be multiples of 8 elements");
      const size_t smem = kTA_SmemFixed + (size_t)(rel_bias ? 2 * S - 1 : 0) * sizeof(float);
      static PerDeviceOnce attr;
      if (attr.first())
        RT_CHECK_CUDA(cudaFuncSetAttribute(text_attn_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024));
      dim3 grid((S + kTA_QT - 1) / kTA_QT, heads, batch);
      text_attn_mma_kernel<<<grid, 128, smem, s>>>(a);
    } else {
      const size_t smem = (size_t)(32 * 65 + 32 * 64 + 32 * 64 + (rel_bias ? 2 * S - 1 : 0)) * sizeof(float);
      static PerDeviceOnce attr;
      if (attr.first())
        RT_CHECK_CUDA(cudaFuncSetAttribute(text_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
      dim3 grid((S + 31) / 32, heads, batch);
      text_attn_kernel<<<grid, 256, smem, s>>>(a);
    }
    RT_POST_LAUNCH();
  });
}

int rt_glu_act(int kind, const void* in, int64_t in_ld, void* out, int64_t out_ld, int64_t rows, int F, void* stream) {
  return guarded([&] {
    RT_REQUIRE(in && out && rows >= 1 && F >= 8 && F % 8 == 0 && in_ld % 8 == 0 && out_ld % 8 == 0 && (kind == 0 || kind == 1),
               "glu_act: bad argument");
    const long long total = rows * (long long)(F / 8);
    cudaStream_t s = (cudaStream_t)stream;
    ProfScope ps(PROF_ELEM, (double)total * 16 * (kind == 0 ? 3 : 2), s);
    long long blocks = (total + 255) / 256;
    const long long cap = (long long)sm_count_t() * 16;
    if (blocks > cap) blocks = cap;
    glu_act_kernel<<<(unsigned)blocks, 256, 0, s>>>((const bf16*)in, in_ld, (bf16*)out, out_ld, F, kind, total);
    RT_POST_LAUNCH();
  });
}

int rt_embedding(const void* table, int64_t vocab, int D, const int64_t* ids, int64_t n, const void* pos_table, int S,
                 void* out, int* bad_flag, void* stream) {
  return guarded([&] {
    RT_REQUIRE(table && ids && out && bad_flag && n >= 1 && vocab >= 1 && D >= 8 && D % 8 == 0 && S >= 1,
               "embedding: bad argument");
    cudaStream_t s = (cudaStream_t)stream;
    ProfScope ps(PROF_ELEM, (double)n * D * 2 * (pos_table ? 3 : 2), s);
    embedding_kernel<<<(unsigned)n, 128, 0, s>>>((const bf16*)table, vocab, D, (const long long*)ids, (const bf16*)pos_table,
                                                 S, (bf16*)out, bad_flag);
    RT_POST_LAUNCH();
  });
}

}  // extern "C"
