// HBM-bound kernels of the hot path: LayerNorm+modulation, grouped GEMV (AdaLN / time-text MLPs),
// RoPE table, QK-RMSNorm+RoPE (unfused form), Euler step, CFG combine, regional mask, glyph blend.
// All are coalesced 16-byte-per-thread kernels; grids are sized in multiples of the SM count.
#include <cmath>

#include "dtype_utils.cuh"
#include "rt_internal.h"

namespace rt {

long long g_launch_count = 0;

static int sm_count() { return device_sm_count(); }

// ------------------------------------------------------------------------------------------------
// LayerNorm (eps 1e-6, no affine) + (1 + scale) * xn + shift.  One warp per row, row cached in
// registers (up to 16 16-byte vectors per lane = D <= 4096 bf16 / 2048 fp32), else 3 passes over L1.
// Reference: diffusers AdaLayerNormZero / AdaLayerNormZeroSingle / AdaLayerNormContinuous, reached
// from RepText/controlnet_flux.py:343-348 via FluxTransformerBlock.
// ------------------------------------------------------------------------------------------------
constexpr int kLnMaxGroups = 2;
struct LnGroups {
  int n;
  LnModGroup g[kLnMaxGroups];
};

template <typename T, int MAXV>
__global__ void __launch_bounds__(256) ln_mod_kernel(const T* __restrict__ x, long long x_bs, int x_ld,
                                                     T* __restrict__ out, long long o_bs, int o_ld, int batch,
                                                     int rows_total, int D, LnGroups groups) {
  constexpr int N = VecT<T>::N;
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  const int nvec = D / N;  // D % N == 0 checked on the host
  for (long long wrow = (long long)blockIdx.x * warps_per_block + (threadIdx.x >> 5);
       wrow < (long long)batch * rows_total; wrow += (long long)gridDim.x * warps_per_block) {
    const int b = (int)(wrow / rows_total);
    const int lr = (int)(wrow % rows_total);
    // map the linear row index onto the groups' row ranges
    int r = -1, gi = 0, acc = 0;
#pragma unroll
    for (int k = 0; k < kLnMaxGroups; ++k) {
      if (k < groups.n) {
        int len = groups.g[k].row_end - groups.g[k].row_begin;
        if (r < 0 && lr < acc + len) {
          r = groups.g[k].row_begin + (lr - acc);
          gi = k;
        }
        acc += len;
      }
    }
    const LnModGroup& G = groups.g[gi];
    const T* xr = x + (long long)b * x_bs + (long long)r * x_ld;
    T* orow = out + (long long)b * o_bs + (long long)r * o_ld;
    const float* sh = G.shift + (long long)b * G.ld;
    const float* sc = G.scale + (long long)b * G.ld;

    float v[MAXV][N];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      int c = lane + 32 * i;
      if (c < nvec) {
        ldvec(xr + c * N, v[i]);
#pragma unroll
        for (int j = 0; j < N; ++j) s += v[i][j];
      }
    }
    s = warp_sum(s);
    const float mean = s / (float)D;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      int c = lane + 32 * i;
      if (c < nvec) {
#pragma unroll
        for (int j = 0; j < N; ++j) {
          float d = v[i][j] - mean;
          q += d * d;
        }
      }
    }
    q = warp_sum(q);
    const float rstd = rsqrtf(q / (float)D + 1e-6f);
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
      int c = lane + 32 * i;
      if (c < nvec) {
        float o[N];
#pragma unroll
        for (int j4 = 0; j4 < N; j4 += 4) {
          float4 a = *reinterpret_cast<const float4*>(sc + c * N + j4);
          float4 h = *reinterpret_cast<const float4*>(sh + c * N + j4);
          o[j4 + 0] = (v[i][j4 + 0] - mean) * rstd * (1.f + a.x) + h.x;
          o[j4 + 1] = (v[i][j4 + 1] - mean) * rstd * (1.f + a.y) + h.y;
          o[j4 + 2] = (v[i][j4 + 2] - mean) * rstd * (1.f + a.z) + h.z;
          o[j4 + 3] = (v[i][j4 + 3] - mean) * rstd * (1.f + a.w) + h.w;
        }
        stvec(orow + c * N, o);
      }
    }
  }
}


// CTA form for rows that fit one 16-byte vector per thread (D <= 8192 bf16 / 4096 fp32): thread t owns columns
// [t*N, t*N+N) of EVERY row its CTA processes, so the (1 + scale) / shift vectors live in registers and are read
// once per CTA instead of once per row (per row they are 4x the bytes of the row itself in bf16).  The kernel is
// written for instruction count (the warp-per-row form issues ~290 instructions per thread and row and is
// issue-bound, ncu r1): packed fp32x2 math, a 4-row butterfly for the warp reductions, 13-instruction cross-warp
// reductions, and the next 4 rows are already in flight while the current 4 are reduced.  Mean and variance are two
// passes over the registers (like torch's LayerNorm).
__device__ __forceinline__ float2 f2(float a, float b) { return make_float2(a, b); }

// sums of 4 independent values over the warp; every lane of lane-group g = lane >> 3 ends with the total of value g
__device__ __forceinline__ float warp_sum4(const float (&p)[4], int lane) {
  const bool hi16 = lane & 16, hi8 = lane & 8;
  const float s0 = hi16 ? p[0] : p[2], s1 = hi16 ? p[1] : p[3];
  const float k0 = hi16 ? p[2] : p[0], k1 = hi16 ? p[3] : p[1];
  const float q0 = k0 + __shfl_xor_sync(0xffffffffu, s0, 16);
  const float q1 = k1 + __shfl_xor_sync(0xffffffffu, s1, 16);
  const float send = hi8 ? q0 : q1, keep = hi8 ? q1 : q0;
  float t = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  t += __shfl_xor_sync(0xffffffffu, t, 4);
  t += __shfl_xor_sync(0xffffffffu, t, 2);
  t += __shfl_xor_sync(0xffffffffu, t, 1);
  return t;
}

template <typename T>
struct RawVec;
template <>
struct RawVec<bf16> {
  uint4 u;
  __device__ __forceinline__ void load(const bf16* p) { u = *reinterpret_cast<const uint4*>(p); }
  __device__ __forceinline__ void zero() { u = make_uint4(0, 0, 0, 0); }
  __device__ __forceinline__ void unpack(float2 (&v)[4]) const {
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = f2(__uint_as_float(w[i] << 16), __uint_as_float(w[i] & 0xFFFF0000u));
  }
  static __device__ __forceinline__ void store(bf16* p, const float2 (&o)[4]) {
    uint4 t;
    __nv_bfloat162 h0 = __float22bfloat162_rn(o[0]), h1 = __float22bfloat162_rn(o[1]);
    __nv_bfloat162 h2 = __float22bfloat162_rn(o[2]), h3 = __float22bfloat162_rn(o[3]);
    t.x = *reinterpret_cast<uint32_t*>(&h0); t.y = *reinterpret_cast<uint32_t*>(&h1);
    t.z = *reinterpret_cast<uint32_t*>(&h2); t.w = *reinterpret_cast<uint32_t*>(&h3);
    *reinterpret_cast<uint4*>(p) = t;
  }
};
template <>
struct RawVec<float> {
  float4 u;
  __device__ __forceinline__ void load(const float* p) { u = *reinterpret_cast<const float4*>(p); }
  __device__ __forceinline__ void zero() { u = make_float4(0.f, 0.f, 0.f, 0.f); }
  __device__ __forceinline__ void unpack(float2 (&v)[2]) const { v[0] = f2(u.x, u.y); v[1] = f2(u.z, u.w); }
  static __device__ __forceinline__ void store(float* p, const float2 (&o)[2]) {
    *reinterpret_cast<float4*>(p) = make_float4(o[0].x, o[0].y, o[1].x, o[1].y);
  }
};

template <typename T, int kMaxThreads, int kMinBlocks>
__global__ void __launch_bounds__(kMaxThreads, kMinBlocks) ln_mod_cta_kernel(const T* __restrict__ x, long long x_bs, int x_ld,
                                                                 T* __restrict__ out, long long o_bs, int o_ld,
                                                                 int batch, int rows_total, int D, LnGroups groups,
                                                                 int rows_per_cta) {
  constexpr int N = VecT<T>::N, H = N / 2, kR = 4;
  __shared__ float red[2][kR][32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  const bool active = tid * N < D;
  const int lin_begin = blockIdx.x * rows_per_cta;
  const int lin_end = min(lin_begin + rows_per_cta, batch * rows_total);
  const float inv_d = 1.f / (float)D;

  // position of a linear row index: batch b, group gi, row r, rows left in this (batch, group)
  int b = 0, r = 0, left = 0, g_ld = 0;
  const float *g_scale = nullptr, *g_shift = nullptr;
  auto locate = [&](int lin) {
    b = lin / rows_total;
    int lr = lin - b * rows_total;
    bool found = false;
#pragma unroll
    for (int k = 0; k < kLnMaxGroups; ++k) {  // static indices only: `groups` stays in the constant bank
      const int len = groups.g[k].row_end - groups.g[k].row_begin;
      if (!found && k < groups.n) {
        if (lr < len) {
          found = true;
          r = groups.g[k].row_begin + lr;
          left = len - lr;
          g_scale = groups.g[k].scale; g_shift = groups.g[k].shift; g_ld = groups.g[k].ld;
        } else {
          lr -= len;
        }
      }
    }
  };
  float2 sc1[H], sh[H];
  auto load_mod = [&]() {
    if (active) {
      const float* scp = g_scale + (long long)b * g_ld + tid * N;
      const float* shp = g_shift + (long long)b * g_ld + tid * N;
#pragma unroll
      for (int j = 0; j < H; j += 2) {
        const float4 a = *reinterpret_cast<const float4*>(scp + 2 * j);
        const float4 h = *reinterpret_cast<const float4*>(shp + 2 * j);
        sc1[j] = f2(1.f + a.x, 1.f + a.y); sc1[j + 1] = f2(1.f + a.z, 1.f + a.w);
        sh[j] = f2(h.x, h.y); sh[j + 1] = f2(h.z, h.w);
      }
    }
  };
  if (lin_begin >= lin_end) return;
  locate(lin_begin);
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");  // programmatic dependent launch (see ptx_sm100.cuh)
  asm volatile("griddepcontrol.wait;" ::: "memory");
  load_mod();

  RawVec<T> cur[kR], nxt[kR];
  int n = min(kR, min(lin_end - lin_begin, left));
  {
    const T* xr = x + (long long)b * x_bs + (long long)r * x_ld + tid * N;
#pragma unroll
    for (int i = 0; i < kR; ++i) {
      if (active && i < n) cur[i].load(xr + (long long)i * x_ld); else cur[i].zero();
    }
  }
  for (int lin = lin_begin; lin < lin_end;) {
    // ---- where the NEXT iteration reads; issue its loads now
    const int b0 = b, r0 = r, n0 = n;
    const int lin_next = lin + n0;
    int n_next = 0;
    bool mod_changes = false;
    if (lin_next < lin_end) {
      if (left > n0) { r += n0; left -= n0; } else { locate(lin_next); mod_changes = true; }
      n_next = min(kR, min(lin_end - lin_next, left));
      const T* xr = x + (long long)b * x_bs + (long long)r * x_ld + tid * N;
#pragma unroll
      for (int i = 0; i < kR; ++i) {
        if (active && i < n_next) nxt[i].load(xr + (long long)i * x_ld); else nxt[i].zero();
      }
    }
    // ---- mean
    float2 v[kR][H];
    float part[kR];
#pragma unroll
    for (int i = 0; i < kR; ++i) {
      cur[i].unpack(v[i]);
      float2 a = v[i][0];
#pragma unroll
      for (int j = 1; j < H; ++j) a = __fadd2_rn(a, v[i][j]);
      part[i] = a.x + a.y;
    }
    float t = warp_sum4(part, lane);
    if ((lane & 7) == 0) red[0][lane >> 3][warp] = t;
    __syncthreads();
    {
      float a = 0.f;
      for (int w = lane & 7; w < nwarps; w += 8) a += red[0][lane >> 3][w];
      a += __shfl_xor_sync(0xffffffffu, a, 4);
      a += __shfl_xor_sync(0xffffffffu, a, 2);
      a += __shfl_xor_sync(0xffffffffu, a, 1);
      t = a;
    }
    float mean[kR];
#pragma unroll
    for (int i = 0; i < kR; ++i) mean[i] = __shfl_sync(0xffffffffu, t, i * 8) * inv_d;
    // ---- variance (v becomes x - mean)
#pragma unroll
    for (int i = 0; i < kR; ++i) {
      const float2 nm = f2(-mean[i], -mean[i]);
      float2 q = f2(0.f, 0.f);
#pragma unroll
      for (int j = 0; j < H; ++j) {
        v[i][j] = __fadd2_rn(v[i][j], nm);
        q = __ffma2_rn(v[i][j], v[i][j], q);
      }
      part[i] = active ? q.x + q.y : 0.f;
    }
    t = warp_sum4(part, lane);
    if ((lane & 7) == 0) red[1][lane >> 3][warp] = t;
    __syncthreads();
    {
      float a = 0.f;
      for (int w = lane & 7; w < nwarps; w += 8) a += red[1][lane >> 3][w];
      a += __shfl_xor_sync(0xffffffffu, a, 4);
      a += __shfl_xor_sync(0xffffffffu, a, 2);
      a += __shfl_xor_sync(0xffffffffu, a, 1);
      t = a;
    }
    // ---- normalise, modulate, store
    T* orow = out + (long long)b0 * o_bs + (long long)r0 * o_ld + tid * N;
#pragma unroll
    for (int i = 0; i < kR; ++i) {
      const float rstd = rsqrtf(__shfl_sync(0xffffffffu, t, i * 8) * inv_d + 1e-6f);
      if (active && i < n0) {
        const float2 rs = f2(rstd, rstd);
        float2 o[H];
#pragma unroll
        for (int j = 0; j < H; ++j) o[j] = __ffma2_rn(v[i][j], __fmul2_rn(rs, sc1[j]), sh[j]);
        RawVec<T>::store(orow + (long long)i * o_ld, o);
      }
    }
    if (mod_changes) load_mod();   // the next rows belong to another (batch, group)
#pragma unroll
    for (int i = 0; i < kR; ++i) cur[i] = nxt[i];
    n = n_next;
    lin = lin_next;
  }
}


// v3: one warp per PAIR of rows, no block-wide reductions.  The CTA stages (1 + scale) and shift for its (batch, group)
// in shared memory once (fp32, 8 D bytes), then each of its warps takes two rows: all 2 x kV 16-byte row loads are
// issued up front, mean and variance are warp-shuffle reductions over registers (two-pass), and the modulation
// vectors are read from shared memory once for both rows.  ~470 instructions per lane and row, no __syncthreads in
// the row loop.  bf16, D a multiple of 256 and <= 4096.
struct LnSeg {  // one (batch, group) segment of rows and the CTAs that process it
  int batch, row_begin, row_end, cta_begin;
  const float *scale, *shift;
};
constexpr int kLnMaxSegs = 16;
struct LnSegs {
  int n, rows_per_cta;
  LnSeg s[kLnMaxSegs];
};

template <int kV>
__global__ void __launch_bounds__(256, 1) ln_mod_rows2_kernel(const bf16* __restrict__ x, long long x_bs, int x_ld,
                                                              bf16* __restrict__ out, long long o_bs, int o_ld, int D,
                                                              const LnSegs segs) {
  extern __shared__ float ln_smem[];  // [D] 1 + scale | [D] shift
  float* s_sc = ln_smem;
  float* s_sh = ln_smem + D;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // which segment does this CTA belong to?  (static indices: the table stays in the constant bank)
  int b = 0, r0 = 0, r_end = 0;
  const float *gsc = nullptr, *gsh = nullptr;
#pragma unroll
  for (int i = 0; i < kLnMaxSegs; ++i) {
    if (i < segs.n && (int)blockIdx.x >= segs.s[i].cta_begin) {
      b = segs.s[i].batch;
      r0 = segs.s[i].row_begin + ((int)blockIdx.x - segs.s[i].cta_begin) * segs.rows_per_cta;
      r_end = segs.s[i].row_end;
      gsc = segs.s[i].scale;
      gsh = segs.s[i].shift;
    }
  }
  r_end = min(r_end, r0 + segs.rows_per_cta);
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");  // programmatic dependent launch (see ptx_sm100.cuh)
  asm volatile("griddepcontrol.wait;" ::: "memory");
  for (int i = threadIdx.x * 4; i < D; i += blockDim.x * 4) {
    const float4 a = *reinterpret_cast<const float4*>(gsc + i);
    const float4 h = *reinterpret_cast<const float4*>(gsh + i);
    *reinterpret_cast<float4*>(s_sc + i) = make_float4(1.f + a.x, 1.f + a.y, 1.f + a.z, 1.f + a.w);
    *reinterpret_cast<float4*>(s_sh + i) = h;
  }
  __syncthreads();
  const float inv_d = 1.f / (float)D;
  for (int r = r0 + warp * 2; r < r_end; r += (blockDim.x >> 5) * 2) {
    const bool two = r + 1 < r_end;
    const bf16* xa = x + (long long)b * x_bs + (long long)r * x_ld + lane * 8;
    const bf16* xb = xa + (two ? x_ld : 0);
    uint4 ra[kV], rb[kV];
#pragma unroll
    for (int i = 0; i < kV; ++i) {
      ra[i] = *reinterpret_cast<const uint4*>(xa + i * 256);
      rb[i] = *reinterpret_cast<const uint4*>(xb + i * 256);
    }
    auto lo = [](uint32_t w) { return __uint_as_float(w << 16); };
    auto hi = [](uint32_t w) { return __uint_as_float(w & 0xFFFF0000u); };
    // four independent partial sums per row, pieces x / y / z / w: the same order as ln_mod_row1_kernel (same bits)
    float2 sa[4] = {f2(0.f, 0.f), f2(0.f, 0.f), f2(0.f, 0.f), f2(0.f, 0.f)};
    float2 sb[4] = {f2(0.f, 0.f), f2(0.f, 0.f), f2(0.f, 0.f), f2(0.f, 0.f)};
#pragma unroll
    for (int i = 0; i < kV; ++i) {
      const uint32_t wa[4] = {ra[i].x, ra[i].y, ra[i].z, ra[i].w}, wb[4] = {rb[i].x, rb[i].y, rb[i].z, rb[i].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        sa[j] = __fadd2_rn(sa[j], f2(lo(wa[j]), hi(wa[j])));
        sb[j] = __fadd2_rn(sb[j], f2(lo(wb[j]), hi(wb[j])));
      }
    }
    const float2 sta = __fadd2_rn(__fadd2_rn(sa[0], sa[1]), __fadd2_rn(sa[2], sa[3]));
    const float2 stb = __fadd2_rn(__fadd2_rn(sb[0], sb[1]), __fadd2_rn(sb[2], sb[3]));
    const float mean_a = warp_sum(sta.x + sta.y) * inv_d, mean_b = warp_sum(stb.x + stb.y) * inv_d;
    const float2 na = f2(-mean_a, -mean_a), nb = f2(-mean_b, -mean_b);
    float2 qa[4] = {f2(0.f, 0.f), f2(0.f, 0.f), f2(0.f, 0.f), f2(0.f, 0.f)};
    float2 qb[4] = {f2(0.f, 0.f), f2(0.f, 0.f), f2(0.f, 0.f), f2(0.f, 0.f)};
#pragma unroll
    for (int i = 0; i < kV; ++i) {
      const uint32_t wa[4] = {ra[i].x, ra[i].y, ra[i].z, ra[i].w}, wb[4] = {rb[i].x, rb[i].y, rb[i].z, rb[i].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 da = __fadd2_rn(f2(lo(wa[j]), hi(wa[j])), na), db = __fadd2_rn(f2(lo(wb[j]), hi(wb[j])), nb);
        qa[j] = __ffma2_rn(da, da, qa[j]);
        qb[j] = __ffma2_rn(db, db, qb[j]);
      }
    }
    const float2 qta = __fadd2_rn(__fadd2_rn(qa[0], qa[1]), __fadd2_rn(qa[2], qa[3]));
    const float2 qtb = __fadd2_rn(__fadd2_rn(qb[0], qb[1]), __fadd2_rn(qb[2], qb[3]));
    const float rstd_a = rsqrtf(warp_sum(qta.x + qta.y) * inv_d + 1e-6f);
    const float rstd_b = rsqrtf(warp_sum(qtb.x + qtb.y) * inv_d + 1e-6f);
    const float2 rsa = f2(rstd_a, rstd_a), rsb = f2(rstd_b, rstd_b);
    bf16* oa = out + (long long)b * o_bs + (long long)r * o_ld + lane * 8;
    bf16* ob = oa + o_ld;
#pragma unroll
    for (int i = 0; i < kV; ++i) {
      const int c = i * 256 + lane * 8;
      const float4 c0 = *reinterpret_cast<const float4*>(s_sc + c), c1 = *reinterpret_cast<const float4*>(s_sc + c + 4);
      const float4 h0 = *reinterpret_cast<const float4*>(s_sh + c), h1 = *reinterpret_cast<const float4*>(s_sh + c + 4);
      const float2 sc[4] = {f2(c0.x, c0.y), f2(c0.z, c0.w), f2(c1.x, c1.y), f2(c1.z, c1.w)};
      const float2 sh[4] = {f2(h0.x, h0.y), f2(h0.z, h0.w), f2(h1.x, h1.y), f2(h1.z, h1.w)};
      const uint32_t wa[4] = {ra[i].x, ra[i].y, ra[i].z, ra[i].w}, wb[4] = {rb[i].x, rb[i].y, rb[i].z, rb[i].w};
      uint32_t pa[4], pb[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 da = __fadd2_rn(f2(lo(wa[j]), hi(wa[j])), na), db = __fadd2_rn(f2(lo(wb[j]), hi(wb[j])), nb);
        const float2 ya = __ffma2_rn(da, __fmul2_rn(rsa, sc[j]), sh[j]);
        const float2 yb = __ffma2_rn(db, __fmul2_rn(rsb, sc[j]), sh[j]);
        __nv_bfloat162 ha = __float22bfloat162_rn(ya), hb = __float22bfloat162_rn(yb);
        pa[j] = *reinterpret_cast<uint32_t*>(&ha);
        pb[j] = *reinterpret_cast<uint32_t*>(&hb);
      }
      *reinterpret_cast<uint4*>(oa + i * 256) = make_uint4(pa[0], pa[1], pa[2], pa[3]);
      if (two) *reinterpret_cast<uint4*>(ob + i * 256) = make_uint4(pb[0], pb[1], pb[2], pb[3]);
    }
  }
}

template <int kV>
static void launch_ln_rows2(const void* x, long long x_bs, int x_ld, void* out, long long o_bs, int o_ld, int D,
                            const LnSegs& S, int grid, cudaStream_t stream) {
  const int smem = 2 * D * (int)sizeof(float);
  static PerDeviceOnce attr_set;
  if (attr_set.first()) {
    RT_CHECK_CUDA(cudaFuncSetAttribute(ln_mod_rows2_kernel<kV>, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 4096 * 4));
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(256);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = get_option("no_pdl") ? 0 : 1;
  RT_CHECK_CUDA(cudaLaunchKernelEx(&cfg, ln_mod_rows2_kernel<kV>, (const bf16*)x, x_bs, x_ld, (bf16*)out, o_bs, o_ld, D, S));
  count_launch();
}

// v4: one warp per ROW, 16 warps per CTA (v3: a PAIR of rows per warp, 8 warps).  ncu on v3 (profiles/
// r2_ncu_full_step_kernels.txt): 11 % warp occupancy, one instruction issued per warp every 10 cycles - two warps per
// scheduler, each walking a 940-instruction chain whose sums hang on ONE accumulator - so the kernel is bound by
// dependent-issue latency, not by bytes.  Here a scheduler has four warps with half the chain each, and the two
// reductions run on four independent partial sums.  Same staging of the modulation vectors, same one wave of CTAs.
template <int kV>
__global__ void __launch_bounds__(512, 1) ln_mod_row1_kernel(const bf16* __restrict__ x, long long x_bs, int x_ld,
                                                             bf16* __restrict__ out, long long o_bs, int o_ld, int D,
                                                             const LnSegs segs) {
  extern __shared__ float ln_smem[];  // [D] 1 + scale | [D] shift
  float* s_sc = ln_smem;
  float* s_sh = ln_smem + D;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int b = 0, r0 = 0, r_end = 0;
  const float *gsc = nullptr, *gsh = nullptr;
#pragma unroll
  for (int i = 0; i < kLnMaxSegs; ++i) {
    if (i < segs.n && (int)blockIdx.x >= segs.s[i].cta_begin) {
      b = segs.s[i].batch;
      r0 = segs.s[i].row_begin + ((int)blockIdx.x - segs.s[i].cta_begin) * segs.rows_per_cta;
      r_end = segs.s[i].row_end;
      gsc = segs.s[i].scale;
      gsh = segs.s[i].shift;
    }
  }
  r_end = min(r_end, r0 + segs.rows_per_cta);
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  for (int i = threadIdx.x * 4; i < D; i += blockDim.x * 4) {
    const float4 a = *reinterpret_cast<const float4*>(gsc + i);
    const float4 h = *reinterpret_cast<const float4*>(gsh + i);
    *reinterpret_cast<float4*>(s_sc + i) = make_float4(1.f + a.x, 1.f + a.y, 1.f + a.z, 1.f + a.w);
    *reinterpret_cast<float4*>(s_sh + i) = h;
  }
  __syncthreads();
  const float inv_d = 1.f / (float)D;
  // volatile: each pass unpacks again (2 ALU ops per pair) instead of keeping 8 kV fp32 values alive across the passes
  auto lo = [](uint32_t w) {
    uint32_t o;
    asm volatile("shl.b32 %0, %1, 16;" : "=r"(o) : "r"(w));
    return __uint_as_float(o);
  };
  auto hi = [](uint32_t w) {
    uint32_t o;
    asm volatile("and.b32 %0, %1, 0xffff0000;" : "=r"(o) : "r"(w));
    return __uint_as_float(o);
  };
  for (int r = r0 + warp; r < r_end; r += 16) {
    uint4 ra[kV];
    {
      const bf16* xa = x + (long long)b * x_bs + (long long)r * x_ld + lane * 8;
#pragma unroll
      for (int i = 0; i < kV; ++i) ra[i] = *reinterpret_cast<const uint4*>(xa + i * 256);
    }
    float2 sa[4] = {f2(0.f, 0.f), f2(0.f, 0.f), f2(0.f, 0.f), f2(0.f, 0.f)};
#pragma unroll
    for (int i = 0; i < kV; ++i) {
      sa[0] = __fadd2_rn(sa[0], f2(lo(ra[i].x), hi(ra[i].x)));
      sa[1] = __fadd2_rn(sa[1], f2(lo(ra[i].y), hi(ra[i].y)));
      sa[2] = __fadd2_rn(sa[2], f2(lo(ra[i].z), hi(ra[i].z)));
      sa[3] = __fadd2_rn(sa[3], f2(lo(ra[i].w), hi(ra[i].w)));
    }
    const float2 st = __fadd2_rn(__fadd2_rn(sa[0], sa[1]), __fadd2_rn(sa[2], sa[3]));
    const float mean = warp_sum(st.x + st.y) * inv_d;
    const float2 nm = f2(-mean, -mean);
    float2 qa[4] = {f2(0.f, 0.f), f2(0.f, 0.f), f2(0.f, 0.f), f2(0.f, 0.f)};
#pragma unroll
    for (int i = 0; i < kV; ++i) {
      const uint32_t w[4] = {ra[i].x, ra[i].y, ra[i].z, ra[i].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 d = __fadd2_rn(f2(lo(w[j]), hi(w[j])), nm);
        qa[j] = __ffma2_rn(d, d, qa[j]);
      }
    }
    const float2 qt = __fadd2_rn(__fadd2_rn(qa[0], qa[1]), __fadd2_rn(qa[2], qa[3]));
    const float rstd = rsqrtf(warp_sum(qt.x + qt.y) * inv_d + 1e-6f);
    const float2 rs = f2(rstd, rstd);
    bf16* oa = out + (long long)b * o_bs + (long long)r * o_ld + lane * 8;
#pragma unroll
    for (int i = 0; i < kV; ++i) {
      const int c = i * 256 + lane * 8;
      const float4 c0 = *reinterpret_cast<const float4*>(s_sc + c), c1 = *reinterpret_cast<const float4*>(s_sc + c + 4);
      const float4 h0 = *reinterpret_cast<const float4*>(s_sh + c), h1 = *reinterpret_cast<const float4*>(s_sh + c + 4);
      const float2 sc[4] = {f2(c0.x, c0.y), f2(c0.z, c0.w), f2(c1.x, c1.y), f2(c1.z, c1.w)};
      const float2 sh[4] = {f2(h0.x, h0.y), f2(h0.z, h0.w), f2(h1.x, h1.y), f2(h1.z, h1.w)};
      const uint32_t w[4] = {ra[i].x, ra[i].y, ra[i].z, ra[i].w};
      uint32_t pk[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 d = __fadd2_rn(f2(lo(w[j]), hi(w[j])), nm);
        const float2 y = __ffma2_rn(d, __fmul2_rn(rs, sc[j]), sh[j]);
        __nv_bfloat162 hh = __float22bfloat162_rn(y);
        pk[j] = *reinterpret_cast<uint32_t*>(&hh);
      }
      *reinterpret_cast<uint4*>(oa + i * 256) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
      if (i % 2 == 1) asm volatile("" ::: "memory");  // keeps the modulation loads of later pieces from piling up in registers
    }
  }
}

template <int kV>
static void launch_ln_row1(const void* x, long long x_bs, int x_ld, void* out, long long o_bs, int o_ld, int D,
                           const LnSegs& S, int grid, cudaStream_t stream) {
  const int smem = 2 * D * (int)sizeof(float);
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(512);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = get_option("no_pdl") ? 0 : 1;
  RT_CHECK_CUDA(cudaLaunchKernelEx(&cfg, ln_mod_row1_kernel<kV>, (const bf16*)x, x_bs, x_ld, (bf16*)out, o_bs, o_ld, D, S));
  count_launch();
}

void launch_ln_mod(int dtype, const void* x, long long x_bs, int x_ld, void* out, long long o_bs, int o_ld, int batch,
                   int D, int ngroups, const LnModGroup* groups, cudaStream_t stream) {
  RT_REQUIRE(ngroups >= 1 && ngroups <= kLnMaxGroups, "ln_mod: 1..2 row groups");
  LnGroups G;
  G.n = ngroups;
  int rows_total = 0;
  for (int i = 0; i < ngroups; ++i) {
    G.g[i] = groups[i];
    rows_total += groups[i].row_end - groups[i].row_begin;
    RT_REQUIRE(groups[i].ld % 4 == 0, "ln_mod: modulation ld must be a multiple of 4");
  }
  if (rows_total == 0 || batch == 0) return;
  ProfScope ps(PROF_LN, 2.0 * batch * rows_total * (double)D * dtype_size(dtype), stream);
  const long long total = (long long)batch * rows_total;
  // bf16, D = kV * 256: the warp-per-row-pair kernel
  if (dtype == RT_BF16 && D % 256 == 0 && D <= 4096 && (D / 256) % 4 == 0 && batch * ngroups <= kLnMaxSegs &&
      x_ld % 8 == 0 && o_ld % 8 == 0 && (get_option("ln_impl") == 0 || get_option("ln_impl") == 3)) {
    LnSegs S{};
    // one wave of CTAs (the kernels keep whole rows in registers: one CTA per SM)
    int rpc = (int)((total + sm_count() - 1) / sm_count());
    rpc = (rpc + 1) / 2 * 2;
    const bool big = rpc > 64;   // more than one wave
    rpc = rpc < 2 ? 2 : (rpc > 64 ? 64 : rpc);
    // v4 (one row per warp, 16 warps) up to one wave of 64-row CTAs and D <= 3072 (128 registers hold the row), v3 (a
    // pair of rows per warp) beyond; ln_impl = 3 forces v3.  Both compute a row with the same operations in the same
    // order, so the choice never shows in the results (a sharded run must equal the unsharded one bit for bit).
    const bool row1 = get_option("ln_impl") == 0 && D <= 3072 && !big;
    S.rows_per_cta = rpc;
    int cta = 0;
    for (int b = 0; b < batch; ++b)
      for (int g = 0; g < ngroups; ++g) {
        const int rows = groups[g].row_end - groups[g].row_begin;
        if (rows <= 0) continue;
        LnSeg& e = S.s[S.n++];
        e.batch = b; e.row_begin = groups[g].row_begin; e.row_end = groups[g].row_end; e.cta_begin = cta;
        e.scale = groups[g].scale + (long long)b * groups[g].ld;
        e.shift = groups[g].shift + (long long)b * groups[g].ld;
        cta += (rows + rpc - 1) / rpc;
      }
    if (row1) {
      switch (D / 256) {
        case 4: launch_ln_row1<4>(x, x_bs, x_ld, out, o_bs, o_ld, D, S, cta, stream); return;
        case 8: launch_ln_row1<8>(x, x_bs, x_ld, out, o_bs, o_ld, D, S, cta, stream); return;
        case 12: launch_ln_row1<12>(x, x_bs, x_ld, out, o_bs, o_ld, D, S, cta, stream); return;
        default: break;
      }
    }
    switch (D / 256) {
      case 4: launch_ln_rows2<4>(x, x_bs, x_ld, out, o_bs, o_ld, D, S, cta, stream); return;
      case 8: launch_ln_rows2<8>(x, x_bs, x_ld, out, o_bs, o_ld, D, S, cta, stream); return;
      case 12: launch_ln_rows2<12>(x, x_bs, x_ld, out, o_bs, o_ld, D, S, cta, stream); return;
      case 16: launch_ln_rows2<16>(x, x_bs, x_ld, out, o_bs, o_ld, D, S, cta, stream); return;
      default: break;
    }
  }
  const int threads = 256, wpb = threads / 32;
  long long blocks = (total + wpb - 1) / wpb;
  long long cap = (long long)sm_count() * 8;
  if (blocks > cap) blocks = cap;
  RT_DISPATCH_DTYPE(dtype, T, {
    constexpr int N = VecT<T>::N;
    RT_REQUIRE(D % N == 0 && x_ld % N == 0 && o_ld % N == 0, "ln_mod: D and lds must be multiples of the vector width");
    const int nvec = D / N;
    if (nvec <= 1024 && nvec >= 64 && !get_option("ln_warp_rows")) {
      constexpr int kR = 4;
      const int cta_threads = (nvec + 31) / 32 * 32;
      const long long cap_cta = (long long)sm_count() * (cta_threads <= 384 ? 2 : 1);
      long long rpc = (total + cap_cta - 1) / cap_cta;
      rpc = (rpc + kR - 1) / kR * kR;
      const long long grid = (total + rpc - 1) / rpc;
      RT_REQUIRE(total < (1ll << 31), "ln_mod: too many rows");
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = dim3((unsigned)grid);
      cfg.blockDim = dim3(cta_threads);
      cfg.stream = stream;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[0].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = attr;
      cfg.numAttrs = get_option("no_pdl") ? 0 : 1;
      const T* xp = (const T*)x;
      T* op = (T*)out;
      const int rpc_i = (int)rpc;
      if (cta_threads <= 384)
        RT_CHECK_CUDA(cudaLaunchKernelEx(&cfg, ln_mod_cta_kernel<T, 384, 2>, xp, x_bs, x_ld, op, o_bs, o_ld, batch,
                                         rows_total, D, G, rpc_i));
      else
        RT_CHECK_CUDA(cudaLaunchKernelEx(&cfg, ln_mod_cta_kernel<T, 1024, 1>, xp, x_bs, x_ld, op, o_bs, o_ld, batch,
                                         rows_total, D, G, rpc_i));
      RT_POST_LAUNCH();
      return;
    }
    const int per_lane = (nvec + 31) / 32;
    RT_REQUIRE(per_lane <= 24, "ln_mod: D too large");
    if (per_lane <= 2)
      ln_mod_kernel<T, 2><<<(int)blocks, threads, 0, stream>>>((const T*)x, x_bs, x_ld, (T*)out, o_bs, o_ld, batch,
                                                               rows_total, D, G);
    else if (per_lane <= 4)
      ln_mod_kernel<T, 4><<<(int)blocks, threads, 0, stream>>>((const T*)x, x_bs, x_ld, (T*)out, o_bs, o_ld, batch,
                                                               rows_total, D, G);
    else if (per_lane <= 12)
      ln_mod_kernel<T, 12><<<(int)blocks, threads, 0, stream>>>((const T*)x, x_bs, x_ld, (T*)out, o_bs, o_ld, batch,
                                                                rows_total, D, G);
    else if (per_lane <= 16)
      ln_mod_kernel<T, 16><<<(int)blocks, threads, 0, stream>>>((const T*)x, x_bs, x_ld, (T*)out, o_bs, o_ld, batch,
                                                                rows_total, D, G);
    else
      ln_mod_kernel<T, 24><<<(int)blocks, threads, 0, stream>>>((const T*)x, x_bs, x_ld, (T*)out, o_bs, o_ld, batch,
                                                                rows_total, D, G);
  });
  RT_POST_LAUNCH();
}

// ------------------------------------------------------------------------------------------------
// Grouped GEMV (M = batch <= 8 per pass): one warp per output row, weights streamed once with 16-byte
// loads; x (fp32, tiny) comes through L1.  Used for the AdaLN modulation vectors of ALL blocks in one
// launch and for the time/guidance/pooled MLPs (diffusers CombinedTimestepGuidanceTextProjEmbeddings,
// RepText/controlnet_flux.py:287-291).
// ------------------------------------------------------------------------------------------------
template <typename T, int NB>
__global__ void __launch_bounds__(256) gemv_grouped_kernel(const float* __restrict__ x, int x_ld, int b0, int K,
                                                           const GemvJob* __restrict__ jobs,
                                                           const int* __restrict__ prefix, int njobs, int total_rows,
                                                           float* __restrict__ out, int out_ld, int silu_out,
                                                           int accumulate, int row_base) {
  constexpr int N = VecT<T>::N;
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (int row_l = blockIdx.x * wpb + (threadIdx.x >> 5); row_l < total_rows; row_l += gridDim.x * wpb) {
    const int row = row_l + row_base;  // prefix[] holds absolute row offsets of the model's whole job table
    int lo = 0, hi = njobs - 1;  // last job with prefix[j] <= row
    while (lo < hi) {
      int mid = (lo + hi + 1) >> 1;
      if (prefix[mid] <= row) lo = mid; else hi = mid - 1;
    }
    const GemvJob J = jobs[lo];
    const int r = row - prefix[lo];
    const T* w = reinterpret_cast<const T*>(J.W) + (long long)r * K;
    float acc[NB];
#pragma unroll
    for (int b = 0; b < NB; ++b) acc[b] = 0.f;
    for (int c = lane * N; c < K; c += 32 * N) {
      float wv[N];
      ldvec(w + c, wv);
#pragma unroll
      for (int b = 0; b < NB; ++b) {
        const float* xb = x + (long long)(b0 + b) * x_ld + c;
#pragma unroll
        for (int j = 0; j < N; j += 4) {
          float4 xv = *reinterpret_cast<const float4*>(xb + j);
          acc[b] += wv[j] * xv.x + wv[j + 1] * xv.y + wv[j + 2] * xv.z + wv[j + 3] * xv.w;
        }
      }
    }
#pragma unroll
    for (int b = 0; b < NB; ++b) acc[b] = warp_sum(acc[b]);
    if (lane == 0) {
      const float bias = J.bias ? to_f(reinterpret_cast<const T*>(J.bias)[r]) : 0.f;
#pragma unroll
      for (int b = 0; b < NB; ++b) {
        float v = acc[b] + bias;
        if (silu_out) v = v / (1.f + expf(-v));
        float* o = out + (long long)(b0 + b) * out_ld + J.out_off + r;
        *o = accumulate ? (*o + v) : v;
      }
    }
  }
}


// kR consecutive output rows per warp: the activation chunk (fp32, L1-resident) is loaded once for kR weight
// rows instead of once per row, and kR independent 16-byte weight loads are in flight per lane per iteration.
// Rows never straddle a job (the host checks rows % kR == 0 for every job).
template <typename T, int NB, int kR>
__global__ void __launch_bounds__(256) gemv_grouped_rows_kernel(const float* __restrict__ x, int x_ld, int b0, int K,
                                                                const GemvJob* __restrict__ jobs,
                                                                const int* __restrict__ prefix, int njobs,
                                                                int total_rows, float* __restrict__ out, int out_ld,
                                                                int silu_out, int accumulate, int row_base,
                                                                const GemvPeers peers) {
  constexpr int N = VecT<T>::N;
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (int row_l = (blockIdx.x * wpb + (threadIdx.x >> 5)) * kR; row_l < total_rows; row_l += gridDim.x * wpb * kR) {
    const int row = row_l + row_base;
    int lo = 0, hi = njobs - 1;  // last job with prefix[j] <= row
    while (lo < hi) {
      int mid = (lo + hi + 1) >> 1;
      if (prefix[mid] <= row) lo = mid; else hi = mid - 1;
    }
    const GemvJob J = jobs[lo];
    const int r = row - prefix[lo];
    const T* w = reinterpret_cast<const T*>(J.W) + (long long)r * K;
    float acc[kR][NB];
#pragma unroll
    for (int i = 0; i < kR; ++i)
#pragma unroll
      for (int b = 0; b < NB; ++b) acc[i][b] = 0.f;
#pragma unroll 2
    for (int c = lane * N; c < K; c += 32 * N) {
      float wv[kR][N];
#pragma unroll
      for (int i = 0; i < kR; ++i) ldvec(w + (long long)i * K + c, wv[i]);
#pragma unroll
      for (int b = 0; b < NB; ++b) {
        const float* xb = x + (long long)(b0 + b) * x_ld + c;
#pragma unroll
        for (int j = 0; j < N; j += 4) {
          const float4 xv = *reinterpret_cast<const float4*>(xb + j);
#pragma unroll
          for (int i = 0; i < kR; ++i)
            acc[i][b] += wv[i][j] * xv.x + wv[i][j + 1] * xv.y + wv[i][j + 2] * xv.z + wv[i][j + 3] * xv.w;
        }
      }
    }
#pragma unroll
    for (int i = 0; i < kR; ++i)
#pragma unroll
      for (int b = 0; b < NB; ++b) acc[i][b] = warp_sum(acc[i][b]);
    if (lane == 0) {
#pragma unroll
      for (int i = 0; i < kR; ++i) {
        const float bias = J.bias ? to_f(reinterpret_cast<const T*>(J.bias)[r + i]) : 0.f;
#pragma unroll
        for (int b = 0; b < NB; ++b) {
          float v = acc[i][b] + bias;
          if (silu_out) v = v / (1.f + expf(-v));
          const long long off = (long long)(b0 + b) * out_ld + J.out_off + r + i;
          if (peers.n > 0) {
            // sequence-parallel row shard: the result goes to EVERY rank's copy (peer-mapped pointers)
            for (int d = 0; d < peers.n; ++d) peers.p[d][off] = v;
          } else {
            float* o = out + off;
            *o = accumulate ? (*o + v) : v;
          }
        }
      }
    }
  }
}

// The same rows for MANY activation vectors (the AdaLN vectors of all the steps of an image in one pass,
// rt_model_build_modulation_table): kR consecutive rows per warp as above, and the warp walks the batch in tiles of NB
// vectors - the first tile streams the rows' weights from HBM, the others find them in L1 / L2 (kR rows of K bf16 =
// 24 KB per warp), so the 6.5 GB of AdaLN weights are read from DRAM once for the whole table instead of once per step.
// Every (row, vector) accumulator sees the same operations in the same order as in gemv_grouped_rows_kernel: the table
// holds the bits a per-step launch would produce (tests/test_model_gpu.py).
template <typename T, int NB, int kR>
__global__ void __launch_bounds__(256) gemv_grouped_rows_table_kernel(const float* __restrict__ x, int x_ld, int batch, int K,
                                                                      const GemvJob* __restrict__ jobs,
                                                                      const int* __restrict__ prefix, int njobs,
                                                                      int total_rows, float* __restrict__ out, int out_ld) {
  constexpr int N = VecT<T>::N;
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (int row = (blockIdx.x * wpb + (threadIdx.x >> 5)) * kR; row < total_rows; row += gridDim.x * wpb * kR) {
    int lo = 0, hi = njobs - 1;  // last job with prefix[j] <= row
    while (lo < hi) {
      int mid = (lo + hi + 1) >> 1;
      if (prefix[mid] <= row) lo = mid; else hi = mid - 1;
    }
    const GemvJob J = jobs[lo];
    const int r = row - prefix[lo];
    const T* w = reinterpret_cast<const T*>(J.W) + (long long)r * K;
    for (int b0 = 0; b0 < batch; b0 += NB) {
      float acc[kR][NB];
#pragma unroll
      for (int i = 0; i < kR; ++i)
#pragma unroll
        for (int b = 0; b < NB; ++b) acc[i][b] = 0.f;
      for (int c = lane * N; c < K; c += 32 * N) {
        float wv[kR][N];
#pragma unroll
        for (int i = 0; i < kR; ++i) ldvec(w + (long long)i * K + c, wv[i]);
#pragma unroll
        for (int b = 0; b < NB; ++b) {
          const int bb = b0 + b < batch ? b0 + b : batch - 1;  // (tail tile: a valid vector again, result dropped)
          const float* xb = x + (long long)bb * x_ld + c;
#pragma unroll
          for (int j = 0; j < N; j += 4) {
            const float4 xv = *reinterpret_cast<const float4*>(xb + j);
#pragma unroll
            for (int i = 0; i < kR; ++i)
              acc[i][b] += wv[i][j] * xv.x + wv[i][j + 1] * xv.y + wv[i][j + 2] * xv.z + wv[i][j + 3] * xv.w;
          }
        }
      }
#pragma unroll
      for (int i = 0; i < kR; ++i)
#pragma unroll
        for (int b = 0; b < NB; ++b) acc[i][b] = warp_sum(acc[i][b]);
      if (lane == 0) {
#pragma unroll
        for (int i = 0; i < kR; ++i) {
          const float bias = J.bias ? to_f(reinterpret_cast<const T*>(J.bias)[r + i]) : 0.f;
#pragma unroll
          for (int b = 0; b < NB; ++b)
            if (b0 + b < batch) out[(long long)(b0 + b) * out_ld + J.out_off + r + i] = acc[i][b] + bias;
        }
      }
    }
  }
}

#define RT_TABLE_CASE(NB_, KR_) \
    gemv_grouped_rows_table_kernel<T, NB_, KR_><<<blocks, threads, 0, stream>>>(x, x_ld, batch, K, jobs_dev, prefix_dev, njobs, \
                                                                              total_rows, out, out_ld)
void launch_gemv_grouped_table(int wdtype, const float* x, int x_ld, int batch, int K, const GemvJob* jobs_dev,
                               const int* prefix_dev, int njobs, int total_rows, float* out, int out_ld,
                               cudaStream_t stream) {
  if (batch == 0 || total_rows == 0) return;
  ProfScope ps(PROF_GEMV, (double)total_rows * K * dtype_size(wdtype), stream);
  RT_REQUIRE(x_ld % 4 == 0 && total_rows % 4 == 0, "gemv table: x_ld and the row count must be multiples of 4");
  const int threads = 256, wpb = threads / 32;
  // 8 rows x 4 vectors per warp pass: measured best of (rows, vectors) = (4, 7) 16.8 ms, (4, 14) 10.0, (8, 7) 9.8,
  // (8, 4) 7.8 for the FLUX.1-dev transformer's 28-step table (profiles/r2_modulation_table.txt); (4, 7) when a model's
  // row count is not a multiple of 8
  const int kRr = total_rows % 8 == 0 ? 8 : 4;
  int blocks = (total_rows / kRr + wpb - 1) / wpb;
  const int cap = sm_count() * 4;
  if (blocks > cap) blocks = cap;
  RT_DISPATCH_DTYPE(wdtype, T, {
    RT_REQUIRE(K % VecT<T>::N == 0, "gemv: K must be a multiple of the vector width");
    if (kRr == 8) RT_TABLE_CASE(4, 8);
    else RT_TABLE_CASE(7, 4);
    RT_POST_LAUNCH();
  });
}

void launch_gemv_grouped(int wdtype, const float* x, int x_ld, int batch, int K, const GemvJob* jobs_dev,
                         const int* prefix_dev, int njobs, int total_rows, float* out, int out_ld, int silu_out,
                         int accumulate, cudaStream_t stream, bool rows_multiple_of_4, int row_base,
                         const GemvPeers* peers) {
  if (batch == 0 || total_rows == 0) return;
  ProfScope ps(PROF_GEMV, (double)total_rows * K * dtype_size(wdtype), stream);
  RT_REQUIRE(x_ld % 4 == 0, "gemv: x_ld must be a multiple of 4");
  const int threads = 256, wpb = threads / 32;
  int blocks = (total_rows + wpb - 1) / wpb;
  int cap = sm_count() * 8;
  if (blocks > cap) blocks = cap;
  RT_DISPATCH_DTYPE(wdtype, T, {
    RT_REQUIRE(K % VecT<T>::N == 0, "gemv: K must be a multiple of the vector width");
    int b0 = 0;
    constexpr int kR = 4;
    const bool multi = rows_multiple_of_4 && total_rows % kR == 0 && (!get_option("gemv_single_row") || peers);
    RT_REQUIRE(!peers || (multi && !accumulate), "gemv: the peer-output form needs 4-row jobs and no accumulation");
    const GemvPeers pv = peers ? *peers : GemvPeers{};
    int mblocks = (total_rows / kR + wpb - 1) / wpb;
    if (mblocks > cap) mblocks = cap;
    while (b0 < batch) {
      int nb = batch - b0;
      if (multi && nb >= 2) {
        gemv_grouped_rows_kernel<T, 2, kR><<<mblocks, threads, 0, stream>>>(x, x_ld, b0, K, jobs_dev, prefix_dev, njobs,
                                                                            total_rows, out, out_ld, silu_out, accumulate, row_base, pv);
        b0 += 2;
      } else if (multi) {
        gemv_grouped_rows_kernel<T, 1, kR><<<mblocks, threads, 0, stream>>>(x, x_ld, b0, K, jobs_dev, prefix_dev, njobs,
                                                                            total_rows, out, out_ld, silu_out, accumulate, row_base, pv);
        b0 += 1;
      } else if (nb >= 4) {
        gemv_grouped_kernel<T, 4><<<blocks, threads, 0, stream>>>(x, x_ld, b0, K, jobs_dev, prefix_dev, njobs,
                                                                  total_rows, out, out_ld, silu_out, accumulate, row_base);
        b0 += 4;
      } else if (nb >= 2) {
        gemv_grouped_kernel<T, 2><<<blocks, threads, 0, stream>>>(x, x_ld, b0, K, jobs_dev, prefix_dev, njobs,
                                                                  total_rows, out, out_ld, silu_out, accumulate, row_base);
        b0 += 2;
      } else {
        gemv_grouped_kernel<T, 1><<<blocks, threads, 0, stream>>>(x, x_ld, b0, K, jobs_dev, prefix_dev, njobs,
                                                                  total_rows, out, out_ld, silu_out, accumulate, row_base);
        b0 += 1;
      }
      RT_POST_LAUNCH();
    }
  });
}

// ------------------------------------------------------------------------------------------------
// Timestep sinusoid: diffusers get_timestep_embedding(256, flip_sin_to_cos=True, shift=0) applied to
// T(t) * 1000 computed IN THE MODEL DTYPE (RepText/controlnet_flux.py:282-284: `timestep.to(dtype) * 1000`).
// ------------------------------------------------------------------------------------------------
template <typename T>
__global__ void time_sinusoid_kernel(const T* __restrict__ t, int t_batch, int batch, float* __restrict__ out) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= batch * 128) return;
  int b = idx / 128, j = idx % 128;
  T tv = t[t_batch == 1 ? 0 : b];
  float v = to_f(from_f<T>(to_f(tv) * 1000.f));
  float f = expf(-logf(10000.f) * (float)j / 128.f);
  float s, c;
  sincosf(v * f, &s, &c);
  out[b * 256 + j] = c;
  out[b * 256 + 128 + j] = s;
}
void launch_time_sinusoid(int dtype, const void* t, int t_batch, int batch, float* out, cudaStream_t stream) {
  if (batch == 0) return;
  int n = batch * 128;
  RT_DISPATCH_DTYPE(dtype, T,
                    (time_sinusoid_kernel<T><<<(n + 127) / 128, 128, 0, stream>>>((const T*)t, t_batch, batch, out)));
  RT_POST_LAUNCH();
}

__global__ void silu_f32_kernel(const float* __restrict__ x, float* __restrict__ out, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float v = x[i];
    out[i] = v / (1.f + expf(-v));
  }
}
void launch_silu_f32(const float* x, float* out, long long n, cudaStream_t stream) {
  if (n == 0) return;
  int blocks = (int)((n + 255) / 256);
  if (blocks > sm_count() * 8) blocks = sm_count() * 8;
  silu_f32_kernel<<<blocks, 256, 0, stream>>>(x, out, n);
  RT_POST_LAUNCH();
}

template <typename T>
__global__ void cast_to_f32_kernel(const T* __restrict__ x, float* __restrict__ out, long long n) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = to_f(x[i]);
}
void launch_cast_to_f32(int dtype, const void* x, float* out, long long n, cudaStream_t stream) {
  if (n == 0) return;
  int blocks = (int)((n + 255) / 256);
  if (blocks > sm_count() * 8) blocks = sm_count() * 8;
  RT_DISPATCH_DTYPE(dtype, T, (cast_to_f32_kernel<T><<<blocks, 256, 0, stream>>>((const T*)x, out, n)));
  RT_POST_LAUNCH();
}

// ------------------------------------------------------------------------------------------------
// RoPE table (FluxPosEmbed, RepText/controlnet_flux.py:65, :316-317): float64 angle, cos/sin to fp32.
// Stored compactly as (cos, sin) per PAIR: [S, hd/2] float2 (the reference repeat_interleaves by 2).
// ------------------------------------------------------------------------------------------------
__global__ void rope_table_kernel(const float* __restrict__ ids, int S, int d0, int d1, int d2,
                                  float2* __restrict__ out) {
  const int half = (d0 + d1 + d2) / 2;
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= S * half) return;
  int s = idx / half, j = idx % half;
  int axis, jj, d;
  if (j < d0 / 2) { axis = 0; jj = j; d = d0; }
  else if (j < (d0 + d1) / 2) { axis = 1; jj = j - d0 / 2; d = d1; }
  else { axis = 2; jj = j - (d0 + d1) / 2; d = d2; }
  double omega = 1.0 / pow(10000.0, (double)(2 * jj) / (double)d);
  double ang = (double)ids[s * 3 + axis] * omega;
  out[idx] = make_float2((float)cos(ang), (float)sin(ang));
}
void launch_rope_table(const float* ids, int S, const int* axes, float2* out, cudaStream_t stream) {
  int half = (axes[0] + axes[1] + axes[2]) / 2;
  int n = S * half;
  if (n == 0) return;
  rope_table_kernel<<<(n + 255) / 256, 256, 0, stream>>>(ids, S, axes[0], axes[1], axes[2], out);
  RT_POST_LAUNCH();
}

// ------------------------------------------------------------------------------------------------
// Unfused QK-RMSNorm (eps 1e-6, over head_dim) * w + interleaved-pair RoPE, in place.  One warp per
// (batch, row, head).  The tcgen05 GEMM fuses this into its epilogue; this form serves the SIMT path.
// ------------------------------------------------------------------------------------------------
template <typename T, int EPL>  // elements per lane (hd = 32 * EPL), EPL even
__global__ void __launch_bounds__(256) qknorm_rope_kernel(T* __restrict__ buf, long long bs, int ld, int col0,
                                                          int batch, int row0, int rows, int heads,
                                                          const T* __restrict__ w, const float2* __restrict__ rope,
                                                          int rope_row0) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  const long long total = (long long)batch * rows * heads;
  constexpr int hd = 32 * EPL;
  for (long long it = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); it < total;
       it += (long long)gridDim.x * wpb) {
    int h = (int)(it % heads);
    long long br = it / heads;
    int r = (int)(br % rows);
    int b = (int)(br / rows);
    T* p = buf + (long long)b * bs + (long long)(row0 + r) * ld + col0 + h * hd + lane * EPL;
    float v[EPL];
    float ss = 0.f;
#pragma unroll
    for (int j = 0; j < EPL; ++j) {
      v[j] = to_f(p[j]);
      ss += v[j] * v[j];
    }
    ss = warp_sum(ss);
    const float rs = rsqrtf(ss / (float)hd + 1e-6f);
#pragma unroll
    for (int j = 0; j < EPL; ++j) v[j] = v[j] * rs * to_f(w[lane * EPL + j]);
    if (rope) {
      const float2* rp = rope + (long long)(rope_row0 + r) * (hd / 2) + lane * (EPL / 2);
#pragma unroll
      for (int j = 0; j < EPL; j += 2) {
        float2 cs = rp[j / 2];
        float a = v[j], c = v[j + 1];
        v[j] = a * cs.x - c * cs.y;
        v[j + 1] = c * cs.x + a * cs.y;
      }
    }
#pragma unroll
    for (int j = 0; j < EPL; ++j) p[j] = from_f<T>(v[j]);
  }
}
void launch_qknorm_rope(int dtype, void* buf, long long bs, int ld, int col0, int batch, int row0, int rows,
                        int heads, int hd, const void* norm_w, const float2* rope, int rope_row0,
                        cudaStream_t stream) {
  long long total = (long long)batch * rows * heads;
  if (total == 0) return;
  RT_REQUIRE(hd == 64 || hd == 128, "qknorm_rope: head_dim must be 64 or 128");
  int blocks = (int)((total + 7) / 8);
  if (blocks > sm_count() * 8) blocks = sm_count() * 8;
  RT_DISPATCH_DTYPE(dtype, T, {
    if (hd == 64)
      qknorm_rope_kernel<T, 2><<<blocks, 256, 0, stream>>>((T*)buf, bs, ld, col0, batch, row0, rows, heads,
                                                           (const T*)norm_w, rope, rope_row0);
    else
      qknorm_rope_kernel<T, 4><<<blocks, 256, 0, stream>>>((T*)buf, bs, ld, col0, batch, row0, rows, heads,
                                                           (const T*)norm_w, rope, rope_row0);
  });
  RT_POST_LAUNCH();
}

// ------------------------------------------------------------------------------------------------
// FlowMatch Euler step (scheduler.step at RepText/pipeline_flux_controlnet.py:1109) and true-CFG
// (pipeline_flux_controlnet_inpaint.py:1264-1270).  16 bytes per thread per tensor, grid-stride.
//   out = T( float(x) + float(T(float(T(dt)) * float(v))) )
// diffusers 0.36 keeps scheduler.sigmas ON THE DEVICE (set_timesteps(..., device=)), so `dt = sigma_next - sigma` is a
// 0-dim fp32 CUDA tensor; torch's type promotion ignores 0-dim tensors of the same category, the multiply runs in the
// dtype of model_output and its fetch casts dt to that dtype first: dt is ROUNDED TO bf16 in a bf16 run.  Measured on
// a B200 with torch 2.11 (tools/euler_dt_probe.py, profiles/r2_euler_dt_probe.txt): that form matches bit for bit, the
// host-scalar form (sigmas on the CPU, as EulerDiscreteScheduler keeps them: dt stays fp32) differs in 12608 of
// 262144 elements.  Option euler_dt_host=1 selects the host-scalar form.
// ------------------------------------------------------------------------------------------------
static float euler_dt(int dtype, float sigma, float sigma_next) {
  const float dt = sigma_next - sigma;
  if (dtype == RT_BF16 && !get_option("euler_dt_host")) return __bfloat162float(__float2bfloat16_rn(dt));
  return dt;
}
template <typename T>
__global__ void __launch_bounds__(256) euler_kernel(const T* __restrict__ v, const T* __restrict__ x,
                                                    T* __restrict__ out, long long n, float dt) {
  constexpr int N = VecT<T>::N;
  const long long nv = n / N;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += (long long)gridDim.x * blockDim.x) {
    float a[N], b[N], o[N];
    ldvec(v + i * N, a);
    ldvec(x + i * N, b);
#pragma unroll
    for (int j = 0; j < N; ++j) o[j] = __fadd_rn(b[j], to_f(from_f<T>(__fmul_rn(dt, a[j]))));
    stvec(out + i * N, o);
  }
  if (blockIdx.x == 0 && threadIdx.x < (int)(n - nv * N)) {
    long long i = nv * N + threadIdx.x;
    out[i] = from_f<T>(__fadd_rn(to_f(x[i]), to_f(from_f<T>(__fmul_rn(dt, to_f(v[i]))))));
  }
}
static int ew_blocks(long long nvec) {
  long long b = (nvec + 255) / 256;
  long long cap = (long long)sm_count() * 8;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return (int)b;
}
void launch_euler_step(int dtype, const void* v, const void* x, void* out, long long n, float sigma,
                       float sigma_next, cudaStream_t stream) {
  if (n == 0) return;
  const float dt = euler_dt(dtype, sigma, sigma_next);
  RT_DISPATCH_DTYPE(dtype, T, (euler_kernel<T><<<ew_blocks(n / VecT<T>::N), 256, 0, stream>>>(
                                  (const T*)v, (const T*)x, (T*)out, n, dt)));
  RT_POST_LAUNCH();
}

// noise_pred = uncond + s * (text - uncond)   (each op rounded to T like the reference's tensor ops)
template <typename T>
__device__ __forceinline__ float cfg_one(float u, float t, float s, int zero_pred) {
  if (zero_pred) return to_f(from_f<T>(__fmul_rn(t, 0.f)));
  float d = to_f(from_f<T>(__fsub_rn(t, u)));
  float m = to_f(from_f<T>(__fmul_rn(s, d)));
  return to_f(from_f<T>(__fadd_rn(u, m)));
}
template <typename T, bool kEuler>
__global__ void __launch_bounds__(256) cfg_kernel(const T* __restrict__ v2, const T* __restrict__ x,
                                                  T* __restrict__ out, long long n, float s, int zero_pred,
                                                  float dt) {
  constexpr int N = VecT<T>::N;
  const long long nv = n / N;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += (long long)gridDim.x * blockDim.x) {
    float u[N], t[N], o[N];
    ldvec(v2 + i * N, u);
    ldvec(v2 + n + i * N, t);
    if constexpr (kEuler) {
      float xv[N];
      ldvec(x + i * N, xv);
#pragma unroll
      for (int j = 0; j < N; ++j)
        o[j] = __fadd_rn(xv[j], to_f(from_f<T>(__fmul_rn(dt, cfg_one<T>(u[j], t[j], s, zero_pred)))));
    } else {
#pragma unroll
      for (int j = 0; j < N; ++j) o[j] = cfg_one<T>(u[j], t[j], s, zero_pred);
    }
    stvec(out + i * N, o);
  }
  if (blockIdx.x == 0 && threadIdx.x < (int)(n - nv * N)) {
    long long i = nv * N + threadIdx.x;
    float p = cfg_one<T>(to_f(v2[i]), to_f(v2[n + i]), s, zero_pred);
    out[i] = kEuler ? from_f<T>(__fadd_rn(to_f(x[i]), to_f(from_f<T>(__fmul_rn(dt, p))))) : from_f<T>(p);
  }
}
void launch_cfg_combine(int dtype, const void* v2, void* out, long long n, float s, int zero_pred,
                        cudaStream_t stream) {
  if (n == 0) return;
  RT_REQUIRE((n * (long long)dtype_size(dtype)) % 16 == 0, "cfg: n*sizeof(T) must be a multiple of 16");
  RT_DISPATCH_DTYPE(dtype, T, (cfg_kernel<T, false><<<ew_blocks(n / VecT<T>::N), 256, 0, stream>>>(
                                  (const T*)v2, nullptr, (T*)out, n, s, zero_pred, 0.f)));
  RT_POST_LAUNCH();
}
void launch_cfg_euler(int dtype, const void* v2, const void* x, void* out, long long n, float s, int zero_pred,
                      float sigma, float sigma_next, cudaStream_t stream) {
  if (n == 0) return;
  RT_REQUIRE((n * (long long)dtype_size(dtype)) % 16 == 0, "cfg: n*sizeof(T) must be a multiple of 16");
  RT_DISPATCH_DTYPE(dtype, T, (cfg_kernel<T, true><<<ew_blocks(n / VecT<T>::N), 256, 0, stream>>>(
                                  (const T*)v2, (const T*)x, (T*)out, n, s, zero_pred, euler_dt(dtype, sigma, sigma_next))));
  RT_POST_LAUNCH();
}

// out[b, r, :] = T(mask[r] * T(scale * x[b, r, :])) (+ acc_in[b, r, :])
template <typename T>
__global__ void __launch_bounds__(256) mask_scale_add_kernel(const T* __restrict__ x, const T* __restrict__ mask,
                                                             const T* __restrict__ acc_in, T* __restrict__ out,
                                                             int rows, int D, long long total_vec, float scale) {
  constexpr int N = VecT<T>::N;
  const int vpr = D / N;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total_vec;
       i += (long long)gridDim.x * blockDim.x) {
    long long row = i / vpr;
    int r = (int)(row % rows);
    float m = mask ? to_f(mask[r]) : 1.f;
    float a[N], o[N];
    ldvec(x + i * N, a);
#pragma unroll
    for (int j = 0; j < N; ++j) o[j] = to_f(from_f<T>(__fmul_rn(m, to_f(from_f<T>(__fmul_rn(scale, a[j]))))));
    if (acc_in) {
      float c[N];
      ldvec(acc_in + i * N, c);
#pragma unroll
      for (int j = 0; j < N; ++j) o[j] = __fadd_rn(o[j], c[j]);
    }
    stvec(out + i * N, o);
  }
}
void launch_mask_scale_add(int dtype, const void* x, const void* mask, const void* acc_in, void* out, int batch,
                           int rows, int D, float scale, cudaStream_t stream) {
  long long n = (long long)batch * rows * D;
  if (n == 0) return;
  RT_DISPATCH_DTYPE(dtype, T, {
    RT_REQUIRE(D % VecT<T>::N == 0, "mask_scale_add: D must be a multiple of the vector width");
    long long nv = n / VecT<T>::N;
    mask_scale_add_kernel<T><<<ew_blocks(nv), 256, 0, stream>>>((const T*)x, (const T*)mask, (const T*)acc_in,
                                                                (T*)out, rows, D, nv, scale);
  });
  RT_POST_LAUNCH();
}

// glyph-latent init: out = mask ? T(T(wg * z) + T(wn * noise)) : noise
template <typename T>
__global__ void __launch_bounds__(256) glyph_blend_kernel(const T* __restrict__ noise, const T* __restrict__ z,
                                                          const unsigned char* __restrict__ mask, T* __restrict__ out,
                                                          long long n, float wg, float wn) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    float nz = to_f(noise[i]);
    float r = nz;
    if (mask[i]) r = __fadd_rn(to_f(from_f<T>(__fmul_rn(wg, to_f(z[i])))), to_f(from_f<T>(__fmul_rn(wn, nz))));
    out[i] = from_f<T>(r);
  }
}
void launch_glyph_blend(int dtype, const void* noise, const void* z, const unsigned char* mask, void* out,
                        long long n, float wg, float wn, cudaStream_t stream) {
  if (n == 0) return;
  RT_DISPATCH_DTYPE(dtype, T, (glyph_blend_kernel<T><<<ew_blocks(n), 256, 0, stream>>>(
                                  (const T*)noise, (const T*)z, mask, (T*)out, n, wg, wn)));
  RT_POST_LAUNCH();
}

template <typename T>
__global__ void __launch_bounds__(256) copy_rows_kernel(const T* __restrict__ src, long long s_bs, int s_ld,
                                                        int s_row0, T* __restrict__ dst, long long d_bs, int d_ld,
                                                        int d_row0, int rows, int vpr, long long total_vec) {
  constexpr int N = VecT<T>::N;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total_vec;
       i += (long long)gridDim.x * blockDim.x) {
    int c = (int)(i % vpr);
    long long row = i / vpr;
    int r = (int)(row % rows);
    int b = (int)(row / rows);
    float a[N];
    ldvec(src + (long long)b * s_bs + (long long)(s_row0 + r) * s_ld + c * N, a);
    stvec(dst + (long long)b * d_bs + (long long)(d_row0 + r) * d_ld + c * N, a);
  }
}
void launch_copy_rows(int dtype, const void* src, long long s_bs, int s_ld, int s_row0, void* dst, long long d_bs,
                      int d_ld, int d_row0, int batch, int rows, int D, cudaStream_t stream) {
  long long n = (long long)batch * rows * D;
  if (n == 0) return;
  RT_DISPATCH_DTYPE(dtype, T, {
    constexpr int N = VecT<T>::N;
    RT_REQUIRE(D % N == 0 && s_ld % N == 0 && d_ld % N == 0, "copy_rows: vector alignment");
    long long nv = n / N;
    copy_rows_kernel<T><<<ew_blocks(nv), 256, 0, stream>>>((const T*)src, s_bs, s_ld, s_row0, (T*)dst, d_bs, d_ld,
                                                           d_row0, rows, D / N, nv);
  });
  RT_POST_LAUNCH();
}

}  // namespace rt
