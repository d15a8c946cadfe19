// A/B kernels of the joint attention (compiled only with -DRT_AB_VARIANTS: python -m reptext_b200.build --ab ->
// csrc/librt_reptext_ab.so, loaded with RT_LIB=...).  None of these is the product: each one is a measured experiment
// that was correct and NOT faster than attn_tc_kernel<0, 0, true, 0x88, true> (profiles/r1_attn_variant_sweep.txt,
// profiles/r2_attn_*).  Included by attn_sm100.cu inside namespace rt::{anonymous}.
#pragma once


// =================================================================================================
// v4: 64-key blocks with a DOUBLE-BUFFERED score tile.
//
// In the kernel above P aliases S, so Q K^T of block j+1 cannot be issued before P V of block j: per query tile the
// chain  Q K^T -> softmax -> P V -> Q K^T  is strictly serial and its latency (two mbarrier hand-offs, TMEM
// round trips, fences - ~2100 cycles even with the arithmetic removed, r1 timing experiments) bounds the tensor pipe
// at 1024 / (1024 + latency) per tile pair.  Here a block is 64 keys: S is 64 columns, and each query tile owns TWO S
// buffers, so Q K^T of block b+2 goes into the buffer that P V of block b has just released while the softmax
// warps are already working on block b+1.  The tensor pipe always has queued work and the softmax warpgroups run
// back to back: the kernel is bound by throughput (MUFU / tensor), not by the hand-off latency.
//
//   TMEM (512 columns): tile t:  S[t][0] = t*256 + 0, S[t][1] = t*256 + 64, O[t] = t*256 + 128 (128 columns)
//   K / V still arrive as 128-key TMA tiles (2 stages each); block b uses half b % 2 of tile b / 2.
//   The lazy rescale of O must not race with P V of the previous block, which may still be running: that (rare)
//   path first waits on pv_done[t].
// =================================================================================================
constexpr int BKB = 64;  // keys per block
constexpr int kNumBars4 = 1 + 4 * kStages + 4 + 4 + 2 + 1;
constexpr int kSmemBytes4 = kSmemTiles * kTileBytes + kNumBars4 * 8 + 16 + 1024;

template <int kPolyMask8, int kDebug>
__global__ void __launch_bounds__(kThreads, 1) attn_tc_kernel_v4(const __grid_constant__ AttnParams P) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_u32 = ptx::smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_u32 + 1023u) & ~1023u) - raw_u32);
  uint8_t* smem_q = smem;                               // 2 tiles
  uint8_t* smem_k = smem + 2 * kTileBytes;              // kStages tiles
  uint8_t* smem_v = smem + (2 + kStages) * kTileBytes;  // kStages tiles
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kSmemTiles * kTileBytes);
  uint64_t* q_full = bars;
  uint64_t* k_full = bars + 1;
  uint64_t* k_empty = k_full + kStages;
  uint64_t* v_full = k_empty + kStages;
  uint64_t* v_empty = v_full + kStages;
  uint64_t* s_full = v_empty + kStages;  // [tile][buf]
  uint64_t* p_full = s_full + 4;         // [tile][buf]
  uint64_t* pv_done = p_full + 4;        // [tile]
  uint64_t* o_full = pv_done + 2;        // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qp = blockIdx.x % P.n_qpairs;
  const int bh = blockIdx.x / P.n_qpairs;
  const int h = bh % P.heads, b = bh / P.heads;
  const int q0 = qp * 2 * BQ;
  const int n_kv = (P.S + BKV - 1) / BKV;  // 128-key K / V tiles
  const int nb = (P.S + BKB - 1) / BKB;    // 64-key blocks

  if (warp == 0 && lane == 0) ptx::prefetch_tmap(&P.tm);
  if (warp == 1 && lane == 0) {
    ptx::mbar_init(q_full, 1);
    for (int i = 0; i < kStages; ++i) {
      ptx::mbar_init(&k_full[i], 1);
      ptx::mbar_init(&k_empty[i], 1);
      ptx::mbar_init(&v_full[i], 1);
      ptx::mbar_init(&v_empty[i], 1);
    }
    for (int i = 0; i < 4; ++i) {
      ptx::mbar_init(&s_full[i], 1);
      ptx::mbar_init(&p_full[i], 4);  // one elected lane per softmax warp
    }
    ptx::mbar_init(&pv_done[0], 1);
    ptx::mbar_init(&pv_done[1], 1);
    ptx::mbar_init(o_full, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 2) ptx::tmem_alloc<1>(tmem_slot, 512);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
    if (warp == 0 && lane == 0) {
      // ===================== TMA producer (same tiles as v1) =====================
      ptx::mbar_arrive_expect_tx(q_full, 2 * kTileBytes);
#pragma unroll
      for (int t = 0; t < 2; ++t)
#pragma unroll
        for (int sub = 0; sub < 2; ++sub)
          ptx::tma_load_3d(&P.tm, q_full, smem_q + t * kTileBytes + sub * kSubBytes, P.q_col0 + h * HD + sub * 64,
                           q0 + t * BQ, b);
      for (int j = 0; j < n_kv; ++j) {
        const int st = j % kStages, ph = (j / kStages) & 1;
        ptx::mbar_wait(&k_empty[st], ph ^ 1);
        ptx::mbar_arrive_expect_tx(&k_full[st], kTileBytes);
#pragma unroll
        for (int sub = 0; sub < 2; ++sub)
          ptx::tma_load_3d(&P.tm, &k_full[st], smem_k + st * kTileBytes + sub * kSubBytes,
                           P.k_col0 + h * HD + sub * 64, j * BKV, b);
        ptx::mbar_wait(&v_empty[st], ph ^ 1);
        ptx::mbar_arrive_expect_tx(&v_full[st], kTileBytes);
#pragma unroll
        for (int sub = 0; sub < 2; ++sub)
          ptx::tma_load_3d(&P.tm, &v_full[st], smem_v + st * kTileBytes + sub * kSubBytes,
                           P.v_col0 + h * HD + sub * 64, j * BKV, b);
      }
    } else if (warp == 1) {
      // ===================== MMA issuer =====================
      constexpr uint32_t idesc_qk = ptx::make_idesc_bf16(BQ, BKB, 0, 0);  // 128 x 64, A and B K-major
      constexpr uint32_t idesc_pv = ptx::make_idesc_bf16(BQ, HD, 0, 1);   // A (= P) from TMEM, B (= V) MN-major
      const uint64_t q_desc = ptx::make_smem_desc_sw128(ptx::smem_u32(smem_q), 0, 1024);
      const uint64_t k_desc = ptx::make_smem_desc_sw128(ptx::smem_u32(smem_k), 0, 1024);
      const uint64_t v_desc = ptx::make_smem_desc_sw128(ptx::smem_u32(smem_v), kSubBytes, 1024);
      constexpr uint32_t kTile16 = kTileBytes >> 4, kSub16 = kSubBytes >> 4;
      constexpr uint32_t kHalfRows16 = (BKB * 128) >> 4;  // 64 key rows of 128 B inside a sub-tile
      auto s_col = [&](int t, int blk) { return tmem + t * 256 + (blk & 1) * BKB; };
      auto issue_qk = [&](int t, int blk) {
        const int st = (blk >> 1) % kStages;
        const uint64_t qa = q_desc + (uint64_t)(t * kTile16);
        const uint64_t ka = k_desc + (uint64_t)(st * kTile16 + (blk & 1) * kHalfRows16);
#pragma unroll
        for (int kk = 0; kk < HD / 16; ++kk) {
          const uint32_t off = (kk >> 2) * kSub16 + (kk & 3) * 2;  // (addr >> 4) units
          ptx::mma_bf16_ss<1>(s_col(t, blk), qa + off, ka + off, idesc_qk, kk != 0 ? 1u : 0u);
        }
      };
      auto issue_pv = [&](int t, int blk) {
        const int st = (blk >> 1) % kStages;
        const uint64_t va = v_desc + (uint64_t)(st * kTile16);
#pragma unroll
        for (int kk = 0; kk < BKB / 16; ++kk) {
          const int key16 = (blk & 1) * (BKB / 16) + kk;  // 16-key group inside the 128-key V tile
          ptx::mma_bf16_ts(tmem + t * 256 + 128, s_col(t, blk) + kk * 8, va + (uint64_t)(key16 * 128), idesc_pv,
                           (kk != 0 || blk > 0) ? 1u : 0u);
        }
      };
      ptx::mbar_wait(q_full, 0);
      ptx::mbar_wait(&k_full[0], 0);
      ptx::tc_fence_after();
      if (ptx::elect_one()) {
        issue_qk(0, 0);
        ptx::mma_commit(&s_full[0]);
        issue_qk(1, 0);
        ptx::mma_commit(&s_full[2]);
        if (nb > 1) {
          issue_qk(0, 1);
          ptx::mma_commit(&s_full[1]);
          issue_qk(1, 1);
          ptx::mma_commit(&s_full[3]);
        }
        ptx::mma_commit(&k_empty[0]);
      }
      __syncwarp();
      for (int blk = 0; blk < nb; ++blk) {
        const int kt = blk >> 1, st = kt % kStages;
        if ((blk & 1) == 0) ptx::mbar_wait(&v_full[st], (kt / kStages) & 1);
        const int nxt = blk + 2;  // the block whose Q K^T reuses this block's S buffer
        const int kt2 = nxt >> 1, st2 = kt2 % kStages;
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          ptx::mbar_wait(&p_full[t * 2 + (blk & 1)], (blk >> 1) & 1);
          if (nxt < nb && t == 0 && (nxt & 1) == 0) ptx::mbar_wait(&k_full[st2], (kt2 / kStages) & 1);
          ptx::tc_fence_after();
          if (ptx::elect_one()) {
            issue_pv(t, blk);
            ptx::mma_commit(&pv_done[t]);
            if (t == 1 && ((blk & 1) == 1 || blk == nb - 1)) ptx::mma_commit(&v_empty[st]);
            if (nxt < nb) {
              issue_qk(t, nxt);
              ptx::mma_commit(&s_full[t * 2 + (nxt & 1)]);
              if (t == 1 && ((nxt & 1) == 1 || nxt == nb - 1)) ptx::mma_commit(&k_empty[st2]);
            }
          }
          __syncwarp();
        }
      }
      if (ptx::elect_one()) ptx::mma_commit(o_full);
      __syncwarp();
    }
  } else {
    // ===================== softmax warpgroups =====================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 216;");
    const int t = (warp - 4) >> 2;  // 0: tile A, 1: tile B
    const int quad = warp & 3;
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t o_addr = tmem + lane_off + t * 256 + 128;
    const float c = P.scale_log2;
    float m_ref = -INFINITY, l = 0.f;
    for (int blk = 0; blk < nb; ++blk) {
      const int buf = blk & 1;
      const uint32_t s_addr = tmem + lane_off + t * 256 + buf * BKB;
      ptx::mbar_wait(&s_full[t * 2 + buf], (blk >> 1) & 1);
      ptx::tc_fence_after();
      const int n_valid = P.S - blk * BKB;  // < 64 only on the last block
      uint32_t s0[32], s1[32];
      if (kDebug & 1) {
#pragma unroll
        for (int i = 0; i < 32; ++i) s0[i] = s1[i] = 0x3f000000u + (uint32_t)(i + blk);
      } else {
        ptx::tmem_ld_32x32b_x32(s_addr, s0);
        ptx::tmem_ld_32x32b_x32(s_addr + 32, s1);
        ptx::tmem_ld_wait();
      }
      if (n_valid < BKB) {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          if (i >= n_valid) s0[i] = 0xff800000u;  // -inf
          if (32 + i >= n_valid) s1[i] = 0xff800000u;
        }
      }
      float mxa = -INFINITY, mxb = -INFINITY;
#pragma unroll
      for (int i = 0; i < 32; i += 2) {
        mxa = fmaxf(mxa, fmaxf(__uint_as_float(s0[i]), __uint_as_float(s0[i + 1])));
        mxb = fmaxf(mxb, fmaxf(__uint_as_float(s1[i]), __uint_as_float(s1[i + 1])));
      }
      const float mx_s = fmaxf(mxa, mxb) * c;
      if (blk == 0) {
        m_ref = ceilf(mx_s);
      } else if (__any_sync(0xffffffffu, mx_s > m_ref + 8.f)) {
        // O[t] may still be receiving P V of the previous block
        ptx::mbar_wait(&pv_done[t], (blk - 1) & 1);
        ptx::tc_fence_after();
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
      const float2 c2 = make_float2(c, c), nm2 = make_float2(-m_ref, -m_ref);
      float2 lsum = make_float2(0.f, 0.f);
      auto exp_chunk = [&](const uint32_t (&sv)[32], int col) {
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float2 x = __ffma2_rn(make_float2(__uint_as_float(sv[2 * i]), __uint_as_float(sv[2 * i + 1])), c2, nm2);
          float2 e;
          if (kDebug & 2) {
            e = x;
          } else if ((kPolyMask8 >> (i & 7)) & 1) {
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
      exp_chunk(s0, 0);
      exp_chunk(s1, 16);
      l += lsum.x + lsum.y;
      ptx::tmem_st_wait();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&p_full[t * 2 + buf]);
    }
    // ---- epilogue: O / l -> bf16 -> shared (row-wise) -> global (2 rows x 256 B per warp instruction); see v1
    ptx::mbar_wait(o_full, 0);
    ptx::tc_fence_after();
    const float inv = 1.f / l;
    constexpr int kPitch = HD * 2 + 16;
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
    const int row0 = q0 + t * BQ + quad * 32;
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


// =================================================================================================
// v5: ONE 128-row query tile per CTA, 128-key blocks, DOUBLE-BUFFERED score tile.
//
// The two-tile kernels above are bound by the latency of the chain Q K^T -> softmax -> P V -> Q K^T (P aliases S, so
// the next Q K^T of a tile cannot be issued before its P V; ~2100 cycles of hand-offs per block even with the
// arithmetic removed), and v4's 64-key blocks halve the efficiency of the Q K^T instruction.  Here TMEM holds S[0],
// S[1] (128 columns each) and O (128 columns) for a single tile: Q K^T of block j+2 goes into the buffer P V of block j
// has just released, Q K^T of block j+1 is already done when the softmax warps finish block j, so they run back to
// back and the tensor pipe only ever waits for the softmax THROUGHPUT (MUFU / issue), not for a round trip.
//   warp 0 TMA (K ring of 3, V ring of 2), warp 1 MMA, warp 2 TMEM allocator, warps 4-7 softmax (thread = row).
// =================================================================================================
constexpr int kStagesK5 = 2, kStagesV5 = 3;
constexpr int kThreads5 = 256;
constexpr int kSmemTiles5 = 1 + kStagesK5 + kStagesV5;
constexpr int kNumBars5 = 1 + 2 * kStagesK5 + 2 * kStagesV5 + 2 + 2 + 1 + 1;
constexpr int kSmemBytes5 = kSmemTiles5 * kTileBytes + kNumBars5 * 8 + 16 + 1024;

template <int kPolyMask8, int kDebug>
__global__ void __launch_bounds__(kThreads5, 1) attn_tc_kernel_v5(const __grid_constant__ AttnParams P) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_u32 = ptx::smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_u32 + 1023u) & ~1023u) - raw_u32);
  uint8_t* smem_q = smem;                                    // 1 tile
  uint8_t* smem_k = smem + kTileBytes;                       // kStagesK5 tiles
  uint8_t* smem_v = smem + (1 + kStagesK5) * kTileBytes;     // kStagesV5 tiles
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kSmemTiles5 * kTileBytes);
  uint64_t* q_full = bars;
  uint64_t* k_full = bars + 1;
  uint64_t* k_empty = k_full + kStagesK5;
  uint64_t* v_full = k_empty + kStagesK5;
  uint64_t* v_empty = v_full + kStagesV5;
  uint64_t* s_full = v_empty + kStagesV5;  // [2] score buffer written
  uint64_t* p_full = s_full + 2;           // [2] P written into the score buffer
  uint64_t* pv_done = p_full + 2;          // [1] P V of a block completed (lazy-rescale path only)
  uint64_t* o_full = pv_done + 1;          // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_qt = (P.S + BQ - 1) / BQ;
  const int qt = blockIdx.x % n_qt;
  const int bh = blockIdx.x / n_qt;
  const int h = bh % P.heads, b = bh / P.heads;
  const int q0 = qt * BQ;
  const int n_kv = (P.S + BKV - 1) / BKV;

  if (warp == 0 && lane == 0) ptx::prefetch_tmap(&P.tm);
  if (warp == 1 && lane == 0) {
    ptx::mbar_init(q_full, 1);
    for (int i = 0; i < kStagesK5; ++i) { ptx::mbar_init(&k_full[i], 1); ptx::mbar_init(&k_empty[i], 1); }
    for (int i = 0; i < kStagesV5; ++i) { ptx::mbar_init(&v_full[i], 1); ptx::mbar_init(&v_empty[i], 1); }
    for (int i = 0; i < 2; ++i) { ptx::mbar_init(&s_full[i], 1); ptx::mbar_init(&p_full[i], 4); }
    ptx::mbar_init(pv_done, 1);
    ptx::mbar_init(o_full, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 2) ptx::tmem_alloc<1>(tmem_slot, 512);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  constexpr uint32_t kOCol = 256;  // S[0] = 0, S[1] = 128, O = 256

  if (warp == 0) {
    if (lane == 0) {
      // ===================== TMA producer =====================
      ptx::mbar_arrive_expect_tx(q_full, kTileBytes);
#pragma unroll
      for (int sub = 0; sub < 2; ++sub)
        ptx::tma_load_3d(&P.tm, q_full, smem_q + sub * kSubBytes, P.q_col0 + h * HD + sub * 64, q0, b);
      for (int j = 0; j < n_kv; ++j) {
        const int sk = j % kStagesK5, sv = j % kStagesV5;
        ptx::mbar_wait(&k_empty[sk], ((j / kStagesK5) & 1) ^ 1);
        ptx::mbar_arrive_expect_tx(&k_full[sk], kTileBytes);
#pragma unroll
        for (int sub = 0; sub < 2; ++sub)
          ptx::tma_load_3d(&P.tm, &k_full[sk], smem_k + sk * kTileBytes + sub * kSubBytes, P.k_col0 + h * HD + sub * 64,
                           j * BKV, b);
        ptx::mbar_wait(&v_empty[sv], ((j / kStagesV5) & 1) ^ 1);
        ptx::mbar_arrive_expect_tx(&v_full[sv], kTileBytes);
#pragma unroll
        for (int sub = 0; sub < 2; ++sub)
          ptx::tma_load_3d(&P.tm, &v_full[sv], smem_v + sv * kTileBytes + sub * kSubBytes, P.v_col0 + h * HD + sub * 64,
                           j * BKV, b);
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc_qk = ptx::make_idesc_bf16(BQ, BKV, 0, 0);  // A, B K-major
    constexpr uint32_t idesc_pv = ptx::make_idesc_bf16(BQ, HD, 0, 1);   // A from TMEM, B (= V) MN-major
    const uint64_t q_desc = ptx::make_smem_desc_sw128(ptx::smem_u32(smem_q), 0, 1024);
    const uint64_t k_desc = ptx::make_smem_desc_sw128(ptx::smem_u32(smem_k), 0, 1024);
    const uint64_t v_desc = ptx::make_smem_desc_sw128(ptx::smem_u32(smem_v), kSubBytes, 1024);
    constexpr uint32_t kTile16 = kTileBytes >> 4, kSub16 = kSubBytes >> 4;
    auto issue_qk = [&](int j) {
      const uint64_t ka = k_desc + (uint64_t)((j % kStagesK5) * kTile16);
#pragma unroll
      for (int kk = 0; kk < HD / 16; ++kk) {
        const uint32_t off = (kk >> 2) * kSub16 + (kk & 3) * 2;  // (addr >> 4) units
        ptx::mma_bf16_ss<1>(tmem + (j & 1) * 128, q_desc + off, ka + off, idesc_qk, kk != 0 ? 1u : 0u);
      }
    };
    auto issue_pv = [&](int j) {
      const uint64_t va = v_desc + (uint64_t)((j % kStagesV5) * kTile16);
#pragma unroll
      for (int kk = 0; kk < BKV / 16; ++kk)
        ptx::mma_bf16_ts(tmem + kOCol, tmem + (j & 1) * 128 + kk * 8, va + (uint64_t)(kk * 128), idesc_pv,
                         (kk != 0 || j > 0) ? 1u : 0u);
    };
    ptx::mbar_wait(q_full, 0);
    ptx::mbar_wait(&k_full[0], 0);
    ptx::tc_fence_after();
    if (ptx::elect_one()) {
      issue_qk(0);
      ptx::mma_commit(&s_full[0]);
      ptx::mma_commit(&k_empty[0]);
    }
    __syncwarp();
    if (n_kv > 1) {
      ptx::mbar_wait(&k_full[1], 0);
      ptx::tc_fence_after();
      if (ptx::elect_one()) {
        issue_qk(1);
        ptx::mma_commit(&s_full[1]);
        ptx::mma_commit(&k_empty[1]);
      }
      __syncwarp();
    }
    for (int j = 0; j < n_kv; ++j) {
      ptx::mbar_wait(&v_full[j % kStagesV5], (j / kStagesV5) & 1);
      ptx::mbar_wait(&p_full[j & 1], (j >> 1) & 1);
      ptx::tc_fence_after();
      if (ptx::elect_one()) {
        issue_pv(j);
        ptx::mma_commit(&v_empty[j % kStagesV5]);
        ptx::mma_commit(pv_done);
      }
      __syncwarp();
      const int nx = j + 2;  // its scores reuse the buffer P V of block j has just consumed
      if (nx < n_kv) {
        ptx::mbar_wait(&k_full[nx % kStagesK5], (nx / kStagesK5) & 1);
        ptx::tc_fence_after();
        if (ptx::elect_one()) {
          issue_qk(nx);
          ptx::mma_commit(&s_full[nx & 1]);
          ptx::mma_commit(&k_empty[nx % kStagesK5]);
        }
        __syncwarp();
      }
    }
    if (ptx::elect_one()) ptx::mma_commit(o_full);
    __syncwarp();
  } else if (warp >= 4) {
    // ===================== softmax warpgroup =====================
    const int quad = warp & 3;
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t o_addr = tmem + lane_off + kOCol;
    const float c = P.scale_log2;
    float m_ref = -INFINITY, l = 0.f;
    for (int j = 0; j < n_kv; ++j) {
      const uint32_t s_addr = tmem + lane_off + (j & 1) * 128;
      ptx::mbar_wait(&s_full[j & 1], (j >> 1) & 1);
      ptx::tc_fence_after();
      const int n_valid = P.S - j * BKV;  // < 128 only on the last block
      uint32_t s0[32], s1[32], s2[32], s3[32];
      if (kDebug & 1) {
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
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        mx0 = fmaxf(mx0, __uint_as_float(s0[i]));
        mx1 = fmaxf(mx1, __uint_as_float(s1[i]));
        mx2 = fmaxf(mx2, __uint_as_float(s2[i]));
        mx3 = fmaxf(mx3, __uint_as_float(s3[i]));
      }
      const float mx_s = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * c;
      if (j == 0) {
        m_ref = ceilf(mx_s);
      } else if (__any_sync(0xffffffffu, mx_s > m_ref + 8.f)) {
        ptx::mbar_wait(pv_done, (j - 1) & 1);  // O may still be receiving P V of the previous block
        ptx::tc_fence_after();
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
      const float2 c2 = make_float2(c, c), nm2 = make_float2(-m_ref, -m_ref);
      float2 lsum = make_float2(0.f, 0.f);
      auto exp_chunk = [&](const uint32_t (&sv)[32], int col) {
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float2 x = __ffma2_rn(make_float2(__uint_as_float(sv[2 * i]), __uint_as_float(sv[2 * i + 1])), c2, nm2);
          float2 e;
          if (kDebug & 2) {
            e = x;
          } else if ((kPolyMask8 >> (i & 7)) & 1) {
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
      exp_chunk(s0, 0);
      exp_chunk(s1, 16);
      exp_chunk(s2, 32);
      exp_chunk(s3, 48);
      l += lsum.x + lsum.y;
      ptx::tmem_st_wait();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&p_full[j & 1]);
    }
    // ---- epilogue: O / l -> bf16 -> shared (row-wise) -> global (2 rows x 256 B per warp instruction)
    ptx::mbar_wait(o_full, 0);
    ptx::tc_fence_after();
    const float inv = 1.f / l;
    constexpr int kPitch = HD * 2 + 16;
    uint8_t* stage = smem + (warp - 4) * (32 * kPitch);  // Q and K tiles are dead once o_full has fired
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
    const int row0 = q0 + quad * 32;
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


// =================================================================================================
// v6: v5's structure on a CTA PAIR (tcgen05 cta_group::2).
//
// Timing experiments (r1) show every variant above ends at the same ~80 B/clk of shared-memory traffic per SM (MMA
// operand reads + TMA writes): two query tiles per CTA read Q and K 4 KB each per MMA, one tile per CTA re-loads K / V
// for half the rows.  With a 2-CTA MMA (M = 256) each CTA still owns ONE query tile (TMEM: S[0], S[1], O, so the
// score tile is double-buffered and the softmax runs back to back, as in v5), but the B operands are split across the
// pair: each CTA loads and holds only 64 of the 128 keys of a K tile and 64 of the 128 head-dim columns of a V tile.
// Per CTA and 128-key block: Q K^T reads 6 KB per MMA instead of 8, P V reads 2 KB instead of 4, TMA writes 32 KB
// instead of 64 - 96 KB of shared-memory traffic per tile and block against 128 KB (two-tile kernel) / 160 KB (v5).
//   leader CTA: issues every MMA for the pair; both CTAs: TMA producer for their halves, softmax warpgroup (thread =
//   row) for their own 128 query rows, P hand-off into the LEADER's barriers (remote arrive from the peer).
// =================================================================================================
constexpr int kStagesK6 = 4, kStagesV6 = 4;
constexpr int kHalfTileBytes = kTileBytes / 2;  // 16 KB: half a K tile (64 keys) or half a V tile (64 columns)
constexpr int kNumBars6 = 1 + 2 * kStagesK6 + 2 * kStagesV6 + 3 + 3 + 2 + 1 + 2;
constexpr int kSmemBytes6 = kTileBytes + (kStagesK6 + kStagesV6) * kHalfTileBytes + kNumBars6 * 8 + 16 + 1024;

template <int kPolyMask8, int kDebug, bool kTwoIssuers = false, int kNS = 2>
__global__ void __launch_bounds__(kThreads5, 1) attn_tc_kernel_v6(const __grid_constant__ AttnParams P) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_u32 = ptx::smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_u32 + 1023u) & ~1023u) - raw_u32);
  uint8_t* smem_q = smem;                                                  // this CTA's 128 x 128 query tile
  uint8_t* smem_k = smem + kTileBytes;                                     // kStagesK6 x [64 keys x 128]
  uint8_t* smem_v = smem + kTileBytes + kStagesK6 * kHalfTileBytes;        // kStagesV6 x [128 keys x 64 columns]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kTileBytes + (kStagesK6 + kStagesV6) * kHalfTileBytes);
  uint64_t* q_full = bars;                 // leader's: both CTAs' TMA bytes
  uint64_t* k_full = bars + 1;             // leader's
  uint64_t* k_empty = k_full + kStagesK6;  // each CTA's own (multicast commit)
  uint64_t* v_full = k_empty + kStagesK6;  // leader's
  uint64_t* v_empty = v_full + kStagesV6;  // each CTA's own
  uint64_t* s_full = v_empty + kStagesV6;  // [kNS] each CTA's own (multicast commit)
  uint64_t* p_full = s_full + 3;           // [kNS] leader's: 4 warps of each CTA
  uint64_t* pv_done = p_full + 3;          // [2] each CTA's own; P V of block j commits pv_done[j & 1]
  uint64_t* o_full = pv_done + 2;          // [1] each CTA's own
  uint64_t* s_free = o_full + 1;           // [2] leader's: P V of a block has consumed the score buffer (kTwoIssuers)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_free + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t cta_rank = ptx::cluster_ctarank();
  const bool leader = cta_rank == 0;
  const int n_qp = (P.S + 2 * BQ - 1) / (2 * BQ);
  const int pair = blockIdx.x >> 1;
  const int qp = pair % n_qp;
  const int bh = pair / n_qp;
  const int h = bh % P.heads, b = bh / P.heads;
  const int q0 = qp * 2 * BQ + (int)cta_rank * BQ;  // this CTA's query rows
  const int n_kv = (P.S + BKV - 1) / BKV;

  if (warp == 0 && lane == 0) ptx::prefetch_tmap(&P.tm);
  if (warp == 1 && lane == 0) {
    ptx::mbar_init(q_full, 1);
    for (int i = 0; i < kStagesK6; ++i) { ptx::mbar_init(&k_full[i], 1); ptx::mbar_init(&k_empty[i], 1); }
    for (int i = 0; i < kStagesV6; ++i) { ptx::mbar_init(&v_full[i], 1); ptx::mbar_init(&v_empty[i], 1); }
    for (int i = 0; i < 3; ++i) { ptx::mbar_init(&s_full[i], 1); ptx::mbar_init(&p_full[i], 8); }
    ptx::mbar_init(&pv_done[0], 1);
    ptx::mbar_init(&pv_done[1], 1);
    ptx::mbar_init(o_full, 1);
    ptx::mbar_init(&s_free[0], 1);
    ptx::mbar_init(&s_free[1], 1);
    ptx::fence_barrier_init();
  }
  if (warp == 2) ptx::tmem_alloc<2>(tmem_slot, 512);
  ptx::tc_fence_before();
  ptx::cluster_sync();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  constexpr uint32_t kOCol = kNS * 128;  // S[i] = i * 128, O behind them (kNS = 3 fills all 512 columns)
  static_assert(kNS == 2 || kNS == 3, "2 or 3 score buffers");
  static_assert(!kTwoIssuers || kNS == 2, "the two-issuer form is written for 2 score buffers");

  if (warp == 0) {
    if (lane == 0) {
      // ===================== TMA producer (both CTAs; the bytes land on the leader's full barriers) ==========
      if (leader) ptx::mbar_arrive_expect_tx(q_full, 2 * kTileBytes);
#pragma unroll
      for (int sub = 0; sub < 2; ++sub)
        ptx::tma_load_3d_2sm(&P.tm, q_full, smem_q + sub * kSubBytes, P.q_col0 + h * HD + sub * 64, q0, b);
      for (int j = 0; j < n_kv; ++j) {
        const int sk = j % kStagesK6, sv = j % kStagesV6;
        // K: this CTA's 64 keys (rows) of the tile, both 64-column halves of the head dimension: box (64, 64)
        ptx::mbar_wait(&k_empty[sk], ((j / kStagesK6) & 1) ^ 1);
        if (leader) ptx::mbar_arrive_expect_tx(&k_full[sk], 2 * kHalfTileBytes);
#pragma unroll
        for (int sub = 0; sub < 2; ++sub)
          ptx::tma_load_3d_2sm(&P.tmh, &k_full[sk], smem_k + sk * kHalfTileBytes + sub * (kHalfTileBytes / 2),
                               P.k_col0 + h * HD + sub * 64, j * BKV + (int)cta_rank * (BKV / 2), b);
        // V: all 128 keys, this CTA's 64 head-dim columns: box (64, 128)
        ptx::mbar_wait(&v_empty[sv], ((j / kStagesV6) & 1) ^ 1);
        if (leader) ptx::mbar_arrive_expect_tx(&v_full[sv], 2 * kHalfTileBytes);
        ptx::tma_load_3d_2sm(&P.tm, &v_full[sv], smem_v + sv * kHalfTileBytes, P.v_col0 + h * HD + (int)cta_rank * 64,
                             j * BKV, b);
      }
    }
  } else if (warp == 1) {
    if (leader) {
      // ===================== MMA issuer (leader CTA, for the pair) =====================
      constexpr uint32_t idesc_qk = ptx::make_idesc_bf16(2 * BQ, BKV, 0, 0);  // 256 x 128, A and B K-major
      constexpr uint32_t idesc_pv = ptx::make_idesc_bf16(2 * BQ, HD, 0, 1);   // A (= P) from TMEM, B (= V) MN-major
      const uint64_t q_desc = ptx::make_smem_desc_sw128(ptx::smem_u32(smem_q), 0, 1024);
      const uint64_t k_desc = ptx::make_smem_desc_sw128(ptx::smem_u32(smem_k), 0, 1024);
      const uint64_t v_desc = ptx::make_smem_desc_sw128(ptx::smem_u32(smem_v), 0, 1024);
      constexpr uint32_t kHalf16 = kHalfTileBytes >> 4, kQSub16 = kSubBytes >> 4, kKSub16 = (kHalfTileBytes / 2) >> 4;
      auto issue_qk = [&](int j) {
        const uint64_t ka = k_desc + (uint64_t)((j % kStagesK6) * kHalf16);
#pragma unroll
        for (int kk = 0; kk < HD / 16; ++kk) {
          const uint32_t qoff = (kk >> 2) * kQSub16 + (kk & 3) * 2;  // (addr >> 4) units
          const uint32_t koff = (kk >> 2) * kKSub16 + (kk & 3) * 2;
          ptx::mma_bf16_ss<2>(tmem + (j % kNS) * 128, q_desc + qoff, ka + koff, idesc_qk, kk != 0 ? 1u : 0u);
        }
      };
      auto issue_pv = [&](int j) {
        const uint64_t va = v_desc + (uint64_t)((j % kStagesV6) * kHalf16);
#pragma unroll
        for (int kk = 0; kk < BKV / 16; ++kk)
          ptx::mma_bf16_ts_2sm(tmem + kOCol, tmem + (j % kNS) * 128 + kk * 8, va + (uint64_t)(kk * 128), idesc_pv,
                               (kk != 0 || j > 0) ? 1u : 0u);
      };
      ptx::mbar_wait(q_full, 0);
      for (int j0 = 0; j0 < kNS && j0 < n_kv; ++j0) {
        ptx::mbar_wait(&k_full[j0 % kStagesK6], (j0 / kStagesK6) & 1);
        ptx::tc_fence_after();
        if (ptx::elect_one()) {
          issue_qk(j0);
          ptx::mma_commit_2sm(&s_full[j0 % kNS], 3);
          ptx::mma_commit_2sm(&k_empty[j0 % kStagesK6], 3);
        }
        __syncwarp();
      }
      if constexpr (kTwoIssuers) {
        // this warp only issues Q K^T; warp 3 issues P V and tells us (s_free) when a score buffer may be overwritten
        for (int nx = 2; nx < n_kv; ++nx) {
          ptx::mbar_wait(&k_full[nx % kStagesK6], (nx / kStagesK6) & 1);
          ptx::mbar_wait(&s_free[nx & 1], ((nx >> 1) - 1) & 1);
          ptx::tc_fence_after();
          if (ptx::elect_one()) {
            issue_qk(nx);
            ptx::mma_commit_2sm(&s_full[nx & 1], 3);
            ptx::mma_commit_2sm(&k_empty[nx % kStagesK6], 3);
          }
          __syncwarp();
        }
      } else {
      for (int j = 0; j < n_kv; ++j) {
        ptx::mbar_wait(&v_full[j % kStagesV6], (j / kStagesV6) & 1);
        ptx::mbar_wait(&p_full[j % kNS], (j / kNS) & 1);
        ptx::tc_fence_after();
        if (ptx::elect_one()) {
          issue_pv(j);
          ptx::mma_commit_2sm(&v_empty[j % kStagesV6], 3);
          ptx::mma_commit_2sm(&pv_done[j & 1], 3);
        }
        __syncwarp();
        const int nx = j + kNS;  // its scores reuse the buffer P V of block j has just consumed
        if (nx < n_kv) {
          ptx::mbar_wait(&k_full[nx % kStagesK6], (nx / kStagesK6) & 1);
          ptx::tc_fence_after();
          if (ptx::elect_one()) {
            issue_qk(nx);
            ptx::mma_commit_2sm(&s_full[nx % kNS], 3);
            ptx::mma_commit_2sm(&k_empty[nx % kStagesK6], 3);
          }
          __syncwarp();
        }
      }
      }
      if constexpr (!kTwoIssuers) {
        if (ptx::elect_one()) ptx::mma_commit_2sm(o_full, 3);
        __syncwarp();
      }
    }
  } else if (warp == 3) {
    if constexpr (kTwoIssuers) {
      if (leader) {
        // ===================== second MMA issuer: P V =====================
        constexpr uint32_t idesc_pv = ptx::make_idesc_bf16(2 * BQ, HD, 0, 1);
        const uint64_t v_desc = ptx::make_smem_desc_sw128(ptx::smem_u32(smem_v), 0, 1024);
        constexpr uint32_t kHalf16 = kHalfTileBytes >> 4;
        for (int j = 0; j < n_kv; ++j) {
          ptx::mbar_wait(&v_full[j % kStagesV6], (j / kStagesV6) & 1);
          ptx::mbar_wait(&p_full[j & 1], (j >> 1) & 1);
          ptx::tc_fence_after();
          if (ptx::elect_one()) {
            const uint64_t va = v_desc + (uint64_t)((j % kStagesV6) * kHalf16);
#pragma unroll
            for (int kk = 0; kk < BKV / 16; ++kk)
              ptx::mma_bf16_ts_2sm(tmem + kOCol, tmem + (j & 1) * 128 + kk * 8, va + (uint64_t)(kk * 128), idesc_pv,
                                   (kk != 0 || j > 0) ? 1u : 0u);
            ptx::mma_commit(&s_free[j & 1]);  // leader-local: the Q K^T warp may reuse this score buffer
            ptx::mma_commit_2sm(&v_empty[j % kStagesV6], 3);
            ptx::mma_commit_2sm(&pv_done[j & 1], 3);
          }
          __syncwarp();
        }
        if (ptx::elect_one()) ptx::mma_commit_2sm(o_full, 3);
        __syncwarp();
      }
    }
  } else if (warp >= 4) {
    // ===================== softmax warpgroup (this CTA's 128 query rows) =====================
    const int quad = warp & 3;
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t o_addr = tmem + lane_off + kOCol;
    const float c = P.scale_log2;
    float m_ref = -INFINITY, l = 0.f;
    for (int j = 0; j < n_kv; ++j) {
      const uint32_t s_addr = tmem + lane_off + (j % kNS) * 128;
      ptx::mbar_wait(&s_full[j % kNS], (j / kNS) & 1);
      ptx::tc_fence_after();
      const int n_valid = P.S - j * BKV;  // < 128 only on the last block
      uint32_t s0[32], s1[32], s2[32], s3[32];
      if (kDebug & 1) {
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
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        mx0 = fmaxf(mx0, __uint_as_float(s0[i]));
        mx1 = fmaxf(mx1, __uint_as_float(s1[i]));
        mx2 = fmaxf(mx2, __uint_as_float(s2[i]));
        mx3 = fmaxf(mx3, __uint_as_float(s3[i]));
      }
      const float mx_s = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * c;
      if (j == 0) {
        m_ref = ceilf(mx_s);
      } else if (__any_sync(0xffffffffu, mx_s > m_ref + 8.f)) {
        // O may still be receiving P V of the previous block(s): with kNS score buffers up to kNS - 1 of them are
        // outstanding.  One barrier per block parity keeps every wait within one phase of its barrier.
        ptx::mbar_wait(&pv_done[(j - 1) & 1], ((j - 1) >> 1) & 1);
        if (j >= 2) ptx::mbar_wait(&pv_done[j & 1], ((j - 2) >> 1) & 1);
        ptx::tc_fence_after();
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
      const float2 c2 = make_float2(c, c), nm2 = make_float2(-m_ref, -m_ref);
      float2 lsum = make_float2(0.f, 0.f);
      auto exp_chunk = [&](const uint32_t (&sv)[32], int col) {
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          float2 x = __ffma2_rn(make_float2(__uint_as_float(sv[2 * i]), __uint_as_float(sv[2 * i + 1])), c2, nm2);
          float2 e;
          if (kDebug & 2) {
            e = x;
          } else if ((kPolyMask8 >> (i & 7)) & 1) {
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
      exp_chunk(s0, 0);
      exp_chunk(s1, 16);
      exp_chunk(s2, 32);
      exp_chunk(s3, 48);
      l += lsum.x + lsum.y;
      ptx::tmem_st_wait();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (leader) ptx::mbar_arrive(&p_full[j % kNS]);
        else ptx::mbar_arrive_cluster_relaxed(&p_full[j % kNS], 0);
      }
    }
    // ---- epilogue: O / l -> bf16 -> shared (row-wise) -> global (2 rows x 256 B per warp instruction)
    ptx::mbar_wait(o_full, 0);
    ptx::tc_fence_after();
    const float inv = 1.f / l;
    constexpr int kPitch = HD * 2 + 16;
    uint8_t* stage = smem + (warp - 4) * (32 * kPitch);  // Q and K tiles are dead once o_full has fired
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
    const int row0 = q0 + quad * 32;
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
  ptx::cluster_sync();
  if (warp == 2) ptx::tmem_dealloc<2>(tmem, 512);
}


// =================================================================================================
// PAIR kernel: the two-tile ping-pong of attn_tc_kernel on a CTA PAIR (tcgen05 cta_group::2).
//
// Why.  With one CTA per SM every Q K^T instruction (128 x 128 x 16, both operands from shared memory) reads 8 KB per 64
// tensor cycles = 128 B/clk - the whole bandwidth of the shared-memory crossbar - while TMA keeps writing the next K / V
// tiles into the same memory: the r1 timing experiments found every single-CTA variant, arithmetic removed, pinned at
// 1320-1460 TFLOP/s for exactly that reason, and cuDNN's Blackwell kernel measures 1510-1540 on the same box and shapes
// (profiles/r2_attn_lib_compare.txt).  A 2-CTA MMA (M = 256: 128 query rows in each CTA) takes its B operand HALF from
// each CTA: per CTA a Q K^T instruction reads 4 KB of Q + 2 KB of K, a P V instruction (P from TMEM) 2 KB of V, and TMA
// writes 16 + 16 KB per 128-key block instead of 32 + 32.  Shared-memory traffic per CTA and block: 160 KB instead of
// 256 KB (78 B/clk at full tensor rate instead of 125).
// v6 (above, A/B only) had the pair but ONE query tile per CTA, which leaves the chain Q K^T -> softmax -> P V of a tile
// exposed; here each CTA keeps the product's two tiles (A, B) and the tensor pipe ping-pongs between MMA-tile A (= tile
// A of both CTAs, 256 rows) and MMA-tile B.
//
//   cluster = 2 CTAs = 4 query tiles (512 rows) of one (batch, head); the LEADER (cluster rank 0) issues every MMA
//   both CTAs: warp 0 TMA producer for their halves - Q: own two tiles; K: 64 of the block's 128 keys; V: 64 of the
//              128 head-dim columns - with the bytes counted on the LEADER's full barriers (cta_group::2 loads);
//              warps 4-7 / 8-11 softmax warpgroups of their own tiles A / B (thread = query row), P handed over by
//              one elected lane per warp arriving on the LEADER's p_half / p_full barriers (remote arrive from the peer)
//   TMEM (both CTAs, same columns): S_A | S_B | O_A | O_B, 128 fp32 columns each; bf16 P overwrites S
//   tcgen05.commit multicasts to both CTAs: each CTA's own k_empty / v_empty / s_full / o_full barriers
// =================================================================================================
constexpr int kStagesK7 = 4, kStagesV7 = 4;
constexpr int kHalfBytes7 = kTileBytes / 2;  // 16 KB: 64 keys x 128 (K half) or 128 keys x 64 columns (V half)
constexpr int kNumBars7 = 1 + 2 * kStagesK7 + 2 * kStagesV7 + 2 + 2 + 2 + 1;
constexpr int kSmemBytes7 = 2 * kTileBytes + (kStagesK7 + kStagesV7) * kHalfBytes7 + kNumBars7 * 8 + 16 + 1024;
static_assert(kNumBars7 * 8 + 16 <= 256, "barrier block");
static_assert(8 * 32 * (HD * 2 + 16) <= 2 * kTileBytes + kStagesK7 * kHalfBytes7, "epilogue staging fits in Q + K");

template <int kPolyMask8>
__global__ void __launch_bounds__(kThreads, 1) attn_tc_pair_kernel(const __grid_constant__ AttnParams P) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_u32 = ptx::smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_u32 + 1023u) & ~1023u) - raw_u32);
  uint8_t* smem_q = smem;                                              // this CTA's tiles A, B
  uint8_t* smem_k = smem + 2 * kTileBytes;                             // kStagesK7 x [64 keys x 128]
  uint8_t* smem_v = smem_k + kStagesK7 * kHalfBytes7;                  // kStagesV7 x [128 keys x 64 columns]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_v + kStagesV7 * kHalfBytes7);
  uint64_t* q_full = bars;                 // leader's: both CTAs' TMA bytes
  uint64_t* k_full = bars + 1;             // leader's
  uint64_t* k_empty = k_full + kStagesK7;  // each CTA's own (multicast commit)
  uint64_t* v_full = k_empty + kStagesK7;  // leader's
  uint64_t* v_empty = v_full + kStagesV7;  // each CTA's own
  uint64_t* s_full = v_empty + kStagesV7;  // [2] each CTA's own
  uint64_t* p_full = s_full + 2;           // [2] leader's: one lane of each of the 8 softmax warps of an MMA tile
  uint64_t* p_half = p_full + 2;           // [2] leader's: first 64 keys of P written
  uint64_t* o_full = p_half + 2;           // [1] each CTA's own
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t cta_rank = ptx::cluster_ctarank();
  const bool leader = cta_rank == 0;
  const int n_qq = (P.S + 4 * BQ - 1) / (4 * BQ);  // clusters per (batch, head)
  const int cl = blockIdx.x >> 1;
  const int qq = cl % n_qq;
  const int bh = cl / n_qq;
  const int h = bh % P.heads, b = bh / P.heads;
  const int q0 = qq * 4 * BQ + (int)cta_rank * 2 * BQ;  // this CTA's 256 query rows
  const int n_kv = (P.S + BKV - 1) / BKV;

  if (warp == 0 && lane == 0) {
    ptx::prefetch_tmap(&P.tm);
    ptx::prefetch_tmap(&P.tmh);
  }
  if (warp == 1 && lane == 0) {
    ptx::mbar_init(q_full, 1);
    for (int i = 0; i < kStagesK7; ++i) { ptx::mbar_init(&k_full[i], 1); ptx::mbar_init(&k_empty[i], 1); }
    for (int i = 0; i < kStagesV7; ++i) { ptx::mbar_init(&v_full[i], 1); ptx::mbar_init(&v_empty[i], 1); }
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&s_full[i], 1);
      ptx::mbar_init(&p_full[i], 8);
      ptx::mbar_init(&p_half[i], 8);
    }
    ptx::mbar_init(o_full, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 2) ptx::tmem_alloc<2>(tmem_slot, 512);
  ptx::tc_fence_before();
  ptx::cluster_sync();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  ptx::grid_launch_dependents();  // programmatic dependent launch: nothing above reads the previous kernel's output
  ptx::grid_dependency_wait();

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
    if (warp == 0 && lane == 0) {
      // ===================== TMA producer (both CTAs; the bytes land on the leader's full barriers) ==========
      if (leader) ptx::mbar_arrive_expect_tx(q_full, 4 * kTileBytes);
#pragma unroll
      for (int t = 0; t < 2; ++t)
#pragma unroll
        for (int sub = 0; sub < 2; ++sub)
          ptx::tma_load_3d_2sm(&P.tm, q_full, smem_q + t * kTileBytes + sub * kSubBytes,
                               P.q_col0 + h * HD + sub * 64, q0 + t * BQ, b);
      for (int j = 0; j < n_kv; ++j) {
        const int sk = j % kStagesK7, sv = j % kStagesV7;
        // K: this CTA's 64 keys of the block, both 64-column halves of the head dimension: box (64, 64)
        ptx::mbar_wait(&k_empty[sk], ((j / kStagesK7) & 1) ^ 1);
        if (leader) ptx::mbar_arrive_expect_tx(&k_full[sk], 2 * kHalfBytes7);
#pragma unroll
        for (int sub = 0; sub < 2; ++sub)
          ptx::tma_load_3d_2sm(&P.tmh, &k_full[sk], smem_k + sk * kHalfBytes7 + sub * (kHalfBytes7 / 2),
                               P.k_col0 + h * HD + sub * 64, j * BKV + (int)cta_rank * (BKV / 2), b);
        // V: all 128 keys, this CTA's 64 head-dim columns: box (64, 128)
        ptx::mbar_wait(&v_empty[sv], ((j / kStagesV7) & 1) ^ 1);
        if (leader) ptx::mbar_arrive_expect_tx(&v_full[sv], 2 * kHalfBytes7);
        ptx::tma_load_3d_2sm(&P.tm, &v_full[sv], smem_v + sv * kHalfBytes7, P.v_col0 + h * HD + (int)cta_rank * 64,
                             j * BKV, b);
      }
    } else if (warp == 1 && leader) {
      // ===================== MMA issuer (leader CTA, for the pair) =====================
      // The whole warp runs the loop (descriptor arithmetic stays on the uniform datapath); one elected lane issues.
      constexpr uint32_t idesc_qk = ptx::make_idesc_bf16(2 * BQ, BKV, 0, 0);  // 256 x 128, A and B K-major
      constexpr uint32_t idesc_pv = ptx::make_idesc_bf16(2 * BQ, HD, 0, 1);   // A (= P) from TMEM, B (= V) MN-major
      const uint64_t q_desc = ptx::make_smem_desc_sw128(ptx::smem_u32(smem_q), 0, 1024);
      const uint64_t k_desc = ptx::make_smem_desc_sw128(ptx::smem_u32(smem_k), 0, 1024);
      const uint64_t v_desc = ptx::make_smem_desc_sw128(ptx::smem_u32(smem_v), 0, 1024);
      constexpr uint32_t kTile16 = kTileBytes >> 4, kHalf16 = kHalfBytes7 >> 4;
      constexpr uint32_t kQSub16 = kSubBytes >> 4, kKSub16 = (kHalfBytes7 / 2) >> 4;
      auto issue_qk = [&](int t, int sk) {
        const uint64_t qa = q_desc + (uint64_t)(t * kTile16), ka = k_desc + (uint64_t)(sk * kHalf16);
#pragma unroll
        for (int kk = 0; kk < HD / 16; ++kk) {
          const uint32_t qoff = (kk >> 2) * kQSub16 + (kk & 3) * 2;  // (addr >> 4) units
          const uint32_t koff = (kk >> 2) * kKSub16 + (kk & 3) * 2;
          ptx::mma_bf16_ss<2>(tmem + t * 128, qa + qoff, ka + koff, idesc_qk, kk != 0 ? 1u : 0u);
        }
      };
      auto issue_pv = [&](int t, int sv, uint32_t acc, int kk0, int kk1) {
        const uint64_t va = v_desc + (uint64_t)(sv * kHalf16);
#pragma unroll
        for (int kk = kk0; kk < kk1; ++kk)  // 16 keys = 16 rows of 128 B in this CTA's 64-column half
          ptx::mma_bf16_ts_2sm(tmem + 256 + t * 128, tmem + t * 128 + kk * 8, va + (uint64_t)(kk * 128), idesc_pv,
                               kk != 0 ? 1u : acc);
      };
      ptx::mbar_wait(q_full, 0);
      ptx::mbar_wait(&k_full[0], 0);
      ptx::tc_fence_after();
      if (ptx::elect_one()) {
        issue_qk(0, 0);
        ptx::mma_commit_2sm(&s_full[0], 3);
        issue_qk(1, 0);
        ptx::mma_commit_2sm(&s_full[1], 3);
        ptx::mma_commit_2sm(&k_empty[0], 3);
      }
      __syncwarp();
      for (int j = 0; j < n_kv; ++j) {
        const int sv = j % kStagesV7, phv = (j / kStagesV7) & 1;
        const int nsk = (j + 1) % kStagesK7, nphk = ((j + 1) / kStagesK7) & 1;
        const bool more = j + 1 < n_kv;
        const uint32_t acc = j > 0 ? 1u : 0u;
        ptx::mbar_wait(&v_full[sv], phv);
        ptx::mbar_wait(&p_half[0], j & 1);
        ptx::tc_fence_after();
        if (ptx::elect_one()) issue_pv(0, sv, acc, 0, 4);
        __syncwarp();
        ptx::mbar_wait(&p_full[0], j & 1);
        if (more) ptx::mbar_wait(&k_full[nsk], nphk);
        ptx::tc_fence_after();
        if (ptx::elect_one()) {
          issue_pv(0, sv, 1u, 4, 8);
          if (more) {
            issue_qk(0, nsk);
            ptx::mma_commit_2sm(&s_full[0], 3);
          }
        }
        __syncwarp();
        ptx::mbar_wait(&p_half[1], j & 1);
        ptx::tc_fence_after();
        if (ptx::elect_one()) issue_pv(1, sv, acc, 0, 4);
        __syncwarp();
        ptx::mbar_wait(&p_full[1], j & 1);
        ptx::tc_fence_after();
        if (ptx::elect_one()) {
          issue_pv(1, sv, 1u, 4, 8);
          ptx::mma_commit_2sm(&v_empty[sv], 3);
          if (more) {
            issue_qk(1, nsk);
            ptx::mma_commit_2sm(&s_full[1], 3);
            ptx::mma_commit_2sm(&k_empty[nsk], 3);
          }
        }
        __syncwarp();
      }
      if (ptx::elect_one()) ptx::mma_commit_2sm(o_full, 3);
      __syncwarp();
    }
  } else {
    // ===================== softmax warpgroups (this CTA's tiles A, B) =====================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 216;");
    const int t = (warp - 4) >> 2;  // 0: tile A, 1: tile B
    const int quad = warp & 3;
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t s_addr = tmem + lane_off + t * 128;
    const uint32_t o_addr = tmem + lane_off + 256 + t * 128;
    const float c = P.scale_log2;
    float m_ref = -INFINITY, l = 0.f;
    auto hand_over = [&](uint64_t* bar) {  // P (or its first half) is in TMEM: tell the leader's MMA warp
      ptx::tmem_st_wait();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (leader) ptx::mbar_arrive(bar);
        else ptx::mbar_arrive_cluster_relaxed(bar, 0);
      }
    };
    for (int j = 0; j < n_kv; ++j) {
      ptx::mbar_wait(&s_full[t], j & 1);
      ptx::tc_fence_after();
      const int n_valid = P.S - j * BKV;  // < 128 only on the last block
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
      ptx::tmem_ld_32x32b_x32(s_addr, s0);
      ptx::tmem_ld_wait();
      ptx::tmem_ld_32x32b_x32(s_addr + 32, s1);
      float mx = chunk_max(s0, 0);
      ptx::tmem_ld_wait();
      ptx::tmem_ld_32x32b_x32(s_addr + 64, s2);
      mx = fmaxf(mx, chunk_max(s1, 32));
      ptx::tmem_ld_wait();
      ptx::tmem_ld_32x32b_x32(s_addr + 96, s3);
      mx = fmaxf(mx, chunk_max(s2, 64));
      ptx::tmem_ld_wait();
      mx = fmaxf(mx, chunk_max(s3, 96));
      const float mx_s = mx * c;
      if (j == 0) {
        m_ref = ceilf(mx_s);
      } else if (__any_sync(0xffffffffu, mx_s > m_ref + 8.f)) {
        // lazy rescale: s_full[t] of block j was committed after P V of block j - 1 (in-order pipe), so O is quiescent
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
      exp_chunk(s0, 0);
      exp_chunk(s1, 16);
      hand_over(&p_half[t]);
      exp_chunk(s2, 32);
      exp_chunk(s3, 48);
      l += lsum.x + lsum.y;
      hand_over(&p_full[t]);
    }
    // ---- epilogue: O / l -> bf16 -> shared (row-wise) -> global (2 rows x 256 B per warp instruction); all MMAs of
    // the pair have completed (o_full), so this CTA's Q tiles and K ring are dead
    ptx::mbar_wait(o_full, 0);
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
  ptx::cluster_sync();
  if (warp == 2) ptx::tmem_dealloc<2>(tmem, 512);
}

