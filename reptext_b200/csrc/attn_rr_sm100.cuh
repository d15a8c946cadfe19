// Round-robin decoupled joint attention ("v9"; included by attn_sm100.cu under RT_AB_VARIANTS; same AttnParams / results
// contract as the product).
//
// Same decoupling as attn_decoupled_sm100.cuh - P through shared memory (SS-mode P V), S released as soon as the softmax
// threads have consumed it, separate Q K^T and P V issuing warps - but ONE thread per query row again, so no row-maximum
// exchange between threads: each query tile has TWO softmax warpgroups that take its key blocks in turn (even / odd).
// Four softmax warps per scheduler, each with two block periods for one block, at different phases - the loads, the
// reference hand-off, the stores and the exponentials of different warps overlap.
//   * the running reference exponent travels block to block through shared memory (mref, one-way: written after the
//     block's row maxima are known, i.e. before the next block's scores even exist; mref_ready barrier per 32 rows);
//     every warpgroup keeps its own partial row sum together with the reference it is relative to;
//   * a thread makes two passes over its 128 scores (tensor memory is read twice: 32 at a time for the maximum, 32 at a
//     time for the exponentials) so that it needs ~100 registers, not 200;
//   * P leaves in two halves through ONE 16 KB half tile per query tile: keys 0-63, then - when P V over them has been
//     read (pvl_done), which the exponentials of keys 64-127 hide - keys 64-127;
//   * 224 KB of shared memory: Q_A Q_B | K ring of 2 | V ring of 2 | P_A half, P_B half.
//   warp 0 TMA producer, warp 1 Q K^T issuer, warp 2 TMEM allocator (pipe observer of the trace build), warp 3 P V
//   issuer, warps 4-7 / 8-11 tile A even / odd blocks, warps 12-15 / 16-19 tile B.
#pragma once

constexpr int kThreads9 = 640;
constexpr int kSmemTiles9 = 7;                      // Q_A Q_B K0 K1 V0 V1 (P_A half, P_B half)
constexpr int kBarOff9 = kSmemTiles9 * kTileBytes;  // 256 B of barriers
constexpr int kXchOff9 = kBarOff9 + 256;            // mref[block parity][tile][128 rows] fp32
constexpr int kXchBytes9 = 2 * 2 * 128 * 4;
constexpr int kSmemPad9 = 768;                      // the dynamic window must be 256-byte aligned (checked)
constexpr int kSmemBytes9 = kXchOff9 + kXchBytes9 + kSmemPad9;
static_assert(kSmemBytes9 <= 232448, "227 KB of shared memory per CTA");

template <int kPolyMask8, bool kTrace = false, bool kOnePass = false>
__global__ void __launch_bounds__(kThreads9, 1) attn_tc_kernel_v9(const __grid_constant__ AttnParams P) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t raw_u32 = ptx::smem_u32(smem_raw);
  const uint32_t pad = ((raw_u32 + 1023u) & ~1023u) - raw_u32;
  if (pad > kSmemPad9) __trap();
  uint8_t* smem = smem_raw + pad;
  uint8_t* smem_q = smem;                     // 2 tiles
  uint8_t* smem_k = smem + 2 * kTileBytes;    // 2 tiles
  uint8_t* smem_v = smem + 4 * kTileBytes;    // 2 tiles
  uint8_t* smem_p = smem + 6 * kTileBytes;    // 2 half tiles: one per query tile, used by its two key halves in turn
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kBarOff9);
  uint64_t* q_full = bars;
  uint64_t* k_full = bars + 1;    // [2]
  uint64_t* k_empty = bars + 3;   // [2]
  uint64_t* v_full = bars + 5;    // [2]
  uint64_t* v_empty = bars + 7;   // [2]
  // A warpgroup visits every OTHER key block, and a parity wait cannot tell phase n - 1 from phase n + 1: every barrier
  // a softmax warp waits on exists once per block parity ([tile * 2 + (j & 1)], phase (j >> 1) & 1), so that each waiter
  // sees consecutive phases.
  uint64_t* s_full = bars + 9;    // [2][2] Q K^T of the tile's block j has completed
  uint64_t* s_free = bars + 13;   // [2] the block's softmax warps hold its scores in registers (4 arrivals)
  uint64_t* p_lo = bars + 15;     // [2] keys 0-63 of P written (4 arrivals)
  uint64_t* p_hi = bars + 17;     // [2] keys 64-127
  uint64_t* pvl_done = bars + 19;  // [2][2] P V over keys 0-63 of the tile's block has completed: the P half tile is free
  uint64_t* pv_done = bars + 23;   // [2][2] ... over keys 64-127 too: the P half tile is free, O may be rescaled
  uint64_t* o_full = bars + 27;
  uint64_t* mref_ready = bars + 28;  // [2][2] the block's reference exponents are in shared memory (4 arrivals)
  float* mref = reinterpret_cast<float*>(smem + kXchOff9);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mref);  // dead after the set-up barrier, long before mref is written

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qp = blockIdx.x % P.n_qpairs;
  const int bh = blockIdx.x / P.n_qpairs;
  const int h = bh % P.heads, b = bh / P.heads;
  const int q0 = qp * 2 * BQ;
  const int n_kv = (P.S + BKV - 1) / BKV;
  auto trace = [&](int j, int slot) {
    if constexpr (kTrace) {
      if (blockIdx.x == 0 && (threadIdx.x & 31) == 0 && j < 128) RT_ATTN_TRACE_STORE(j, slot);
    }
  };

  if (warp == 0 && lane == 0) ptx::prefetch_tmap(&P.tm);
  if (warp == 1 && lane == 0) {
    ptx::mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&k_full[i], 1);
      ptx::mbar_init(&k_empty[i], 1);
      ptx::mbar_init(&v_full[i], 1);
      ptx::mbar_init(&v_empty[i], 1);
      ptx::mbar_init(&s_free[i], 4);
      ptx::mbar_init(&p_lo[i], 4);
      ptx::mbar_init(&p_hi[i], 4);
    }
    for (int i = 0; i < 4; ++i) {
      ptx::mbar_init(&s_full[i], 1);
      ptx::mbar_init(&pvl_done[i], 1);
      ptx::mbar_init(&pv_done[i], 1);
      ptx::mbar_init(&mref_ready[i], 4);
    }
    ptx::mbar_init(o_full, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 2) ptx::tmem_alloc<1>(tmem_slot, 512);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  ptx::grid_launch_dependents();
  ptx::grid_dependency_wait();

  if (warp < 4) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
    if (kTrace && warp == 2) {
      for (int j = 0; j < n_kv; ++j) {
        if (j + 1 < n_kv) { ptx::mbar_wait(&s_full[(j + 1) & 1], ((j + 1) >> 1) & 1); trace(j, 22); }
        if (j + 1 < n_kv) { ptx::mbar_wait(&s_full[2 + ((j + 1) & 1)], ((j + 1) >> 1) & 1); trace(j, 25); }
        ptx::mbar_wait(&pv_done[j & 1], (j >> 1) & 1); trace(j, 21);
        ptx::mbar_wait(&pv_done[2 + (j & 1)], (j >> 1) & 1); trace(j, 24);
      }
    }
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
        const int st = j & 1, ph = (j >> 1) & 1;
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
      // ===================== Q K^T issuer =====================
      constexpr uint32_t idesc_qk = ptx::make_idesc_bf16(BQ, BKV, 0, 0);  // A, B K-major
      const uint64_t q_desc = ptx::make_smem_desc_sw128(ptx::smem_u32(smem_q), 0, 1024);
      const uint64_t k_desc = ptx::make_smem_desc_sw128(ptx::smem_u32(smem_k), 0, 1024);
      constexpr uint32_t kTile16 = kTileBytes >> 4, kSub16 = kSubBytes >> 4;
      auto issue_qk = [&](int t, int kst) {
        const uint64_t qa = q_desc + (uint64_t)(t * kTile16), ka = k_desc + (uint64_t)(kst * kTile16);
#pragma unroll
        for (int kk = 0; kk < HD / 16; ++kk) {
          const uint32_t off = (kk >> 2) * kSub16 + (kk & 3) * 2;
          ptx::mma_bf16_ss<1>(tmem + t * 128, qa + off, ka + off, idesc_qk, kk != 0 ? 1u : 0u);
        }
      };
      ptx::mbar_wait(q_full, 0);
      for (int j = 0; j < n_kv; ++j) {  // scores of block j: as soon as the softmax threads have block j-1 in registers
        const int st = j & 1, ph = (j >> 1) & 1;
        ptx::mbar_wait(&k_full[st], ph);
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          if (j > 0) ptx::mbar_wait(&s_free[t], (j - 1) & 1);
          trace(j, 30 + t);
          ptx::tc_fence_after();
          if (ptx::elect_one()) {
            issue_qk(t, st);
            ptx::mma_commit(&s_full[t * 2 + st]);
            if (t == 1) ptx::mma_commit(&k_empty[st]);
          }
          __syncwarp();
        }
      }
    } else if (warp == 3) {
      // ===================== P V issuer =====================
      constexpr uint32_t idesc_pv = ptx::make_idesc_bf16(BQ, HD, 0, 1);   // A (= P) K-major, B (= V) MN-major
      const uint64_t p_desc = ptx::make_smem_desc_sw128(ptx::smem_u32(smem_p), 0, 1024);
      const uint64_t v_desc = ptx::make_smem_desc_sw128(ptx::smem_u32(smem_v), kSubBytes, 1024);
      constexpr uint32_t kTile16 = kTileBytes >> 4, kSub16 = kSubBytes >> 4;
      auto issue_pv = [&](int t, int st, uint32_t acc, int kk0) {  // 64 keys: k-steps kk0 .. kk0 + 3
        const uint64_t pa = p_desc + (uint64_t)(t * kSub16), va = v_desc + (uint64_t)(st * kTile16);
#pragma unroll
        for (int kk = kk0; kk < kk0 + 4; ++kk)
          ptx::mma_bf16_ss<1>(tmem + 256 + t * 128, pa + (uint64_t)((kk & 3) * 2), va + (uint64_t)(kk * 128), idesc_pv,
                              kk != kk0 ? 1u : acc);
      };
      for (int j = 0; j < n_kv; ++j) {
        const int st = j & 1, ph = (j >> 1) & 1;
        ptx::mbar_wait(&v_full[st], ph);
        trace(j, 0);
#pragma unroll
        for (int t = 0; t < 2; ++t) {
          ptx::mbar_wait(&p_lo[t], j & 1);
          trace(j, 1 + t * 3);
          ptx::tc_fence_after();  // a (rare) rescale of O by the softmax threads precedes their arrival
          if (ptx::elect_one()) {
            issue_pv(t, st, j > 0 ? 1u : 0u, 0);
            ptx::mma_commit(&pvl_done[t * 2 + st]);
          }
          __syncwarp();
          ptx::mbar_wait(&p_hi[t], j & 1);
          trace(j, 2 + t * 3);
          ptx::tc_fence_after();
          if (ptx::elect_one()) {
            issue_pv(t, st, 1u, 4);
            ptx::mma_commit(&pv_done[t * 2 + st]);
            if (t == 1) ptx::mma_commit(&v_empty[st]);
          }
          __syncwarp();
        }
      }
      if (ptx::elect_one()) ptx::mma_commit(o_full);
      __syncwarp();
    }
  } else {
    // ===================== softmax: one thread per query row, two warpgroups per tile take the key blocks in turn ======
    asm volatile("setmaxnreg.inc.sync.aligned.u32 104;");
    const int idx = warp - 4;
    const int quad = idx & 3;         // TMEM lane quadrant (= warp % 4)
    const int par = (idx >> 2) & 1;   // this warpgroup takes blocks j = par, par + 2, ...
    const int t = idx >> 3;           // 0: tile A, 1: tile B
    const int row = quad * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t s_addr = tmem + lane_off + t * 128;
    const uint32_t o_addr = tmem + lane_off + 256 + t * 128;
    uint8_t* p_row = smem_p + t * kSubBytes + row * 128;  // the tile's P half tile (keys 0-63, then keys 64-127)
    const int sw = row & 7;
    uint64_t* mr_mine = &mref_ready[t * 2 + par];
    uint64_t* mr_other = &mref_ready[t * 2 + (par ^ 1)];
    uint64_t* s_full_mine = &s_full[t * 2 + par];
    uint64_t* pvl_mine = &pvl_done[t * 2 + par];
    uint64_t* pv_other = &pv_done[t * 2 + (par ^ 1)];   // of the blocks j - 1, j + 1, ... (the other warpgroup's)
    const float c = P.scale_log2;
    auto ph2 = [](int n) { return static_cast<uint32_t>((n >> 1) & 1); };  // phase of block n on its parity's barrier
    float m_ref = -INFINITY, l = 0.f;   // l is relative to m_ref
    if constexpr (kOnePass) {
      // ONE pass over the scores: block j is exponentiated against the reference block j - 1 left behind (known before
      // S(j) exists), 32 scores at a time; a chunk whose maximum exceeds the reference by more than 2^8 raises it first
      // (rare: everything accumulated so far - l, this block's partial sum, the not yet published half of P, O - is
      // brought to the new reference at a point where the tensor pipe is known to be idle on this tile).  The integer
      // reference makes P's mantissa independent of the reference, so this form gives the two-pass form's values.
      for (int j = par; j < n_kv; j += 2) {
        ptx::mbar_wait(s_full_mine, ph2(j));
        if (quad == 0) trace(j, 8 + t * 4);
        ptx::tc_fence_after();
        const int n_valid = P.S - j * BKV;
        if (j == 0) {   // the first block has no predecessor: its row maximum, one extra pass
          float mx = -INFINITY;
#pragma unroll 1
          for (int ch = 0; ch < 4; ++ch) {
            uint32_t sv[32];
            ptx::tmem_ld_32x32b_x32(s_addr + ch * 32, sv);
            ptx::tmem_ld_wait();
            if (n_valid < BKV) {
#pragma unroll
              for (int i = 0; i < 32; ++i)
                if (ch * 32 + i >= n_valid) sv[i] = 0xff800000u;
            }
            float a = -INFINITY, b2 = -INFINITY;
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              a = fmaxf(a, fmaxf(__uint_as_float(sv[i]), __uint_as_float(sv[i + 1])));
              b2 = fmaxf(b2, fmaxf(__uint_as_float(sv[i + 2]), __uint_as_float(sv[i + 3])));
            }
            mx = fmaxf(mx, fmaxf(a, b2));
          }
          m_ref = ceilf(mx * c);
        } else {
          ptx::mbar_wait(mr_other, ph2(j - 1));
          const float m_prev = mref[((j - 1) & 1) * 256 + t * 128 + row];
          l = (j == 1) ? 0.f : l * ptx::ex2_approx(m_ref - m_prev);
          m_ref = m_prev;
        }
        if (quad == 0) trace(j, 9 + t * 4);
        const float2 c2 = make_float2(c, c);
        float2 lsum[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
        uint32_t pa[16], pb[16];
        auto chunk = [&](auto ch_tag, uint32_t (&pk)[16], uint32_t (&pk_sib)[16]) {
          constexpr int ch = decltype(ch_tag)::value;
          uint32_t sv[32];
          ptx::tmem_ld_32x32b_x32(s_addr + ch * 32, sv);
          ptx::tmem_ld_wait();
          if (ch == 3) {  // the scores are consumed: the tensor pipe may overwrite S with the tile's next block
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&s_free[t]);
          }
          if (n_valid < BKV) {
#pragma unroll
            for (int i = 0; i < 32; ++i)
              if (ch * 32 + i >= n_valid) sv[i] = 0xff800000u;
          }
          float a = -INFINITY, b2 = -INFINITY;
#pragma unroll
          for (int i = 0; i < 32; i += 4) {
            a = fmaxf(a, fmaxf(__uint_as_float(sv[i]), __uint_as_float(sv[i + 1])));
            b2 = fmaxf(b2, fmaxf(__uint_as_float(sv[i + 2]), __uint_as_float(sv[i + 3])));
          }
          const float cm_s = fmaxf(a, b2) * c;
          if (__any_sync(0xffffffffu, cm_s > m_ref + 8.f)) {
            const float m_new = ceilf(fmaxf(m_ref, cm_s));   // = m_ref for the lanes that did not ask
            const float f = ptx::ex2_approx(m_ref - m_new);
            l *= f;
            lsum[0].x *= f; lsum[0].y *= f; lsum[1].x *= f; lsum[1].y *= f;
            if (ch & 1) {   // the sibling chunk of this half of P is still in registers: a power of two, exact
              const __nv_bfloat162 f2 = __float2bfloat162_rn(f);
#pragma unroll
              for (int i = 0; i < 16; ++i) {
                __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&pk_sib[i]);
                v = __hmul2(v, f2);
                pk_sib[i] = *reinterpret_cast<uint32_t*>(&v);
              }
            }
            // O: every P V issued so far on this tile must have completed (and none can be issued before this warp's
            // next arrival on p_lo / p_hi)
            if (ch < 2) {
              if (j > 0) ptx::mbar_wait(pv_other, ph2(j - 1));
            } else {
              ptx::mbar_wait(pvl_mine, ph2(j));
            }
            if (j > 0 || ch >= 2) {
              ptx::tc_fence_after();
#pragma unroll 1
              for (int oc = 0; oc < 8; ++oc) {
                uint32_t r[16];
                ptx::tmem_ld_32x32b_x16(o_addr + oc * 16, r);
                ptx::tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 16; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * f);
                ptx::tmem_st_32x32b_x16(o_addr + oc * 16, r);
              }
              ptx::tmem_st_wait();
              ptx::tc_fence_before();
            }
            m_ref = m_new;
          }
          if (ch == 3) {   // the block's reference is final: the other warpgroup's next block starts from it
            mref[(j & 1) * 256 + t * 128 + row] = m_ref;
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(mr_mine);
          }
          const float2 nm2 = make_float2(-m_ref, -m_ref);
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
            lsum[i & 1] = __fadd2_rn(lsum[i & 1], e);
            pk[i] = ptx::pack_bf16x2(e.x, e.y);
          }
        };
        auto store_chunk = [&](int ck0, const uint32_t (&pk)[16]) {  // 32 keys = four 16-byte chunks, swizzled by the row
#pragma unroll
          for (int q = 0; q < 4; ++q)
            *reinterpret_cast<uint4*>(p_row + (((ck0 + q) ^ sw) << 4)) =
                make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
        };
        chunk(std::integral_constant<int, 0>{}, pa, pb);
        chunk(std::integral_constant<int, 1>{}, pb, pa);
        if (j > 0) ptx::mbar_wait(pv_other, ph2(j - 1));  // the P half tile: P V over keys 64-127 of block j - 1 has read it
        store_chunk(0, pa);
        store_chunk(4, pb);
        ptx::fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&p_lo[t]);
        if (quad == 0) trace(j, 10 + t * 4);
        chunk(std::integral_constant<int, 2>{}, pa, pb);
        chunk(std::integral_constant<int, 3>{}, pb, pa);
        ptx::mbar_wait(pvl_mine, ph2(j));   // P V over keys 0-63 of THIS block has read the half tile
        store_chunk(0, pa);
        store_chunk(4, pb);
        ptx::fence_proxy_async_smem();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&p_hi[t]);
        if (quad == 0) trace(j, 11 + t * 4);
        l += (lsum[0].x + lsum[0].y) + (lsum[1].x + lsum[1].y);
      }
    } else {
    for (int j = par; j < n_kv; j += 2) {
      ptx::mbar_wait(s_full_mine, ph2(j));
      if (quad == 0) trace(j, 8 + t * 4);
      ptx::tc_fence_after();
      const int n_valid = P.S - j * BKV;  // < 128 only on the last block
      // ---- pass 1: the row maximum, 32 scores at a time
      float mx = -INFINITY;
#pragma unroll 1
      for (int ch = 0; ch < 4; ++ch) {
        uint32_t sv[32];
        ptx::tmem_ld_32x32b_x32(s_addr + ch * 32, sv);
        ptx::tmem_ld_wait();
        if (n_valid < BKV) {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (ch * 32 + i >= n_valid) sv[i] = 0xff800000u;  // -inf
        }
        float a = -INFINITY, b2 = -INFINITY;
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          a = fmaxf(a, fmaxf(__uint_as_float(sv[i]), __uint_as_float(sv[i + 1])));
          b2 = fmaxf(b2, fmaxf(__uint_as_float(sv[i + 2]), __uint_as_float(sv[i + 3])));
        }
        mx = fmaxf(mx, fmaxf(a, b2));
      }
      const float mx_s = mx * c;
      // ---- the reference exponent of this block: the previous block's (from the other warpgroup), raised lazily
      if (j == 0) {
        m_ref = ceilf(mx_s);
      } else {
        ptx::mbar_wait(mr_other, ph2(j - 1));
        const float m_prev = mref[((j - 1) & 1) * 256 + t * 128 + row];
        l *= ptx::ex2_approx(m_ref - m_prev);  // this warpgroup's partial sum follows the reference (x 1 when unchanged)
        if (j == 1) l = 0.f;                    // first block of the odd warpgroup: m_ref was -inf
        m_ref = m_prev;
        if (__any_sync(0xffffffffu, mx_s > m_ref + 8.f)) {
          const float m_new = ceilf(fmaxf(m_ref, mx_s));
          const float f = ptx::ex2_approx(m_ref - m_new);
          l *= f;
          ptx::mbar_wait(pv_other, ph2(j - 1));  // O += P V of every earlier block has completed
          ptx::tc_fence_after();
#pragma unroll 1
          for (int ch = 0; ch < 8; ++ch) {
            uint32_t r[16];
            ptx::tmem_ld_32x32b_x16(o_addr + ch * 16, r);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * f);
            ptx::tmem_st_32x32b_x16(o_addr + ch * 16, r);
          }
          ptx::tmem_st_wait();
          ptx::tc_fence_before();
          m_ref = m_new;
        }
      }
      mref[(j & 1) * 256 + t * 128 + row] = m_ref;
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(mr_mine);
      if (quad == 0) trace(j, 9 + t * 4);
      // ---- pass 2: P = 2^(S c - m), 32 scores at a time; keys 0-63 leave first
      const float2 c2 = make_float2(c, c), nm2 = make_float2(-m_ref, -m_ref);
      float2 lsum[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
      auto exp_chunk = [&](int ch, uint32_t (&pk)[16]) {
        uint32_t sv[32];
        ptx::tmem_ld_32x32b_x32(s_addr + ch * 32, sv);
        ptx::tmem_ld_wait();
        if (ch == 3) {  // the scores are consumed: the tensor pipe may overwrite S with the tile's next block
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(&s_free[t]);
        }
        if (n_valid < BKV) {
#pragma unroll
          for (int i = 0; i < 32; ++i)
            if (ch * 32 + i >= n_valid) sv[i] = 0xff800000u;
        }
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
          lsum[i & 1] = __fadd2_rn(lsum[i & 1], e);
          pk[i] = ptx::pack_bf16x2(e.x, e.y);
        }
      };
      auto store_chunk = [&](int ck0, const uint32_t (&pk)[16]) {  // 32 keys = four 16-byte chunks, swizzled by the row
#pragma unroll
        for (int q = 0; q < 4; ++q)
          *reinterpret_cast<uint4*>(p_row + (((ck0 + q) ^ sw) << 4)) =
              make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
      };
      uint32_t pa[16], pb[16];
      exp_chunk(0, pa);
      if (j > 0) ptx::mbar_wait(pv_other, ph2(j - 1));  // the P half tile: P V over keys 64-127 of the previous block has read it
      store_chunk(0, pa);
      exp_chunk(1, pa);
      store_chunk(4, pa);
      ptx::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&p_lo[t]);
      if (quad == 0) trace(j, 10 + t * 4);
      exp_chunk(2, pa);
      exp_chunk(3, pb);
      ptx::mbar_wait(pvl_mine, ph2(j));   // P V over keys 0-63 of THIS block has read the half tile
      store_chunk(0, pa);
      store_chunk(4, pb);
      ptx::fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&p_hi[t]);
      if (quad == 0) trace(j, 11 + t * 4);
      l += (lsum[0].x + lsum[0].y) + (lsum[1].x + lsum[1].y);
    }
    }
    // ---- epilogue: bring both partial sums to the LAST block's reference, add them, then O / l -> bf16 -> shared ->
    // global; the two warpgroups of a tile share one staging tile and split the 128 columns
    const int jl = n_kv - 1;
    if (par != (jl & 1)) {
      ptx::mbar_wait(mr_other, ph2(jl));
      const float m_last = mref[(jl & 1) * 256 + t * 128 + row];
      l = (l > 0.f) ? l * ptx::ex2_approx(m_ref - m_last) : 0.f;
    }
    ptx::mbar_wait(o_full, 0);
    ptx::tc_fence_after();
    const int hf = par;
    const int bar_id = 1 + t * 4 + quad;  // named barrier of the two warps that share these 32 rows
    auto pair_sync = [&]() { asm volatile("bar.sync %0, 64;" ::"r"(bar_id) : "memory"); };
    float* xsum = reinterpret_cast<float*>(smem_p);  // every MMA has completed: the P half tiles are dead
    xsum[(t * 2 + hf) * 128 + row] = l;
    pair_sync();
    const float inv = 1.f / (l + xsum[(t * 2 + (hf ^ 1)) * 128 + row]);
    constexpr int kPitch = HD * 2 + 16;
    uint8_t* stage = smem + (t * 4 + quad) * (32 * kPitch);  // Q / K / V tiles are dead (o_full)
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
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 2) ptx::tmem_dealloc<1>(tmem, 512);
}
