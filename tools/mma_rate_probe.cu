// Stand-alone probe: how many SM cycles does one tcgen05.mma (bf16, M = 128, K = 16) take in the streams the attention
// kernel issues?  One CTA per SM, one issuing thread, no softmax, no hand-offs - only the tensor pipe and its operand
// fetch.  Build + run:  nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -o tools/mma_rate_probe.bin
//                            tools/mma_rate_probe.cu && tools/mma_rate_probe.bin
//   mode 0  SS  128x128x16, A (Q tile) and B (K tile) K-major in shared memory, 8 k-steps per accumulator   (Q K^T)
//   mode 1  TS  128x128x16, A from TMEM, B (V tile) MN-major in shared memory                                (P V)
//   mode 2  the attention order: PV_A x8, QK_A x8, PV_B x8, QK_B x8
//   mode 3  SS  128x256x16 (the GEMM's instruction)
//   mode 4  SS  128x64x16
//   mode 5  mode 2 with 64 KB of bulk copies into shared memory per 32 MMAs (the K / V stream)
//   mode 6  mode 0 with the same bulk-copy stream
//   mode 7  TS  128x256x16 (hypothetical head_dim 256 / two V tiles)
//   mode 8  SS  128x128x16 with A == B tile (half the distinct bytes)
//   mode 9  attention order with P V in SS mode too (P from shared memory: 8 KB of operands per MMA for all 32)
//   mode 10 mode 9 + the bulk-copy stream + two warps writing 64 KB / 32 MMAs with st.shared.v4 (the P tiles)
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../reptext_b200/csrc/ptx_sm100.cuh"


constexpr int kTile = 128 * 128 * 2;  // 32 KB
constexpr int kSub = kTile / 2;

__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   ptx::smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(ptx::smem_u32(bar))
               : "memory");
}

__global__ void __launch_bounds__(128, 1) probe(int mode, int rounds, const uint8_t* src, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = ptx::smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw + 1023u) & ~1023u) - raw);
  uint8_t* q = smem;                 // 2 tiles
  uint8_t* k = smem + 2 * kTile;     // 2 tiles (also the 256-row B of mode 3)
  uint8_t* v = smem + 4 * kTile;     // 2 tiles: stream target in modes 5 / 6
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 6 * kTile);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bars + 8);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 6 * kTile / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + i * 7u;
  if (threadIdx.x == 0) {
    ptx::mbar_init(&bars[0], 1);
    ptx::mbar_init(&bars[1], 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc<1>(slot, 512);
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *slot;
  const bool stream = mode == 5 || mode == 6 || mode == 10;
  if (warp == 0) {
    constexpr uint32_t id_ss128 = ptx::make_idesc_bf16(128, 128, 0, 0);
    constexpr uint32_t id_ts128 = ptx::make_idesc_bf16(128, 128, 0, 1);
    constexpr uint32_t id_ss256 = ptx::make_idesc_bf16(128, 256, 0, 0);
    constexpr uint32_t id_ss64 = ptx::make_idesc_bf16(128, 64, 0, 0);
    constexpr uint32_t id_ts256 = ptx::make_idesc_bf16(128, 256, 0, 1);
    const uint64_t qd = ptx::make_smem_desc_sw128(ptx::smem_u32(q), 0, 1024);
    const uint64_t kd = ptx::make_smem_desc_sw128(ptx::smem_u32(k), 0, 1024);
    const uint64_t vd = ptx::make_smem_desc_sw128(ptx::smem_u32(k), kSub, 1024);  // V-style descriptor over the k area
    constexpr uint32_t kTile16 = kTile >> 4, kSub16 = kSub >> 4;
    auto qk = [&](int t, uint32_t idesc, uint64_t a_desc) {
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {
        const uint32_t off = (kk >> 2) * kSub16 + (kk & 3) * 2;
        ptx::mma_bf16_ss<1>(tmem + t * 128, a_desc + (uint64_t)(t * kTile16) + off, kd + off, idesc, kk != 0);
      }
    };
    auto pv_ss = [&](int t, uint32_t d_col) {  // A = a "P" tile in the q area (K-major), B = V-style descriptor
#pragma unroll
      for (int kk = 0; kk < 8; ++kk) {
        const uint32_t off = (kk >> 2) * kSub16 + (kk & 3) * 2;
        ptx::mma_bf16_ss<1>(tmem + d_col, qd + (uint64_t)(t * kTile16) + off, vd + (uint64_t)(kk * 128), id_ts128, 1u);
      }
    };
    auto pv = [&](int t, uint32_t idesc, uint32_t d_col) {
#pragma unroll
      for (int kk = 0; kk < 8; ++kk)
        ptx::mma_bf16_ts(tmem + d_col, tmem + t * 128 + kk * 8, vd + (uint64_t)(kk * 128), idesc, 1u);
    };
    __syncwarp();
    long long t0 = clock64();
    if (ptx::elect_one()) {
      for (int r = 0; r < rounds; ++r) {
        switch (mode) {
          case 0: case 6: qk(0, id_ss128, qd); qk(1, id_ss128, qd); qk(0, id_ss128, qd); qk(1, id_ss128, qd); break;
          case 1: pv(0, id_ts128, 256); pv(1, id_ts128, 384); pv(0, id_ts128, 256); pv(1, id_ts128, 384); break;
          case 2: case 5: pv(0, id_ts128, 256); qk(0, id_ss128, qd); pv(1, id_ts128, 384); qk(1, id_ss128, qd); break;
          case 3: qk(0, id_ss256, qd); qk(0, id_ss256, qd); break;                      // 16 MMAs of twice the work
          case 4: qk(0, id_ss64, qd); qk(1, id_ss64, qd); qk(0, id_ss64, qd); qk(1, id_ss64, qd); break;
          case 7: pv(0, id_ts256, 256); pv(1, id_ts256, 256); break;                    // 16 MMAs of twice the work
          case 9: case 10: pv_ss(0, 256); qk(0, id_ss128, qd); pv_ss(1, 384); qk(1, id_ss128, qd); break;
          case 8: qk(0, id_ss128, kd - (uint64_t)0); qk(0, id_ss128, kd); qk(0, id_ss128, kd); qk(0, id_ss128, kd); break;
        }
      }
      ptx::mma_commit(&bars[0]);
    }
    __syncwarp();
    ptx::mbar_wait(&bars[0], 0);
    long long t1 = clock64();
    if (lane == 0) out[blockIdx.x] = t1 - t0;
  } else if (warp == 3 && mode == 10) {
    // the P tiles: 128 threads would write 64 KB per block; here one warp writes 16 B per lane, conflict-free rows,
    // paced by nothing - an upper bound on the LSU traffic the softmax warps can add
    uint4 val = make_uint4(lane, lane, lane, lane);
    for (int r = 0; r < rounds * 128; ++r) {  // 128 x 512 B = 64 KB per round
      *reinterpret_cast<uint4*>(v + kTile + ((r & 63) * 512) + lane * 16) = val;
      val.x += r;
    }
  } else if (warp == 2 && lane == 0 && stream) {
    // the K / V stream: 4 x 16 KB per 32 MMAs (~2048 cycles); not synchronised with the MMAs (timing only)
    for (int r = 0; r < rounds; ++r) {
      ptx::mbar_arrive_expect_tx(&bars[1], 4 * kSub);
      for (int c = 0; c < 4; ++c) bulk_g2s(v + c * kSub, src + ((size_t)blockIdx.x * 4 + c) * kSub, kSub, &bars[1]);
      ptx::mbar_wait(&bars[1], r & 1);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc<1>(tmem, 512);
}

int main(int argc, char** argv) {
  const int rounds = argc > 1 ? atoi(argv[1]) : 256;
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int smem = 6 * kTile + 256 + 1024;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  uint8_t* src;
  long long* out;
  cudaMalloc(&src, (size_t)sms * 4 * kSub);
  cudaMemset(src, 0x3c, (size_t)sms * 4 * kSub);
  cudaMalloc(&out, sms * sizeof(long long));
  const char* names[] = {"SS 128x128x16 (Q K^T)", "TS 128x128x16 (P V)", "attention order PV8 QK8 PV8 QK8", "SS 128x256x16",
                         "SS 128x64x16", "attention order + 64 KB smem stream / 32 MMAs", "SS 128x128x16 + stream",
                         "TS 128x256x16", "SS 128x128x16, A == B tile", "attention order, P V in SS mode",
                         "SS attention order + TMA stream + st.shared P stream"};
  for (int grid : {1, sms}) {
    for (int mode = 0; mode <= 10; ++mode) {
      for (int rep = 0; rep < 2; ++rep) probe<<<grid, 128, smem>>>(mode, rounds, src, out);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) {
        printf("mode %d: %s\n", mode, cudaGetErrorString(e));
        return 1;
      }
      std::vector<long long> h(grid);
      cudaMemcpy(h.data(), out, grid * sizeof(long long), cudaMemcpyDeviceToHost);
      double sum = 0;
      for (long long x : h) sum += (double)x;
      const int per_round = (mode == 3 || mode == 7) ? 16 : 32;
      const double cyc = sum / grid / ((double)rounds * per_round);
      const double work = (mode == 3 || mode == 7) ? 2.0 : (mode == 4 ? 0.5 : 1.0);  // in units of 128x128x16
      printf("grid %3d  mode %d  %-48s %7.1f cycles / MMA   %6.1f cycles per 128x128x16 of work (nominal 64)\n", grid, mode,
             names[mode], cyc, cyc / work);
    }
  }
  return 0;
}
