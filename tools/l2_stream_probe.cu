// Stand-alone probe: how fast can every SM stream K / V-sized tiles from L2 into shared memory?  One CTA per SM issues
// cp.async.bulk copies of 16 KB (4 in flight) over a region that all CTAs share (`span` bytes, L2-resident), the way 148
// attention CTAs re-read the same head's K / V.  Prints bytes per SM clock per SM and the aggregate TB/s.
//   nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -o tools/l2_stream_probe.bin tools/l2_stream_probe.cu
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../reptext_b200/csrc/ptx_sm100.cuh"

constexpr int kChunk = 16 * 1024;
constexpr int kDepth = 8;

__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   ptx::smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(ptx::smem_u32(bar))
               : "memory");
}

__global__ void __launch_bounds__(32, 1) probe(const uint8_t* src, size_t span, int chunks, int stride_ctas, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kDepth * kChunk);
  if (threadIdx.x == 0) {
    for (int i = 0; i < kDepth; ++i) ptx::mbar_init(&bars[i], 1);
    ptx::fence_barrier_init();
  }
  __syncwarp();
  if (threadIdx.x == 0) {
    // CTAs that share a "head" read the same bytes (stride_ctas CTAs per distinct region start)
    size_t off = ((size_t)(blockIdx.x / stride_ctas) * 2654435761u) % (span / kChunk) * kChunk;
    const long long t0 = clock64();
    for (int c = 0; c < chunks + kDepth; ++c) {
      const int st = c % kDepth;
      if (c >= kDepth) ptx::mbar_wait(&bars[st], ((c / kDepth) - 1) & 1);
      if (c < chunks) {
        ptx::mbar_arrive_expect_tx(&bars[st], kChunk);
        bulk_g2s(smem + st * kChunk, src + off, kChunk, &bars[st]);
        off += kChunk;
        if (off >= span) off = 0;
      }
    }
    out[blockIdx.x] = clock64() - t0;
  }
}

int main(int argc, char** argv) {
  int dev = 0, sms = 0, khz = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
  const int smem = kDepth * kChunk + 256;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const size_t span_max = 256ull << 20;
  uint8_t* src;
  long long* out;
  cudaMalloc(&src, span_max);
  cudaMemset(src, 1, span_max);
  cudaMalloc(&out, sms * sizeof(long long));
  const int chunks = 4096;  // 64 MB per CTA
  printf("# %d SMs, %d x 16 KB bulk copies in flight per SM, %d chunks per CTA\n", sms, kDepth, chunks);
  for (size_t span : {size_t(2) << 20, size_t(8) << 20, size_t(32) << 20, size_t(96) << 20, size_t(256) << 20}) {
    for (int share : {148, 6, 1}) {  // CTAs reading the same stream: all / one head's worth / none
      for (int rep = 0; rep < 2; ++rep) probe<<<sms, 32, smem>>>(src, span, chunks, share, out);
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0); cudaEventCreate(&e1);
      cudaEventRecord(e0);
      probe<<<sms, 32, smem>>>(src, span, chunks, share, out);
      cudaEventRecord(e1);
      if (cudaDeviceSynchronize() != cudaSuccess) { printf("error\n"); return 1; }
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      std::vector<long long> h(sms);
      cudaMemcpy(h.data(), out, sms * sizeof(long long), cudaMemcpyDeviceToHost);
      double cyc = 0; for (auto x : h) cyc += (double)x; cyc /= sms;
      const double bytes = (double)chunks * kChunk;
      printf("span %4zu MB, %3d CTAs per stream: %6.1f B/clk/SM, aggregate %6.2f TB/s (event time %.3f ms)\n", span >> 20, share,
             bytes / cyc, bytes * sms / (ms * 1e-3) / 1e12, ms);
    }
  }
  return 0;
}
