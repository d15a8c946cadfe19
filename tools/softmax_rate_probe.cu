// Stand-alone probe: instruction throughput of the attention kernel's softmax inner loop on one SM sub-partition (SMSP),
// without any tensor-core work.  Each thread owns a row of 128 scores in registers (as in attn_sm100.cu) and runs
//     x = s * c - m;  e = 2^x (MUFU.EX2 or the cubic on the FMA pipe);  sum += e;  pack to bf16
// `rounds` times; the result is SM cycles per 32-lane element step per SMSP for 1, 2 and 4 resident warps per SMSP.
//   nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -o tools/softmax_rate_probe.bin tools/softmax_rate_probe.cu
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../reptext_b200/csrc/ptx_sm100.cuh"

__device__ __forceinline__ float2 exp2_poly2(float2 x) {
  x.x = fmaxf(x.x, -125.f);
  x.y = fmaxf(x.y, -125.f);
  const float2 magic = make_float2(12582912.f, 12582912.f);
  const float2 xi = __fadd2_rn(x, magic);
  const float2 n = __fadd2_rn(xi, make_float2(-12582912.f, -12582912.f));
  const float2 f = __ffma2_rn(n, make_float2(-1.f, -1.f), x);
  float2 p = __ffma2_rn(f, make_float2(0.0550440177f, 0.0550440177f), make_float2(0.24229379f, 0.24229379f));
  p = __ffma2_rn(p, f, make_float2(0.69325459f, 0.69325459f));
  p = __ffma2_rn(p, f, make_float2(0.99994999f, 0.99994999f));
  return make_float2(__int_as_float(__float_as_int(p.x) + (__float_as_int(xi.x) << 23)),
                     __int_as_float(__float_as_int(p.y) + (__float_as_int(xi.y) << 23)));
}

__device__ __forceinline__ uint32_t ex2_f16x2(uint32_t x) {
  uint32_t y;
  asm volatile("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x));
  return y;
}
__device__ __forceinline__ uint32_t ex2_bf16x2(uint32_t x) {
  uint32_t y;
  asm volatile("ex2.approx.ftz.bf16x2 %0, %1;" : "=r"(y) : "r"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
  uint32_t y;
  asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(y) : "f"(hi), "f"(lo));
  return y;
}
__device__ __forceinline__ uint32_t hadd2(uint32_t a, uint32_t b) {
  uint32_t y;
  asm volatile("add.rn.f16x2 %0, %1, %2;" : "=r"(y) : "r"(a), "r"(b));
  return y;
}
__device__ __forceinline__ float2 unpack_f16x2(uint32_t v) {
  float lo, hi;
  asm volatile("{.reg .b16 l, h; mov.b32 {l, h}, %2; cvt.f32.f16 %0, l; cvt.f32.f16 %1, h;}" : "=f"(lo), "=f"(hi) : "r"(v));
  return make_float2(lo, hi);
}

// kMask8: bit (i % 8) set = pair i uses the polynomial.  kSum / kPack: keep the row sum / the bf16 packing.
// kMode: 0 = the softmax step above; 1 = MUFU.EX2 only; 2 = F2FP only; 3 = FFMA2 only; 4 = scalar FFMA only
template <int kMask8, bool kSum, bool kPack, int kMode>
__global__ void __launch_bounds__(512, 1) probe(const float* in, int rounds, float* out, long long* cyc) {
  float s[128];
#pragma unroll
  for (int i = 0; i < 128; ++i) s[i] = in[(threadIdx.x * 128 + i) & 4095];
  float c = in[4096], m = in[4097];
  float2 lsum = make_float2(0.f, 0.f);
  uint32_t keep = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int r = 0; r < rounds; ++r) {
    if constexpr (kMode == 0) {
      const float2 c2 = make_float2(c, c), nm2 = make_float2(-m, -m);
#pragma unroll
      for (int i = 0; i < 64; ++i) {
        float2 x = __ffma2_rn(make_float2(s[2 * i], s[2 * i + 1]), c2, nm2);
        float2 e;
        if ((kMask8 >> (i & 7)) & 1) {
          e = exp2_poly2(x);
        } else {
          e.x = ptx::ex2_approx(x.x);
          e.y = ptx::ex2_approx(x.y);
        }
        if constexpr (kSum) lsum = __fadd2_rn(lsum, e);
        if constexpr (kPack) keep ^= ptx::pack_bf16x2(e.x, e.y);
        else keep ^= __float_as_uint(e.x) ^ __float_as_uint(e.y);
      }
      m += 1e-6f;  // a new reference every round: nothing can be hoisted
    } else if constexpr (kMode == 5 || kMode == 6 || kMode == 7) {
      // packed half-precision exponential: x (fp32) -> f16x2 -> MUFU.EX2 on the pair -> the packed P itself;
      // mode 5: f16x2 with a 3-level HADD2 tree + fp32 widening for the row sum; 6: f16x2, no sum; 7: bf16x2, no sum
      const float2 c2 = make_float2(c, c), nm2 = make_float2(-m, -m);
#pragma unroll
      for (int g = 0; g < 8; ++g) {
        uint32_t pk[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float2 x = __ffma2_rn(make_float2(s[16 * g + 2 * i], s[16 * g + 2 * i + 1]), c2, nm2);
          if constexpr (kMode == 7) pk[i] = ex2_bf16x2(ptx::pack_bf16x2(x.x, x.y));
          else pk[i] = ex2_f16x2(pack_f16x2(x.x, x.y));
          keep ^= pk[i];
        }
        if constexpr (kMode == 5) {
          const uint32_t t = hadd2(hadd2(hadd2(pk[0], pk[1]), hadd2(pk[2], pk[3])), hadd2(hadd2(pk[4], pk[5]), hadd2(pk[6], pk[7])));
          lsum = __fadd2_rn(lsum, unpack_f16x2(t));
        }
      }
      m += 1e-6f;
    } else if constexpr (kMode == 8) {
#pragma unroll
      for (int i = 0; i < 128; ++i) s[i] = __uint_as_float(ex2_f16x2(__float_as_uint(s[i])));
    } else if constexpr (kMode == 1) {
#pragma unroll
      for (int i = 0; i < 128; ++i) s[i] = ptx::ex2_approx(s[i]);
    } else if constexpr (kMode == 2) {
#pragma unroll
      for (int i = 0; i < 64; ++i) keep ^= ptx::pack_bf16x2(s[2 * i] + m, s[2 * i + 1]);
      m += 1.f;
    } else if constexpr (kMode == 3) {
      const float2 c2 = make_float2(c, c), nm2 = make_float2(-m, -m);
#pragma unroll
      for (int i = 0; i < 64; ++i) {
        float2 x = __ffma2_rn(make_float2(s[2 * i], s[2 * i + 1]), c2, nm2);
        s[2 * i] = x.x;
        s[2 * i + 1] = x.y;
      }
    } else {
#pragma unroll
      for (int i = 0; i < 128; ++i) s[i] = fmaf(s[i], c, -m);
    }
  }
  const long long t1 = clock64();
  float acc = lsum.x + lsum.y + __uint_as_float(keep & 0x007fffffu);
#pragma unroll
  for (int i = 0; i < 128; ++i) acc += s[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}


// Straight-line form over registers (the product's) with CHAINED references: the scale-subtract of chunk k takes its
// reference through fma(e, 0, -m) from an exponential of chunk k - kLag, which equals -m but is only available once that
// chunk's exponentials have been issued - the scheduler can no longer hoist every FFMA2 to the front and leave the
// MUFU-bound tail on its own.  Same values.
template <int kMask8, int kCh, int kLag>
__global__ void __launch_bounds__(512, 1) probe_chained(const float* in, int rounds, float* out, long long* cyc) {
  float s[128];
#pragma unroll
  for (int i = 0; i < 128; ++i) s[i] = in[(threadIdx.x * 128 + i) & 4095];
  float c = in[4096], m = in[4097];
  float2 lsum0 = make_float2(0.f, 0.f), lsum1 = make_float2(0.f, 0.f);
  uint32_t keep = 0;
  __syncthreads();
  const long long t0 = clock64();
  for (int r = 0; r < rounds; ++r) {
    const float2 c2 = make_float2(c, c);
    constexpr int kNC = 128 / kCh, kP = kCh / 2;
    float tail[kNC];   // one exponential of every chunk
#pragma unroll
    for (int k = 0; k < kNC; ++k) {
      float nm = -m;
      if (k >= kLag) nm = fmaf(tail[k - kLag], 0.f, nm);
      const float2 nm2 = make_float2(nm, nm);
#pragma unroll
      for (int i = 0; i < kP; ++i) {
        const int p = k * kP + i;
        float2 x = __ffma2_rn(make_float2(s[2 * p], s[2 * p + 1]), c2, nm2);
        float2 e;
        if ((kMask8 >> (i & 7)) & 1) {
          e = exp2_poly2(x);
        } else {
          e.x = ptx::ex2_approx(x.x);
          e.y = ptx::ex2_approx(x.y);
        }
        if (i & 1) lsum1 = __fadd2_rn(lsum1, e);
        else lsum0 = __fadd2_rn(lsum0, e);
        keep ^= ptx::pack_bf16x2(e.x, e.y);
        if (i == kP - 1) tail[k] = e.y;
      }
    }
    m += 1e-6f;
  }
  const long long t1 = clock64();
  float acc = lsum0.x + lsum0.y + lsum1.x + lsum1.y + __uint_as_float(keep & 0x007fffffu);
#pragma unroll
  for (int i = 0; i < 128; ++i) acc += s[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int kMask8, int kCh, int kLag>
void run_chained(const char* name, const float* in, float* out, long long* cyc, int rounds) {
  for (int threads : {128, 256, 512}) {
    for (int rep = 0; rep < 2; ++rep) probe_chained<kMask8, kCh, kLag><<<1, threads>>>(in, rounds, out, cyc);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("%s: %s\n", name, cudaGetErrorString(e));
      exit(1);
    }
    long long h;
    cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    const int warps_per_smsp = threads / 128;
    printf("%-44s %d warp(s) / SMSP: %6.2f cycles per 32-lane element step per SMSP  (%7.0f cycles per 128-key row block per warp)\n",
           name, warps_per_smsp, (double)h / ((double)rounds * 128 * warps_per_smsp), (double)h / rounds);
  }
}

// Software-pipelined form of the same step, the way a kernel that reads S from tensor memory and writes P back can run
// it: a ROLLED loop over chunks of kCh keys (shared memory stands in for tensor memory: dynamic addresses, static
// registers).  Iteration c issues the exponentials of chunk c, the scale-subtract of chunk c + 1 and the row sum / bf16
// packing / store of chunk c - 1: every consumer sits one whole iteration behind its producer, so one warp alone can keep
// the MUFU pipe and the FMA pipe busy at the same time (the straight-line form above leaves that to the scheduler's
// luck: 15-18 cycles per key with one warp).
template <int kMask8, int kCh>
__global__ void __launch_bounds__(512, 1) probe_pipelined(const float* in, int rounds, int zero, float* out, long long* cyc) {
  extern __shared__ float4 sm[];
  constexpr int kP = kCh / 2;                           // pairs per chunk
  constexpr int kNC = 128 / kCh;                        // chunks per 128-key block
  float4* s_in = sm;                                    // [kCh / 4][blockDim]
  uint4* s_out = reinterpret_cast<uint4*>(sm + (kCh / 4) * blockDim.x);  // [kCh / 8][blockDim]
  for (int k = 0; k < kCh / 4; ++k)
    s_in[k * blockDim.x + threadIdx.x] = make_float4(in[(threadIdx.x * 4 + k) & 4095], in[(threadIdx.x * 4 + k + 1) & 4095],
                                                     in[(threadIdx.x * 4 + k + 2) & 4095], in[(threadIdx.x * 4 + k + 3) & 4095]);
  float c = in[4096], m = in[4097];
  float2 lsum0 = make_float2(0.f, 0.f), lsum1 = make_float2(0.f, 0.f);
  __syncthreads();
  auto load = [&](float2 (&raw)[kP], int ch) {
    const float4* src = s_in + (ch * zero) * blockDim.x + threadIdx.x;
#pragma unroll
    for (int k = 0; k < kCh / 4; ++k) {
      const float4 v = src[k * blockDim.x];
      raw[2 * k] = make_float2(v.x, v.y);
      raw[2 * k + 1] = make_float2(v.z, v.w);
    }
  };
  const long long t0 = clock64();
  for (int r = 0; r < rounds; ++r) {
    const float2 c2 = make_float2(c, c), nm2 = make_float2(-m, -m);
    // two register sets (A / B) alternate roles, so that the rolled loop needs no register moves:
    //   half-iteration on (cur, nxt): exponentials of cur.x -> cur.e, scale-subtract nxt.raw -> nxt.x, then load the
    //   chunk after next into cur.raw... each pair's three stages are INTERLEAVED in the source (one basic block).
    float2 rawA[kP], xA[kP], eA[kP], rawB[kP], xB[kP], eB[kP];
    load(rawA, 0);
#pragma unroll
    for (int i = 0; i < kP; ++i) xA[i] = __ffma2_rn(rawA[i], c2, nm2);
    load(rawB, 1);
#pragma unroll
    for (int i = 0; i < kP; ++i) eB[i] = make_float2(0.f, 0.f);
    auto half = [&](float2 (&xc)[kP], float2 (&ec)[kP], float2 (&rawn)[kP], float2 (&xn)[kP], float2 (&ep)[kP], int ch) {
      // exps of chunk ch (xc -> ec) | scale-subtract of chunk ch + 1 (rawn -> xn) | sum / pack / store of chunk ch - 1 (ep)
      uint32_t pk[kP];
#pragma unroll
      for (int i = 0; i < kP; ++i) {
        const bool poly = (kMask8 >> (i & 7)) & 1;
        if (!poly) ec[i].x = ptx::ex2_approx(xc[i].x);
        xn[i] = __ffma2_rn(rawn[i], c2, nm2);
        if (i & 1) lsum1 = __fadd2_rn(lsum1, ep[i]);
        else lsum0 = __fadd2_rn(lsum0, ep[i]);
        if (!poly) ec[i].y = ptx::ex2_approx(xc[i].y);
        else ec[i] = exp2_poly2(xc[i]);
        pk[i] = ptx::pack_bf16x2(ep[i].x, ep[i].y);
      }
      uint4* dst = s_out + (ch * zero) * blockDim.x + threadIdx.x;
#pragma unroll
      for (int k = 0; k < kP / 4; ++k) dst[k * blockDim.x] = make_uint4(pk[4 * k], pk[4 * k + 1], pk[4 * k + 2], pk[4 * k + 3]);
    };
#pragma unroll 1
    for (int ch = 0; ch < kNC; ch += 2) {
      half(xA, eA, rawB, xB, eB, ch);       // (the first pass sums / stores zeros: the pipeline's fill)
      load(rawA, ch + 2);
      half(xB, eB, rawA, xA, eA, ch + 1);
      load(rawB, ch + 3);
    }
#pragma unroll
    for (int i = 0; i < kP; ++i) {          // drain: the last chunk's sum
      if (i & 1) lsum1 = __fadd2_rn(lsum1, eB[i]);
      else lsum0 = __fadd2_rn(lsum0, eB[i]);
    }
    m += 1e-6f;
  }
  const long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = lsum0.x + lsum0.y + lsum1.x + lsum1.y + __uint_as_float(s_out[threadIdx.x].x & 0x007fffffu);
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int kMask8, int kCh>
void run_pipelined(const char* name, const float* in, float* out, long long* cyc, int rounds) {
  cudaFuncSetAttribute(probe_pipelined<kMask8, kCh>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
  for (int threads : {128, 256, 512}) {
    const size_t smem = (size_t)(kCh / 4 + kCh / 8) * threads * 16;
    for (int rep = 0; rep < 2; ++rep) probe_pipelined<kMask8, kCh><<<1, threads, smem>>>(in, rounds, 0, out, cyc);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("%s: %s\n", name, cudaGetErrorString(e));
      exit(1);
    }
    long long h;
    cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    const int warps_per_smsp = threads / 128;
    printf("%-44s %d warp(s) / SMSP: %6.2f cycles per 32-lane element step per SMSP  (%7.0f cycles per 128-key row block per warp)\n",
           name, warps_per_smsp, (double)h / ((double)rounds * 128 * warps_per_smsp), (double)h / rounds);
  }
}

template <int kMask8, bool kSum, bool kPack, int kMode>
void run(const char* name, const float* in, float* out, long long* cyc, int rounds) {
  for (int threads : {128, 256, 512}) {
    for (int rep = 0; rep < 2; ++rep) probe<kMask8, kSum, kPack, kMode><<<1, threads>>>(in, rounds, out, cyc);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
      printf("%s: %s\n", name, cudaGetErrorString(e));
      exit(1);
    }
    long long h;
    cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    const int warps_per_smsp = threads / 128;
    // per SMSP: warps_per_smsp warps x 128 elements x rounds element steps of 32 lanes
    printf("%-44s %d warp(s) / SMSP: %6.2f cycles per 32-lane element step per SMSP  (%7.0f cycles per 128-key row block per warp)\n",
           name, warps_per_smsp, (double)h / ((double)rounds * 128 * warps_per_smsp), (double)h / rounds);
  }
}

int main(int argc, char** argv) {
  const int rounds = argc > 1 ? atoi(argv[1]) : 2000;
  float* in;
  float* out;
  long long* cyc;
  cudaMalloc(&in, 4100 * sizeof(float));
  std::vector<float> h(4100);
  for (int i = 0; i < 4096; ++i) h[i] = -1.f - (float)(i % 97) * 0.01f;
  h[4096] = 0.127f;
  h[4097] = 0.5f;
  cudaMemcpy(in, h.data(), 4100 * sizeof(float), cudaMemcpyHostToDevice);
  cudaMalloc(&out, 512 * sizeof(float));
  cudaMalloc(&cyc, sizeof(long long));
  run_chained<0x88, 16, 1>("CHAINED 16-key chunks lag 1, 25 % poly", in, out, cyc, rounds);
  run_chained<0x88, 16, 2>("CHAINED 16-key chunks lag 2, 25 % poly", in, out, cyc, rounds);
  run_chained<0x88, 8, 2>("CHAINED 8-key chunks lag 2, 25 % poly", in, out, cyc, rounds);
  run_chained<0x88, 8, 3>("CHAINED 8-key chunks lag 3, 25 % poly", in, out, cyc, rounds);
  run_chained<0x88, 32, 1>("CHAINED 32-key chunks lag 1, 25 % poly", in, out, cyc, rounds);
  run_chained<0x00, 16, 2>("CHAINED 16-key chunks lag 2, MUFU only", in, out, cyc, rounds);
  run_chained<0xAA, 16, 2>("CHAINED 16-key chunks lag 2, 50 % poly", in, out, cyc, rounds);
  run_pipelined<0x00, 32>("PIPELINED 32-key chunks, MUFU only", in, out, cyc, rounds);
  run_pipelined<0x88, 32>("PIPELINED 32-key chunks, 25 % polynomial", in, out, cyc, rounds);
  run_pipelined<0x92, 32>("PIPELINED 32-key chunks, 37.5 % polynomial", in, out, cyc, rounds);
  run_pipelined<0xAA, 32>("PIPELINED 32-key chunks, 50 % polynomial", in, out, cyc, rounds);
  run_pipelined<0x00, 16>("PIPELINED 16-key chunks, MUFU only", in, out, cyc, rounds);
  run_pipelined<0x88, 16>("PIPELINED 16-key chunks, 25 % polynomial", in, out, cyc, rounds);
  run_pipelined<0xAA, 16>("PIPELINED 16-key chunks, 50 % polynomial", in, out, cyc, rounds);
  run<0x00, true, true, 0>("softmax step, MUFU only", in, out, cyc, rounds);
  run<0x88, true, true, 0>("softmax step, 25 % polynomial (product)", in, out, cyc, rounds);
  run<0xAA, true, true, 0>("softmax step, 50 % polynomial", in, out, cyc, rounds);
  run<0xFF, true, true, 0>("softmax step, polynomial only", in, out, cyc, rounds);
  run<0x00, false, true, 0>("MUFU only, no row sum", in, out, cyc, rounds);
  run<0x00, true, false, 0>("MUFU only, no bf16 pack", in, out, cyc, rounds);
  run<0x00, false, false, 0>("MUFU only, no sum, no pack", in, out, cyc, rounds);
  run<0, false, false, 5>("f16x2 step: FFMA2, cvt, EX2.F16x2, HADD2-tree sum", in, out, cyc, rounds);
  run<0, false, false, 6>("f16x2 step, no row sum", in, out, cyc, rounds);
  run<0, false, false, 7>("bf16x2 step, no row sum", in, out, cyc, rounds);
  run<0, false, false, 8>("MUFU.EX2.F16x2 alone (per PAIR)", in, out, cyc, rounds);
  run<0, false, false, 1>("MUFU.EX2 alone", in, out, cyc, rounds);
  run<0, false, false, 2>("F2FP.BF16 pack alone (+1 FADD)", in, out, cyc, rounds);
  run<0, false, false, 3>("FFMA2 alone", in, out, cyc, rounds);
  run<0, false, false, 4>("FFMA (scalar) alone", in, out, cyc, rounds);
  return 0;
}
