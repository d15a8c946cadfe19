// HBM-bound kernels of the VAE path (diffusers AutoencoderKL as the RepText pipelines call it,
// RepText/pipeline_flux_controlnet.py:705-715 encode, :1136-1140 decode).  Activations are NHWC bf16 so that a 3x3
// convolution is an implicit GEMM over pixels (gemm_sm100.cu, conv mode) and a 1x1 convolution / attention projection
// is a plain GEMM.  Here: GroupNorm (+ SiLU), nearest x2 upsampling, the row softmax of the single-head mid-block
// attention, and the im2col gather used for the few convolutions TMA cannot address (stride 2, 3 input channels).
#include <cmath>

#include "dtype_utils.cuh"
#include "rt_internal.h"

namespace rt {
namespace {

int sm_count_v() { return device_sm_count(); }

// ---- GroupNorm statistics: x [B, HW, C] bf16, stats [B, G, 2] double (sum, sum of squares), zeroed by the launcher.
// A thread owns 8 consecutive channels (one 16-byte vector) of a strided set of pixels; C / G is a multiple of ... 1:
// channel c belongs to group c / (C / G), so a vector may straddle groups only if C / G < 8 (then 8 % (C / G) == 0).
__global__ void __launch_bounds__(256) gn_stats_kernel(const bf16* __restrict__ x, long long hw, int C, int G,
                                                       double* __restrict__ stats, int pixels_per_block) {
  const int b = blockIdx.y;
  const int vecs = C / 8;                 // vectors per pixel
  const int cpg = C / G;                  // channels per group
  const int v = threadIdx.x % vecs;       // this thread's vector (blockDim.x is a multiple of vecs)
  const int lanes = blockDim.x / vecs;    // pixels processed per step by the block
  const long long p0 = (long long)blockIdx.x * pixels_per_block;
  const long long p1 = min(p0 + pixels_per_block, hw);
  float s[8], q[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = q[j] = 0.f;
  const bf16* base = x + (long long)b * hw * C + v * 8;
  long long p = p0 + threadIdx.x / vecs;
  for (; p + 3LL * lanes < p1; p += 4LL * lanes) {      // four independent 16-byte loads in flight per thread
    float t[4][8];
#pragma unroll
    for (int u = 0; u < 4; ++u) ldvec(base + (p + (long long)u * lanes) * C, t[u]);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
#pragma unroll
      for (int j = 0; j < 8; ++j) { s[j] += t[u][j]; q[j] = fmaf(t[u][j], t[u][j], q[j]); }
    }
  }
  for (; p < p1; p += lanes) {
    float t[8];
    ldvec(base + p * C, t);
#pragma unroll
    for (int j = 0; j < 8; ++j) { s[j] += t[j]; q[j] = fmaf(t[j], t[j], q[j]); }
  }
  // block-level fold in shared memory (per channel), then ONE pair of double atomics per (block, group)
  __shared__ float ssum[512], ssq[512];
  for (int c = threadIdx.x; c < C; c += blockDim.x) { ssum[c] = 0.f; ssq[c] = 0.f; }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    atomicAdd(&ssum[v * 8 + j], s[j]);
    atomicAdd(&ssq[v * 8 + j], q[j]);
  }
  __syncthreads();
  for (int g = threadIdx.x; g < G; g += blockDim.x) {
    double a = 0.0, b2 = 0.0;
    for (int c = g * cpg; c < (g + 1) * cpg; ++c) { a += (double)ssum[c]; b2 += (double)ssq[c]; }
    atomicAdd(&stats[((long long)b * G + g) * 2 + 0], a);
    atomicAdd(&stats[((long long)b * G + g) * 2 + 1], b2);
  }
}

// (sum, sum of squares) -> (mean, rstd) per (batch, group)
__global__ void gn_finalize_kernel(const double* __restrict__ stats, float2* __restrict__ mr, int n_groups, double n,
                                   float eps) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_groups) return;
  const double mean = stats[2 * i] / n;
  const double var = fmax(stats[2 * i + 1] / n - mean * mean, 0.0);
  mr[i] = make_float2((float)mean, (float)(1.0 / sqrt(var + (double)eps)));
}

// y = act((x - mean) * rstd * gamma + beta), act = SiLU or identity.  Like the statistics kernel, a thread owns ONE
// 16-byte channel vector of a strided set of pixels, so the affine form y = a * x + b (a = rstd * gamma,
// b = beta - mean * a) is computed once per thread and the pixel loop is load - 8 FMAs - (8 tanh) - store, four pixels
// in flight per thread.  silu(y) = y * sigmoid(y) = 0.5 * y * (1 + tanh(y / 2)): one MUFU op per element.
__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <bool kSilu>
__global__ void __launch_bounds__(256) gn_apply_kernel(const bf16* __restrict__ x, bf16* __restrict__ out, long long hw,
                                                       int C, int G, const float2* __restrict__ mr,
                                                       const bf16* __restrict__ gamma, const bf16* __restrict__ beta,
                                                       int pixels_per_block) {
  const int b = blockIdx.y;
  const int vecs = C / 8, cpg = C / G;
  const int v = threadIdx.x % vecs;
  const int lanes = blockDim.x / vecs;
  float ga[8], be[8], ca[8], cb[8];
  ldvec(gamma + v * 8, ga);
  ldvec(beta + v * 8, be);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float2 m = __ldg(&mr[(long long)b * G + (v * 8 + j) / cpg]);
    ca[j] = m.y * ga[j];
    cb[j] = be[j] - m.x * ca[j];
  }
  const long long p0 = (long long)blockIdx.x * pixels_per_block;
  const long long p1 = min(p0 + pixels_per_block, hw);
  const bf16* xb = x + (long long)b * hw * C + v * 8;
  bf16* ob = out + (long long)b * hw * C + v * 8;
  long long p = p0 + threadIdx.x / vecs;
  for (; p + 3LL * lanes < p1; p += 4LL * lanes) {
    float t[4][8];
#pragma unroll
    for (int u = 0; u < 4; ++u) ldvec(xb + (p + (long long)u * lanes) * C, t[u]);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float y = fmaf(t[u][j], ca[j], cb[j]);
        if (kSilu) y = 0.5f * y * (1.f + tanh_approx(0.5f * y));
        t[u][j] = y;
      }
      stvec(ob + (p + (long long)u * lanes) * C, t[u]);
    }
  }
  for (; p < p1; p += lanes) {
    float t[8];
    ldvec(xb + p * C, t);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float y = fmaf(t[j], ca[j], cb[j]);
      if (kSilu) y = 0.5f * y * (1.f + tanh_approx(0.5f * y));
      t[j] = y;
    }
    stvec(ob + p * C, t);
  }
}

// nearest-neighbour x2: in [B, H, W, C] -> out [B, 2H, 2W, C]
__global__ void __launch_bounds__(256) upsample2x_kernel(const bf16* __restrict__ in, bf16* __restrict__ out, int H, int W,
                                                         int C, long long total_vecs) {
  const int vecs = C / 8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total_vecs;
       i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % vecs);
    long long p = i / vecs;                // output pixel index: (b, y, x)
    const int x = (int)(p % (2 * W));
    p /= 2 * W;
    const int y = (int)(p % (2 * H));
    const long long b = p / (2 * H);
    const uint4 t = *reinterpret_cast<const uint4*>(in + ((b * H + y / 2) * W + x / 2) * C + v * 8);
    *reinterpret_cast<uint4*>(out + (i / vecs) * C + v * 8) = t;
  }
}

// softmax over each row of a [rows, cols] bf16 matrix, in place (fp32 arithmetic); one CTA per row
__global__ void __launch_bounds__(1024) softmax_rows_kernel(bf16* __restrict__ x, long long ld, int cols) {
  __shared__ float red[32];
  bf16* row = x + (long long)blockIdx.x * ld;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  float mx = -INFINITY;
  for (int c = threadIdx.x * 8; c < cols; c += blockDim.x * 8) {
    float t[8];
    ldvec(row + c, t);
#pragma unroll
    for (int j = 0; j < 8; ++j) mx = fmaxf(mx, t[j]);
  }
  mx = warp_max(mx);
  if (lane == 0) red[warp] = mx;
  __syncthreads();
  mx = red[0];
  for (int w = 1; w < nw; ++w) mx = fmaxf(mx, red[w]);
  __syncthreads();
  float sum = 0.f;
  for (int c = threadIdx.x * 8; c < cols; c += blockDim.x * 8) {
    float t[8];
    ldvec(row + c, t);
#pragma unroll
    for (int j = 0; j < 8; ++j) sum += __expf(t[j] - mx);
  }
  sum = warp_sum(sum);
  if (lane == 0) red[warp] = sum;
  __syncthreads();
  sum = 0.f;
  for (int w = 0; w < nw; ++w) sum += red[w];
  const float inv = 1.f / sum;
  for (int c = threadIdx.x * 8; c < cols; c += blockDim.x * 8) {
    float t[8], o[8];
    ldvec(row + c, t);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = __expf(t[j] - mx) * inv;
    stvec(row + c, o);
  }
}

// softmax over each row of an fp32 [rows, cols] matrix -> bf16; one CTA per row, the row lives in registers (kV float4
// per thread): ONE read of the logits, which never see a bf16 rounding (the reference's SDPA keeps them in fp32)
template <int kV>
__global__ void __launch_bounds__(1024) softmax_rows_f32_kernel(const float* __restrict__ x, long long ld_in,
                                                                bf16* __restrict__ out, long long ld_out, int cols) {
  __shared__ float red[32];
  const float4* row = reinterpret_cast<const float4*>(x + (long long)blockIdx.x * ld_in);
  bf16* orow = out + (long long)blockIdx.x * ld_out;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int nvec = cols >> 2;
  float4 v[kV];
  float mx = -INFINITY;
#pragma unroll
  for (int i = 0; i < kV; ++i) {
    const int c = threadIdx.x + i * blockDim.x;
    v[i] = c < nvec ? row[c] : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
    mx = fmaxf(fmaxf(mx, fmaxf(v[i].x, v[i].y)), fmaxf(v[i].z, v[i].w));
  }
  mx = warp_max(mx);
  if (lane == 0) red[warp] = mx;
  __syncthreads();
  mx = red[0];
  for (int w = 1; w < nw; ++w) mx = fmaxf(mx, red[w]);
  __syncthreads();
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < kV; ++i) {
    v[i].x = __expf(v[i].x - mx); v[i].y = __expf(v[i].y - mx); v[i].z = __expf(v[i].z - mx); v[i].w = __expf(v[i].w - mx);
    sum += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  }
  sum = warp_sum(sum);
  if (lane == 0) red[warp] = sum;
  __syncthreads();
  sum = 0.f;
  for (int w = 0; w < nw; ++w) sum += red[w];
  const float inv = 1.f / sum;
#pragma unroll
  for (int i = 0; i < kV; ++i) {
    const int c = threadIdx.x + i * blockDim.x;
    if (c < nvec) {
      __nv_bfloat162 lo = __floats2bfloat162_rn(v[i].x * inv, v[i].y * inv), hi = __floats2bfloat162_rn(v[i].z * inv, v[i].w * inv);
      uint2 u;
      u.x = *reinterpret_cast<uint32_t*>(&lo);
      u.y = *reinterpret_cast<uint32_t*>(&hi);
      *reinterpret_cast<uint2*>(orow + 4 * c) = u;
    }
  }
}

// im2col for a 3x3 convolution with stride `stride` and padding (top / left = pad_lo, bottom / right as needed):
// in [B, H, W, Cin_ld] (first C channels used) -> out [B, Ho * Wo, Kp], K index = tap * C + c, zero beyond 9 * C
__global__ void __launch_bounds__(256) im2col3x3_kernel(const bf16* __restrict__ in, bf16* __restrict__ out, int H, int W,
                                                        int C, int c_ld, int Ho, int Wo, int stride, int pad_lo, int Kp,
                                                        long long total) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int k = (int)(i % Kp);
    long long p = i / Kp;
    const int xo = (int)(p % Wo);
    p /= Wo;
    const int yo = (int)(p % Ho);
    const long long b = p / Ho;
    bf16 v = __float2bfloat16(0.f);
    if (k < 9 * C) {
      const int tap = k / C, c = k - tap * C;
      const int y = yo * stride + tap / 3 - pad_lo, x = xo * stride + tap % 3 - pad_lo;
      if (y >= 0 && y < H && x >= 0 && x < W) v = in[((b * H + y) * W + x) * c_ld + c];
    }
    out[i] = v;
  }
}


// NCHW (fp32 or bf16) -> NHWC bf16 [B, H * W, c_pad], channels >= C zero: the image / latent entering the VAE
template <typename T>
__global__ void __launch_bounds__(256) nchw_to_nhwc_kernel(const T* __restrict__ in, bf16* __restrict__ out, int C,
                                                           long long hw, int c_pad, long long total) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % c_pad);
    const long long bp = i / c_pad;       // b * hw + p
    const long long b = bp / hw, p = bp - b * hw;
    float v = 0.f;
    if (c < C) v = to_f(in[(b * C + c) * hw + p]);
    out[i] = __float2bfloat16(v);
  }
}

// NHWC bf16 [B, H * W, ld] (first C channels) -> NCHW (fp32 or bf16): the image / moments leaving the VAE.
// A warp reads 32 pixels of one channel-block through shared memory so that both sides are coalesced.
template <typename T>
__global__ void __launch_bounds__(256) nhwc_to_nchw_kernel(const bf16* __restrict__ in, T* __restrict__ out, int C, int ld,
                                                           long long hw) {
  __shared__ float tile[32][33];
  const long long b = blockIdx.z;
  const long long p0 = (long long)blockIdx.x * 32;
  const int c0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const long long p = p0 + r;
    const int c = c0 + threadIdx.x;
    tile[r][threadIdx.x] = (p < hw && c < C) ? __bfloat162float(in[(b * hw + p) * ld + c]) : 0.f;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int c = c0 + r;
    const long long p = p0 + threadIdx.x;
    if (c < C && p < hw) out[(b * C + c) * hw + p] = from_f<T>(tile[threadIdx.x][r]);
  }
}

// DiagonalGaussianDistribution.sample on NHWC moments [B, hw, ld] (mean = channels 0..L, logvar = L..2L):
// out NCHW [B, L, hw] = mean + exp(0.5 * clamp(logvar, -30, 20)) * noise (noise NCHW, same dtype as out; NULL = mode)
template <typename T>
__global__ void __launch_bounds__(256) posterior_sample_kernel(const bf16* __restrict__ mom, int ld, int L, long long hw,
                                                               const T* __restrict__ noise, T* __restrict__ out,
                                                               long long total) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const long long p = i % hw;
    const long long bc = i / hw;
    const int c = (int)(bc % L);
    const long long b = bc / L;
    const bf16* m = mom + (b * hw + p) * ld;
    float v = __bfloat162float(m[c]);
    if (noise) {
      const float lv = fminf(fmaxf(__bfloat162float(m[L + c]), -30.f), 20.f);
      v += expf(0.5f * lv) * to_f(noise[i]);
    }
    out[i] = from_f<T>(v);
  }
}

// the same gather, 8 channels (16 bytes) per thread: C, c_ld multiples of 8 and Kp == 9 * C
__global__ void __launch_bounds__(256) im2col3x3_vec8_kernel(const bf16* __restrict__ in, bf16* __restrict__ out, int H,
                                                             int W, int C, int c_ld, int Ho, int Wo, int stride,
                                                             int pad_lo, long long total) {
  const int vecs = C / 8;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int v = (int)(i % vecs);
    long long r = i / vecs;
    const int tap = (int)(r % 9);
    long long p = r / 9;                    // (b, yo, xo)
    const long long pix = p;
    const int xo = (int)(p % Wo);
    p /= Wo;
    const int yo = (int)(p % Ho);
    const long long b = p / Ho;
    const int y = yo * stride + tap / 3 - pad_lo, x = xo * stride + tap % 3 - pad_lo;
    uint4 t = make_uint4(0u, 0u, 0u, 0u);
    if (y >= 0 && y < H && x >= 0 && x < W)
      t = *reinterpret_cast<const uint4*>(in + ((b * H + y) * W + x) * c_ld + v * 8);
    *reinterpret_cast<uint4*>(out + (pix * 9 + tap) * C + v * 8) = t;
  }
}

int grid_for(long long n, int threads) {
  long long b = (n + threads - 1) / threads;
  const long long cap = (long long)sm_count_v() * 16;
  return (int)(b > cap ? cap : (b < 1 ? 1 : b));
}

}  // namespace
}  // namespace rt

using namespace rt;

extern "C" {

int rt_groupnorm_nhwc(const void* x, void* out, int batch, int64_t hw, int C, int groups, const void* gamma,
                      const void* beta, float eps, int silu, void* stats_ws, void* stream) {
  return guarded([&] {
    RT_REQUIRE(x && out && gamma && beta && stats_ws, "groupnorm: null argument");
    RT_REQUIRE(batch >= 1 && hw >= 1 && C % 8 == 0 && groups >= 1 && C % groups == 0, "groupnorm: bad shape");
    const int cpg = C / groups;
    RT_REQUIRE(cpg % 8 == 0 || 8 % cpg == 0, "groupnorm: channels per group must divide or be a multiple of 8");
    const int vecs = C / 8;
    RT_REQUIRE(vecs <= 256 && 256 % vecs == 0, "groupnorm: C / 8 must divide 256");
    cudaStream_t s = (cudaStream_t)stream;
    ProfScope ps(PROF_ELEM, 3.0 * batch * (double)hw * C * 2, s);
    RT_REQUIRE(C <= 512, "groupnorm: C <= 512");
    RT_CHECK_CUDA(cudaMemsetAsync(stats_ws, 0, (size_t)batch * groups * 2 * sizeof(double), s));
    float2* mr = reinterpret_cast<float2*>(reinterpret_cast<double*>(stats_ws) + (size_t)batch * groups * 2);
    const int lanes = 256 / vecs;
    long long ppb = (hw + (long long)sm_count_v() * 4 - 1) / ((long long)sm_count_v() * 4);
    ppb = (ppb + lanes - 1) / lanes * lanes;
    if (ppb < lanes) ppb = lanes;
    dim3 grid((unsigned)((hw + ppb - 1) / ppb), (unsigned)batch);
    gn_stats_kernel<<<grid, 256, 0, s>>>((const bf16*)x, hw, C, groups, (double*)stats_ws, (int)ppb);
    RT_POST_LAUNCH();
    gn_finalize_kernel<<<(batch * groups + 127) / 128, 128, 0, s>>>((const double*)stats_ws, mr, batch * groups,
                                                                    (double)hw * cpg, eps);
    RT_POST_LAUNCH();
    if (silu)
      gn_apply_kernel<true><<<grid, 256, 0, s>>>((const bf16*)x, (bf16*)out, hw, C, groups, mr, (const bf16*)gamma,
                                                 (const bf16*)beta, (int)ppb);
    else
      gn_apply_kernel<false><<<grid, 256, 0, s>>>((const bf16*)x, (bf16*)out, hw, C, groups, mr, (const bf16*)gamma,
                                                  (const bf16*)beta, (int)ppb);
    RT_POST_LAUNCH();
  });
}

int rt_upsample_nearest2x_nhwc(const void* in, void* out, int batch, int H, int W, int C, void* stream) {
  return guarded([&] {
    RT_REQUIRE(in && out && batch >= 1 && H >= 1 && W >= 1 && C % 8 == 0, "upsample: bad argument");
    const long long total = (long long)batch * 4 * H * W * (C / 8);
    ProfScope ps(PROF_ELEM, (double)total * 16 * 1.25, (cudaStream_t)stream);
    upsample2x_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>((const bf16*)in, (bf16*)out, H, W, C, total);
    RT_POST_LAUNCH();
  });
}

int rt_softmax_rows_f32(const void* x, int64_t rows, int cols, int64_t ld_in, void* out, int64_t ld_out, void* stream) {
  return guarded([&] {
    RT_REQUIRE(x && out && rows >= 1 && cols >= 4 && cols % 4 == 0 && cols <= 65536 && ld_in % 4 == 0 && ld_out % 4 == 0,
               "softmax_rows_f32: bad argument");
    RT_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 7) == 0, "softmax_rows_f32: alignment");
    const int nvec = cols / 4;
    const int threads = nvec >= 1024 ? 1024 : (nvec + 31) / 32 * 32;
    const int per = (nvec + threads - 1) / threads;  // float4 per thread
    cudaStream_t s = (cudaStream_t)stream;
    const float* xf = (const float*)x;
    bf16* ob = (bf16*)out;
    if (per <= 1) softmax_rows_f32_kernel<1><<<(unsigned)rows, threads, 0, s>>>(xf, ld_in, ob, ld_out, cols);
    else if (per <= 2) softmax_rows_f32_kernel<2><<<(unsigned)rows, threads, 0, s>>>(xf, ld_in, ob, ld_out, cols);
    else if (per <= 4) softmax_rows_f32_kernel<4><<<(unsigned)rows, threads, 0, s>>>(xf, ld_in, ob, ld_out, cols);
    else if (per <= 8) softmax_rows_f32_kernel<8><<<(unsigned)rows, threads, 0, s>>>(xf, ld_in, ob, ld_out, cols);
    else softmax_rows_f32_kernel<16><<<(unsigned)rows, threads, 0, s>>>(xf, ld_in, ob, ld_out, cols);
    RT_POST_LAUNCH();
  });
}

int rt_softmax_rows(void* x, int64_t rows, int cols, int64_t ld, void* stream) {
  return guarded([&] {
    RT_REQUIRE(x && rows >= 1 && cols >= 8 && cols % 8 == 0 && ld % 8 == 0, "softmax_rows: bad argument");
    ProfScope ps(PROF_ELEM, 4.0 * rows * (double)cols * 2, (cudaStream_t)stream);
    const int threads = cols >= 8192 ? 1024 : (cols >= 2048 ? 256 : 128);
    softmax_rows_kernel<<<(unsigned)rows, threads, 0, (cudaStream_t)stream>>>((bf16*)x, ld, cols);
    RT_POST_LAUNCH();
  });
}

int rt_im2col3x3_nhwc(const void* in, void* out, int batch, int H, int W, int C, int c_ld, int Ho, int Wo, int stride,
                      int pad_lo, int Kp, void* stream) {
  return guarded([&] {
    RT_REQUIRE(in && out && batch >= 1 && C >= 1 && c_ld >= C && Kp >= 9 * C && Kp % 8 == 0 && stride >= 1,
               "im2col: bad argument");
    const long long total = (long long)batch * Ho * Wo * Kp;
    ProfScope ps(PROF_ELEM, (double)total * 2 * 2, (cudaStream_t)stream);
    if (C % 8 == 0 && c_ld % 8 == 0 && Kp == 9 * C && (reinterpret_cast<uintptr_t>(in) & 15) == 0 &&
        (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
      const long long tv = total / 8;
      im2col3x3_vec8_kernel<<<grid_for(tv, 256), 256, 0, (cudaStream_t)stream>>>((const bf16*)in, (bf16*)out, H, W, C, c_ld,
                                                                                Ho, Wo, stride, pad_lo, tv);
    } else {
      im2col3x3_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>((const bf16*)in, (bf16*)out, H, W, C, c_ld, Ho,
                                                                              Wo, stride, pad_lo, Kp, total);
    }
    RT_POST_LAUNCH();
  });
}

int rt_nchw_to_nhwc(int src_dtype, const void* in, void* out, int batch, int C, int64_t hw, int c_pad, void* stream) {
  return guarded([&] {
    RT_REQUIRE(in && out && batch >= 1 && C >= 1 && hw >= 1 && c_pad >= C, "nchw_to_nhwc: bad argument");
    const long long total = (long long)batch * hw * c_pad;
    cudaStream_t s = (cudaStream_t)stream;
    ProfScope ps(PROF_ELEM, (double)total * 2 * 2, s);
    if (src_dtype == RT_F32)
      nchw_to_nhwc_kernel<float><<<grid_for(total, 256), 256, 0, s>>>((const float*)in, (bf16*)out, C, hw, c_pad, total);
    else if (src_dtype == RT_BF16)
      nchw_to_nhwc_kernel<bf16><<<grid_for(total, 256), 256, 0, s>>>((const bf16*)in, (bf16*)out, C, hw, c_pad, total);
    else
      throw Error(RT_ERR_INVALID, "nchw_to_nhwc: dtype");
    RT_POST_LAUNCH();
  });
}

int rt_nhwc_to_nchw(const void* in, int ld, void* out, int dst_dtype, int batch, int C, int64_t hw, void* stream) {
  return guarded([&] {
    RT_REQUIRE(in && out && batch >= 1 && C >= 1 && hw >= 1 && ld >= C, "nhwc_to_nchw: bad argument");
    cudaStream_t s = (cudaStream_t)stream;
    ProfScope ps(PROF_ELEM, (double)batch * hw * C * 2 * 2, s);
    dim3 grid((unsigned)((hw + 31) / 32), (unsigned)((C + 31) / 32), (unsigned)batch), block(32, 8);
    if (dst_dtype == RT_F32)
      nhwc_to_nchw_kernel<float><<<grid, block, 0, s>>>((const bf16*)in, (float*)out, C, ld, hw);
    else if (dst_dtype == RT_BF16)
      nhwc_to_nchw_kernel<bf16><<<grid, block, 0, s>>>((const bf16*)in, (bf16*)out, C, ld, hw);
    else
      throw Error(RT_ERR_INVALID, "nhwc_to_nchw: dtype");
    RT_POST_LAUNCH();
  });
}

int rt_vae_posterior_sample(const void* moments, int ld, int latent_channels, int batch, int64_t hw, const void* noise,
                            void* out, int dtype, void* stream) {
  return guarded([&] {
    RT_REQUIRE(moments && out && batch >= 1 && hw >= 1 && latent_channels >= 1 && ld >= 2 * latent_channels,
               "posterior_sample: bad argument");
    const long long total = (long long)batch * latent_channels * hw;
    cudaStream_t s = (cudaStream_t)stream;
    ProfScope ps(PROF_ELEM, (double)total * 2 * 4, s);
    if (dtype == RT_F32)
      posterior_sample_kernel<float><<<grid_for(total, 256), 256, 0, s>>>((const bf16*)moments, ld, latent_channels, hw,
                                                                        (const float*)noise, (float*)out, total);
    else if (dtype == RT_BF16)
      posterior_sample_kernel<bf16><<<grid_for(total, 256), 256, 0, s>>>((const bf16*)moments, ld, latent_channels, hw,
                                                                       (const bf16*)noise, (bf16*)out, total);
    else
      throw Error(RT_ERR_INVALID, "posterior_sample: dtype");
    RT_POST_LAUNCH();
  });
}

}  // extern "C"
