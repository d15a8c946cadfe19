// CUDA-core (SIMT) GEMM and attention.  These serve (a) the fp32 tiny configuration, where the parity
// bar is 1e-4 and bf16 tensor cores cannot be used, and (b) shapes the tcgen05 kernels do not cover
// (K not a multiple of 8, head_dim != 128, ...).  They are not the performance path.
#include "dtype_utils.cuh"
#include "rt_internal.h"

namespace rt {

int gemm_total_n(const GemmProblem& p) {
  int n = 0;
  for (int i = 0; i < p.nseg; ++i) n = p.seg[i].n_end > n ? p.seg[i].n_end : n;
  return n;
}

// ------------------------------------------------------------------------------------------------
// 64x64x16 tiled GEMM, 256 threads, 4x4 outputs per thread, fp32 accumulate.
// ------------------------------------------------------------------------------------------------
struct SimtGemmArgs {
  const void* A; long long a_bs; int a_ld; int a_row0;
  const void* W; const void* bias;
  int M, N, K;
  int mode;
  void* out; long long o_bs; int o_ld; int o_row0; int o_col0;
  const float* gate; int gate_ld; int gate_col0;
  const void* extra; long long e_bs; int e_ld; int e_row0;
  float scale; const void* mask; int accumulate;
};

template <typename T>
__global__ void __launch_bounds__(256) gemm_simt_kernel(SimtGemmArgs g) {
  __shared__ float As[16][64 + 1];
  __shared__ float Ws[16][64 + 1];
  const int b = blockIdx.z;
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const T* A = reinterpret_cast<const T*>(g.A) + (long long)b * g.a_bs + (long long)g.a_row0 * g.a_ld;
  const T* W = reinterpret_cast<const T*>(g.W);
  float acc[4][4] = {};
  for (int k0 = 0; k0 < g.K; k0 += 16) {
    // 64 rows x 16 k = 1024 elements per operand, 4 per thread
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      int e = threadIdx.x + 256 * i;
      int r = e >> 4, k = e & 15;
      float av = 0.f, wv = 0.f;
      if (m0 + r < g.M && k0 + k < g.K) av = to_f(A[(long long)(m0 + r) * g.a_ld + k0 + k]);
      if (n0 + r < g.N && k0 + k < g.K) wv = to_f(W[(long long)(n0 + r) * g.K + k0 + k]);
      As[k][r] = av;
      Ws[k][r] = wv;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      float a[4], w[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        a[i] = As[k][ty * 4 + i];
        w[i] = Ws[k][tx * 4 + i];
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] += a[i] * w[j];
    }
    __syncthreads();
  }
  const T* bias = reinterpret_cast<const T*>(g.bias);
  T* out = reinterpret_cast<T*>(g.out) + (long long)b * g.o_bs;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int m = m0 + ty * 4 + i;
    if (m >= g.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int n = n0 + tx * 4 + j;
      if (n >= g.N) continue;
      float v = acc[i][j] + (bias ? to_f(bias[n]) : 0.f);
      T* o = out + (long long)(g.o_row0 + m) * g.o_ld + g.o_col0 + n;
      switch (g.mode) {
        case EPI_GELU:
          v = gelu_tanh_ref(v);
          break;
        case EPI_GATE_RESID: {
          float gt = g.gate ? g.gate[(long long)b * g.gate_ld + g.gate_col0 + n] : 1.f;
          v = to_f(*o) + gt * v;
          if (g.extra && m >= g.e_row0)
            v += to_f(reinterpret_cast<const T*>(g.extra)[(long long)b * g.e_bs + (long long)(m - g.e_row0) * g.e_ld +
                                                          g.gate_col0 + n]);
          break;
        }
        case EPI_SCALE_MASK: {
          float mk = g.mask ? to_f(reinterpret_cast<const T*>(g.mask)[m]) : 1.f;
          v = v * g.scale * mk;
          if (g.accumulate) v += to_f(*o);
          break;
        }
        default:
          break;
      }
      *o = from_f<T>(v);
    }
  }
}

void launch_gemm_simt(const GemmLaunch& L, cudaStream_t stream) {
  for (int pi = 0; pi < L.nprob; ++pi) {
    const GemmProblem& P = L.prob[pi];
    if (P.m_rows == 0 || L.batch == 0) continue;
    if (P.conv_w > 0) throw Error(RT_ERR_UNSUPPORTED, "the implicit 3x3 convolution exists on the tcgen05 GEMM only");
    for (int si = 0; si < P.nseg; ++si) {
      const GemmSegment& S = P.seg[si];
      if (S.scatter) throw Error(RT_ERR_UNSUPPORTED, "sequence-parallel scatter exists on the tcgen05 GEMM only");
      if (S.out_f32) throw Error(RT_ERR_UNSUPPORTED, "fp32 output of bf16 operands (out_f32) exists on the tcgen05 GEMM only");
      SimtGemmArgs g;
      g.A = P.A; g.a_bs = P.a_batch_stride; g.a_ld = P.a_ld; g.a_row0 = P.a_row0;
      g.W = S.W; g.bias = S.bias;
      g.M = P.m_rows; g.N = S.n_end - S.n_begin; g.K = P.K;
      g.mode = S.mode;
      g.out = S.out; g.o_bs = S.out_batch_stride; g.o_ld = S.out_ld; g.o_row0 = P.out_row0; g.o_col0 = S.out_col0;
      g.gate = P.gate; g.gate_ld = P.gate_ld; g.gate_col0 = S.n_begin;
      g.extra = P.extra; g.e_bs = P.extra_batch_stride; g.e_ld = P.extra_ld; g.e_row0 = P.extra_row0;
      g.scale = P.scale; g.mask = P.mask; g.accumulate = P.accumulate;
      dim3 grid((g.N + 63) / 64, (g.M + 63) / 64, L.batch);
      RT_DISPATCH_DTYPE(L.dtype, T, (gemm_simt_kernel<T><<<grid, 256, 0, stream>>>(g)));
      RT_POST_LAUNCH();
      if (S.mode == EPI_QKNORM_ROPE) {
        RT_REQUIRE(L.head_dim > 0 && g.N % L.head_dim == 0, "qknorm segment must be whole heads");
        launch_qknorm_rope(L.dtype, S.out, S.out_batch_stride, S.out_ld, S.out_col0, L.batch, P.out_row0, P.m_rows,
                           g.N / L.head_dim, L.head_dim, S.norm_w, reinterpret_cast<const float2*>(L.rope),
                           P.out_row0, stream);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Attention: block = (batch, head, 32 query rows), 8 warps x 4 rows; K/V tiles of 32 keys in smem.
// Non-causal, no mask, scale 1/sqrt(hd), fp32 online softmax.  (diffusers dispatch_attention_fn ->
// F.scaled_dot_product_attention)
// ------------------------------------------------------------------------------------------------
template <typename T, int HD>
__global__ void __launch_bounds__(256) attn_simt_kernel(AttnArgs a) {
  constexpr int KT = 32, RPW = 4, DPL = HD / 32;
  extern __shared__ float smem[];
  float* Ks = smem;                      // [KT][HD+1]
  float* Vs = Ks + KT * (HD + 1);        // [KT][HD]
  float* Qs = Vs + KT * HD;              // [32][HD]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int h = blockIdx.y, b = blockIdx.z;
  const int q0 = blockIdx.x * 32;
  const T* base = reinterpret_cast<const T*>(a.qkv) + (long long)b * a.batch_stride;
  const float scale = rsqrtf((float)HD);
  for (int e = threadIdx.x; e < 32 * HD; e += 256) {
    int r = e / HD, d = e % HD;
    Qs[e] = (q0 + r < a.S) ? to_f(base[(long long)(q0 + r) * a.ld + a.q_col0 + h * HD + d]) * scale : 0.f;
  }
  float m[RPW], l[RPW], o[RPW][DPL];
#pragma unroll
  for (int r = 0; r < RPW; ++r) {
    m[r] = -INFINITY;
    l[r] = 0.f;
#pragma unroll
    for (int i = 0; i < DPL; ++i) o[r][i] = 0.f;
  }
  for (int k0 = 0; k0 < a.S; k0 += KT) {
    __syncthreads();
    for (int e = threadIdx.x; e < KT * HD; e += 256) {
      int j = e / HD, d = e % HD;
      bool ok = k0 + j < a.S;
      long long off = (long long)(k0 + j) * a.ld + h * HD + d;
      Ks[j * (HD + 1) + d] = ok ? to_f(base[off + a.k_col0]) : 0.f;
      Vs[j * HD + d] = ok ? to_f(base[off + a.v_col0]) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < RPW; ++r) {
      const float* q = Qs + (warp * RPW + r) * HD;
      float s = 0.f;
#pragma unroll 8
      for (int d = 0; d < HD; ++d) s += q[d] * Ks[lane * (HD + 1) + d];
      if (k0 + lane >= a.S) s = -INFINITY;
      float mx = fmaxf(m[r], warp_max(s));
      float p = __expf(s - mx);
      float corr = __expf(m[r] - mx);
      float ps = warp_sum(p);
      l[r] = l[r] * corr + ps;
      m[r] = mx;
#pragma unroll
      for (int i = 0; i < DPL; ++i) o[r][i] *= corr;
#pragma unroll 8
      for (int j = 0; j < KT; ++j) {
        float pj = __shfl_sync(0xffffffffu, p, j);
#pragma unroll
        for (int i = 0; i < DPL; ++i) o[r][i] += pj * Vs[j * HD + lane + 32 * i];
      }
    }
  }
  T* ob = reinterpret_cast<T*>(a.out) + (long long)b * a.out_batch_stride;
#pragma unroll
  for (int r = 0; r < RPW; ++r) {
    int row = q0 + warp * RPW + r;
    if (row >= a.S) continue;
    float inv = 1.f / l[r];
#pragma unroll
    for (int i = 0; i < DPL; ++i)
      ob[(long long)row * a.out_ld + a.out_col0 + h * HD + lane + 32 * i] = from_f<T>(o[r][i] * inv);
  }
}

void launch_attention_simt(const AttnArgs& a, cudaStream_t stream) {
  if (a.batch == 0 || a.S == 0) return;
  if (a.sp_rows > 0) throw Error(RT_ERR_UNSUPPORTED, "sequence-parallel scatter exists on the tcgen05 attention only");
  RT_REQUIRE(a.hd == 64 || a.hd == 128, "attention: head_dim must be 64 or 128");
  dim3 grid((a.S + 31) / 32, a.heads, a.batch);
  size_t smem = (size_t)(32 * (a.hd + 1) + 32 * a.hd + 32 * a.hd) * sizeof(float);
  RT_DISPATCH_DTYPE(a.dtype, T, {
    if (a.hd == 64) {
      attn_simt_kernel<T, 64><<<grid, 256, smem, stream>>>(a);
    } else {
      RT_CHECK_CUDA(cudaFuncSetAttribute(attn_simt_kernel<T, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem));
      attn_simt_kernel<T, 128><<<grid, 256, smem, stream>>>(a);
    }
  });
  RT_POST_LAUNCH();
}

}  // namespace rt
