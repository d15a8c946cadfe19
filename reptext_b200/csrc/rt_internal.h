// Internal declarations shared by the kernels, the C++ runtime (model.cu) and the C-ABI (api.cu).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>

#include "../../include/reptext_rt.h"

namespace rt {

// ------------------------------------------------------------------ errors
void set_last_error(const std::string& msg);
struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};
#define RT_CHECK_CUDA(expr)                                                                              \
  do {                                                                                                   \
    cudaError_t _e = (expr);                                                                             \
    if (_e != cudaSuccess)                                                                               \
      throw ::rt::Error(RT_ERR_CUDA, std::string(#expr) + " -> " + cudaGetErrorString(_e) + " (" + __FILE__ + \
                                         ":" + std::to_string(__LINE__) + ")");                          \
  } while (0)
#define RT_REQUIRE(cond, msg)                                                              \
  do {                                                                                     \
    if (!(cond)) throw ::rt::Error(RT_ERR_INVALID, std::string(msg) + " [" #cond "]");     \
  } while (0)

inline size_t dtype_size(int dt) { return dt == RT_BF16 ? 2 : 4; }

// ------------------------------------------------------------------ per-device state
// cudaFuncSetAttribute and the SM count belong to ONE device; a process may place models on several
// (from_pretrained(device=...)), so one-time set-up is remembered per device, keyed by the device that is
// current at the launch (the Python layer makes the tensors' device current around every native call).
inline int current_device() {
  int dev = 0;
  RT_CHECK_CUDA(cudaGetDevice(&dev));
  return dev;
}
struct PerDeviceOnce {  // if (once.first()) { ...set attributes...; }  - true once per device
  unsigned long long done = 0;
  bool first() {
    const int dev = current_device();
    if (done >> dev & 1) return false;
    done |= 1ull << dev;
    return true;
  }
};
int device_sm_count();  // of the current device (api.cu)
int get_option(const char* name);

// run f, translating exceptions into a status code + rt_last_error()
template <typename F>
inline int guarded(F&& f) {
  try {
    f();
    return RT_OK;
  } catch (const Error& e) {
    set_last_error(e.what());
    return e.code;
  } catch (const std::exception& e) {
    set_last_error(e.what());
    return RT_ERR_INTERNAL;
  }
}

// ------------------------------------------------------------------ GEMM description
// out[b, out_row0 + m, out_col0 + (n - n_begin)] = epilogue( sum_k A[b, a_row0 + m, k] * W_seg[n - n_begin, k] )
// (structs are the public ones of include/reptext_rt.h)
enum EpiMode : int {
  EPI_BIAS = RT_EPI_BIAS,
  EPI_GELU = RT_EPI_GELU,
  EPI_QKNORM_ROPE = RT_EPI_QKNORM_ROPE,  // fused form exists on the tcgen05 path only
  EPI_GATE_RESID = RT_EPI_GATE_RESID,
  EPI_SCALE_MASK = RT_EPI_SCALE_MASK,
};
using GemmSegment = rt_gemm_segment;
using GemmProblem = rt_gemm_problem;
using GemmLaunch = rt_gemm_launch;

int gemm_total_n(const GemmProblem& p);

// kernels (each launches on `stream`; throws rt::Error)
void launch_gemm_simt(const GemmLaunch& g, cudaStream_t stream);
bool gemm_tc_supported(const GemmLaunch& g, std::string* why);
struct SpSyncParams;  // sp_sync.cuh: sequence-parallel phase synchronisation inside the tcgen05 kernels (null: none)
void launch_gemm_tc(const GemmLaunch& g, cudaStream_t stream, int force_cta_group /*0 auto,1,2*/,
                    const SpSyncParams* sync = nullptr);
// picks tcgen05 when supported; `sync` on the other paths becomes stand-alone kernels (same protocol)
void launch_gemm(const GemmLaunch& g, cudaStream_t stream, const SpSyncParams* sync = nullptr);
extern long long g_launch_count;                             // kernels launched by this library

// ------------------------------------------------------------------ other kernels
using LnModGroup = rt_lnmod_group;
// out[b, r, :] = LN(x[b, r, :]) * (1 + scale) + shift   (eps 1e-6, no affine)
void launch_ln_mod(int dtype, const void* x, long long x_bs, int x_ld, void* out, long long o_bs, int o_ld, int batch,
                   int D, int ngroups, const LnModGroup* groups, cudaStream_t stream);

// in-place per-head RMSNorm * w and RoPE on columns [col0, col0 + heads*hd) of buf (SIMT path)
void launch_qknorm_rope(int dtype, void* buf, long long bs, int ld, int col0, int batch, int row0, int rows,
                        int heads, int hd, const void* norm_w, const float2* rope, int rope_row0,
                        cudaStream_t stream);

// O[b, r, h*hd + d] = softmax(Q K^T / sqrt(hd)) V over the joint sequence; q/k/v live in one buffer
using AttnArgs = rt_attention_args;
void launch_attention_simt(const AttnArgs& a, cudaStream_t stream);
bool attention_tc_supported(const AttnArgs& a, std::string* why);
void launch_attention_tc(const AttnArgs& a, cudaStream_t stream, int variant, const SpSyncParams* sync = nullptr);
void launch_attention(const AttnArgs& a, cudaStream_t stream, const SpSyncParams* sync = nullptr);

// grouped GEMV: out[b, off + r] = dot(act(x[b, :]), W[r, :]) + bias[r]   (fp32 in/out, weights dtype T)
struct GemvJob {
  const void* W;
  const void* bias;
  int rows;
  int out_off;
};
struct GemvPeers {  // sequence-parallel row shard: every result is stored to all `n` ranks' output buffers
  float* p[RT_SP_MAX_RANKS];
  int n;
};
void launch_gemv_grouped(int wdtype, const float* x, int x_ld, int batch, int K, const GemvJob* jobs_dev,
                         const int* job_row_prefix_dev, int njobs, int total_rows, float* out, int out_ld,
                         int silu_out, int accumulate, cudaStream_t stream, bool rows_multiple_of_4 = false,
                         int row_base = 0, const GemvPeers* peers = nullptr);
// the same rows for `batch` activation vectors, weights read from DRAM once (4-row jobs; absolute rows, no accumulate)
void launch_gemv_grouped_table(int wdtype, const float* x, int x_ld, int batch, int K, const GemvJob* jobs_dev,
                               const int* job_row_prefix_dev, int njobs, int total_rows, float* out, int out_ld,
                               cudaStream_t stream);
void launch_silu_f32(const float* x, float* out, long long n, cudaStream_t stream);

void launch_time_sinusoid(int dtype, const void* t, int t_batch, int batch, float* out /*[batch,256]*/,
                          cudaStream_t stream);
void launch_cast_to_f32(int dtype, const void* x, float* out, long long n, cudaStream_t stream);
void launch_rope_table(const float* ids /*[S,3]*/, int S, const int* axes /*3*/, float2* out /*[S, hd/2]*/,
                       cudaStream_t stream);
void launch_euler_step(int dtype, const void* v, const void* x, void* out, long long n, float sigma,
                       float sigma_next, cudaStream_t stream);
void launch_cfg_euler(int dtype, const void* v2 /*[2,n] uncond,text*/, const void* x, void* out, long long n,
                      float true_scale, int zero_pred, float sigma, float sigma_next, cudaStream_t stream);
void launch_cfg_combine(int dtype, const void* v2, void* out, long long n, float true_scale, int zero_pred,
                        cudaStream_t stream);
void launch_mask_scale_add(int dtype, const void* x, const void* mask, const void* acc_in, void* out, int batch,
                           int rows, int D, float scale, cudaStream_t stream);
void launch_glyph_blend(int dtype, const void* noise, const void* glyph_lat, const unsigned char* mask, void* out,
                        long long n, float w_glyph, float w_noise, cudaStream_t stream);
void launch_copy_rows(int dtype, const void* src, long long s_bs, int s_ld, int s_row0, void* dst, long long d_bs,
                      int d_ld, int d_row0, int batch, int rows, int D, cudaStream_t stream);

// cross-GPU flag barrier of the sequence-parallel mode (sp.cu)
void launch_sp_barrier(const rt_sp_group& g, cudaStream_t stream);
void launch_sp_sync_before(const SpSyncParams& s, cudaStream_t stream);  // the in-kernel form's barrier / announcement
void launch_sp_sync_after(const SpSyncParams& s, cudaStream_t stream);   // as kernels (launches off the tcgen05 path)

// ------------------------------------------------------------------ per-class device timing (option "profile")
// When the option is on, every launch of a class is bracketed by CUDA events on ITS stream; bench.py reads
// the per-class sums (rt_profile_read) for the roofline line.  Off by default (zero overhead).
enum ProfClass : int { PROF_GEMM_TC = 0, PROF_GEMM_SIMT, PROF_ATTN_TC, PROF_ATTN_SIMT, PROF_LN, PROF_GEMV, PROF_ELEM,
                       PROF_ATTN_TEXT, PROF_NCLS };
struct ProfScope {
  ProfScope(int cls, double work, cudaStream_t s);
  ~ProfScope();
  int cls_;
  double work_;
  cudaStream_t s_;
  cudaEvent_t e0_ = nullptr, e1_ = nullptr;
};

// ------------------------------------------------------------------ TMA descriptor encode (driver entry point)
void encode_tmap_bf16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_b,
                      const uint32_t* box);

}  // namespace rt
