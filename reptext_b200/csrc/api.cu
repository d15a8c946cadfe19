// extern "C" surface of librt_reptext.so (see include/reptext_rt.h) — operator-level entry points,
// options and error reporting.  Model-level entry points live in model.cu.
#include <cstring>
#include <map>
#include <mutex>
#include <vector>

#include "dtype_utils.cuh"
#include "rt_internal.h"
#include "sp_sync.cuh"

namespace rt {

static thread_local std::string t_last_error;
void set_last_error(const std::string& msg) { t_last_error = msg; }

static std::map<std::string, int>& options() {
  static std::map<std::string, int> o = {{"force_simt", 0}, {"gemm_cta_group", 0}, {"attn_variant", 0},
                                         {"profile", 0}, {"ln_warp_rows", 0}, {"gemv_single_row", 0}, {"gemm_debug", 0}, {"mod_inline", 0}, {"sp_replicate_mod", 0}, {"no_pdl", 0}, {"ln_impl", 0}, {"text_attn_simt", 0}, {"text_attn_mma", 0}, {"gemm_band", 0}, {"gemm_dyn_bn", 0}, {"sp_sync_kernels", 0}, {"mod_debug_skip", 0}, {"gemm_epi_warps", 0}, {"euler_dt_host", 0}};
  return o;
}
int device_sm_count() {
  static int n[64] = {0};
  const int dev = current_device();
  RT_REQUIRE(dev >= 0 && dev < 64, "device index out of range");
  if (!n[dev]) {
    RT_CHECK_CUDA(cudaDeviceGetAttribute(&n[dev], cudaDevAttrMultiProcessorCount, dev));
    if (n[dev] <= 0) n[dev] = 148;
  }
  return n[dev];
}

int get_option(const char* name) {
  auto it = options().find(name);
  return it == options().end() ? 0 : it->second;
}

// ---- per-class timing ----------------------------------------------------------------------------
struct ProfRec {
  int cls;
  double work;
  cudaEvent_t e0, e1;
};
static std::vector<ProfRec>& prof_recs() {
  static std::vector<ProfRec> r;
  return r;
}
static std::mutex g_prof_mu;
ProfScope::ProfScope(int cls, double work, cudaStream_t s) : cls_(cls), work_(work), s_(s) {
  if (!get_option("profile")) return;
  if (cudaEventCreate(&e0_) != cudaSuccess || cudaEventCreate(&e1_) != cudaSuccess) { e0_ = e1_ = nullptr; return; }
  cudaEventRecord(e0_, s_);
}
ProfScope::~ProfScope() {
  if (!e0_) return;
  cudaEventRecord(e1_, s_);
  std::lock_guard<std::mutex> lk(g_prof_mu);
  prof_recs().push_back({cls_, work_, e0_, e1_});
}

static double gemm_flops(const GemmLaunch& g) {
  double f = 0;
  for (int p = 0; p < g.nprob; ++p)
    f += 2.0 * g.batch * (double)g.prob[p].m_rows * gemm_total_n(g.prob[p]) * g.prob[p].K;
  return f;
}

// `sync` (sequence-parallel mode): the kernel runs the phase barrier at its head and / or announces the next barrier's
// epoch (sp_sync.cuh).  Only the tcgen05 kernels carry that; on any other path the same two steps run as stand-alone
// kernels before / after the launch - every rank takes the same path (same shapes), so the epochs stay in step.
void launch_gemm(const GemmLaunch& g, cudaStream_t stream, const SpSyncParams* sync) {
  if (!get_option("force_simt") && gemm_tc_supported(g, nullptr)) {
    ProfScope ps(PROF_GEMM_TC, gemm_flops(g), stream);
    launch_gemm_tc(g, stream, get_option("gemm_cta_group"), sync);
  } else {
    if (sync) launch_sp_sync_before(*sync, stream);
    {
      ProfScope ps(PROF_GEMM_SIMT, gemm_flops(g), stream);
      launch_gemm_simt(g, stream);
    }
    if (sync) launch_sp_sync_after(*sync, stream);
  }
}

void launch_attention(const AttnArgs& a, cudaStream_t stream, const SpSyncParams* sync) {
  const double flops = 4.0 * a.batch * a.heads * (double)a.S * a.S * a.hd;
  const int variant = get_option("attn_variant");
  if (!get_option("force_simt") && attention_tc_supported(a, nullptr) && (!sync || variant == 0)) {
    ProfScope ps(PROF_ATTN_TC, flops, stream);
    launch_attention_tc(a, stream, variant, sync);
  } else {
    if (sync) launch_sp_sync_before(*sync, stream);
    if (!get_option("force_simt") && attention_tc_supported(a, nullptr)) {
      ProfScope ps(PROF_ATTN_TC, flops, stream);
      launch_attention_tc(a, stream, variant);
    } else {
      ProfScope ps(PROF_ATTN_SIMT, flops, stream);
      launch_attention_simt(a, stream);
    }
    if (sync) launch_sp_sync_after(*sync, stream);
  }
}

}  // namespace rt

using namespace rt;

extern "C" {

const char* rt_last_error(void) { return t_last_error.c_str(); }
int rt_abi_version(void) { return RT_ABI_VERSION; }
int rt_struct_size(int which) {
  switch (which) {
    case 0: return (int)sizeof(rt_model_config);
    case 1: return (int)sizeof(rt_forward_args);
    case 2: return (int)sizeof(rt_sp_group);
    case 3: return (int)sizeof(rt_controlnet_call);
    case 4: return (int)sizeof(rt_transformer_call);
    case 5: return (int)sizeof(rt_gemm_segment);
    case 6: return (int)sizeof(rt_gemm_problem);
    case 7: return (int)sizeof(rt_gemm_launch);
    case 8: return (int)sizeof(rt_attention_args);
    case 9: return (int)sizeof(rt_lnmod_group);
    default: return -1;
  }
}
long long rt_launch_count(void) { return g_launch_count; }

int rt_set_option(const char* name, int value) {
  return guarded([&] {
    RT_REQUIRE(name && options().count(name), "unknown option");
    options()[name] = value;
  });
}
int rt_get_option(const char* name, int* value) {
  return guarded([&] {
    RT_REQUIRE(name && value && options().count(name), "unknown option");
    *value = options()[name];
  });
}

int rt_profile_reset(void) {
  return guarded([&] {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    for (auto& r : prof_recs()) {
      cudaEventDestroy(r.e0);
      cudaEventDestroy(r.e1);
    }
    prof_recs().clear();
  });
}
int rt_profile_read(int cls, double* ms, double* work, long long* count) {
  return guarded([&] {
    RT_REQUIRE(cls >= 0 && cls < PROF_NCLS && ms && work && count, "profile_read: bad argument");
    std::lock_guard<std::mutex> lk(g_prof_mu);
    double t = 0, w = 0;
    long long n = 0;
    for (auto& r : prof_recs()) {
      if (r.cls != cls) continue;
      RT_CHECK_CUDA(cudaEventSynchronize(r.e1));
      float f = 0;
      RT_CHECK_CUDA(cudaEventElapsedTime(&f, r.e0, r.e1));
      t += f;
      w += r.work;
      ++n;
    }
    *ms = t;
    *work = w;
    *count = n;
  });
}

int rt_euler_step(int dtype, const void* model_output, const void* sample, void* out, int64_t n, float sigma,
                  float sigma_next, void* stream) {
  return guarded([&] {
    RT_REQUIRE(model_output && sample && out && n >= 0, "euler_step: null argument");
    RT_REQUIRE((n * (int64_t)dtype_size(dtype)) % 16 == 0, "euler_step: n*sizeof(T) must be a multiple of 16");
    launch_euler_step(dtype, model_output, sample, out, n, sigma, sigma_next, (cudaStream_t)stream);
  });
}
int rt_cfg_combine(int dtype, const void* v2, void* out, int64_t n, float s, int zero_pred, void* stream) {
  return guarded([&] {
    RT_REQUIRE(v2 && out && n >= 0, "cfg_combine: null argument");
    launch_cfg_combine(dtype, v2, out, n, s, zero_pred, (cudaStream_t)stream);
  });
}
int rt_cfg_euler_step(int dtype, const void* v2, const void* sample, void* out, int64_t n, float s, int zero_pred,
                      float sigma, float sigma_next, void* stream) {
  return guarded([&] {
    RT_REQUIRE(v2 && sample && out && n >= 0, "cfg_euler_step: null argument");
    launch_cfg_euler(dtype, v2, sample, out, n, s, zero_pred, sigma, sigma_next, (cudaStream_t)stream);
  });
}
int rt_mask_scale_add(int dtype, const void* x, const void* mask, const void* acc_in, void* out, int batch, int rows,
                      int D, float scale, void* stream) {
  return guarded([&] {
    RT_REQUIRE(x && out && batch >= 0 && rows >= 0 && D >= 0, "mask_scale_add: bad argument");
    launch_mask_scale_add(dtype, x, mask, acc_in, out, batch, rows, D, scale, (cudaStream_t)stream);
  });
}
int rt_glyph_init_blend(int dtype, const void* noise, const void* glyph_latents, const unsigned char* mask, void* out,
                        int64_t n, float w_glyph, float w_noise, void* stream) {
  return guarded([&] {
    RT_REQUIRE(noise && glyph_latents && mask && out && n >= 0, "glyph_init_blend: null argument");
    launch_glyph_blend(dtype, noise, glyph_latents, mask, out, n, w_glyph, w_noise, (cudaStream_t)stream);
  });
}

int rt_gemm(const rt_gemm_launch* g, int impl, void* stream) {
  return guarded([&] {
    RT_REQUIRE(g, "gemm: null launch");
    RT_REQUIRE(g->nprob >= 1 && g->nprob <= 2, "gemm: nprob must be 1 or 2");
    for (int p = 0; p < g->nprob; ++p) {
      RT_REQUIRE(g->prob[p].nseg >= 1 && g->prob[p].nseg <= 4, "gemm: nseg must be 1..4");
      RT_REQUIRE(g->prob[p].A && g->prob[p].K > 0 && g->prob[p].m_rows >= 0, "gemm: bad problem");
    }
    cudaStream_t s = (cudaStream_t)stream;
    switch (impl) {
      case 0: launch_gemm(*g, s); break;
      case 1: launch_gemm_simt(*g, s); break;
      case 2: launch_gemm_tc(*g, s, 1); break;
      case 3: launch_gemm_tc(*g, s, 2); break;
      default: throw Error(RT_ERR_INVALID, "gemm: impl must be 0..3");
    }
  });
}

int rt_attention(const rt_attention_args* a, int impl, void* stream) {
  return guarded([&] {
    RT_REQUIRE(a && a->qkv && (a->out || a->sp_rows > 0), "attention: null argument");
    cudaStream_t s = (cudaStream_t)stream;
    if (impl == 0) launch_attention(*a, s);
    else if (impl == 1) launch_attention_simt(*a, s);
    else {
      std::string why;
      if (!attention_tc_supported(*a, &why)) throw Error(RT_ERR_UNSUPPORTED, "tcgen05 attention: " + why);
      launch_attention_tc(*a, s, impl - 2);
    }
  });
}

int rt_layernorm_modulate(int dtype, const void* x, int64_t x_bs, int x_ld, void* out, int64_t o_bs, int o_ld,
                          int batch, int D, int ngroups, const rt_lnmod_group* groups, void* stream) {
  return guarded([&] {
    RT_REQUIRE(x && out && groups, "layernorm_modulate: null argument");
    launch_ln_mod(dtype, x, x_bs, x_ld, out, o_bs, o_ld, batch, D, ngroups, groups, (cudaStream_t)stream);
  });
}

int rt_rope_table(const float* ids, int S, const int* axes_dims, float* out, void* stream) {
  return guarded([&] {
    RT_REQUIRE(ids && axes_dims && out && S >= 0, "rope_table: null argument");
    RT_REQUIRE(axes_dims[0] % 2 == 0 && axes_dims[1] % 2 == 0 && axes_dims[2] % 2 == 0, "rope axes must be even");
    launch_rope_table(ids, S, axes_dims, reinterpret_cast<float2*>(out), (cudaStream_t)stream);
  });
}

int rt_qknorm_rope(int dtype, void* buf, int64_t bs, int ld, int col0, int batch, int row0, int rows, int heads,
                   int hd, const void* norm_w, const float* rope, int rope_row0, void* stream) {
  return guarded([&] {
    RT_REQUIRE(buf && norm_w, "qknorm_rope: null argument");
    launch_qknorm_rope(dtype, buf, bs, ld, col0, batch, row0, rows, heads, hd, norm_w,
                       reinterpret_cast<const float2*>(rope), rope_row0, (cudaStream_t)stream);
  });
}

}  // extern "C"
