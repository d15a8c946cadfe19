// Sequence-parallel plumbing (include/reptext_rt.h, rt_sp_group): the cross-GPU flag barrier that orders the
// peer stores of the fused QKV-GEMM / attention epilogues, and the CUDA-IPC helpers that make one rank's
// workspace addressable from the other ranks' processes.  The reference has no multi-GPU mode;
// the exchange pattern is the head <-> token re-partition of the joint attention
// (diffusers FluxTransformerBlock / FluxSingleTransformerBlock, reached from RepText/controlnet_flux.py:343-348).
#include <cstring>

#include "dtype_utils.cuh"
#include "rt_internal.h"
#include "sp_sync.cuh"
#include "sp_sync.cuh"

namespace rt {
namespace {

struct BarrierParams {
  unsigned long long* flags[RT_SP_MAX_RANKS];
  int world, rank;
};

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// One warp.  Every earlier kernel of this stream has completed (stream order), so its peer stores are performed;
// the system-scope fence + release stores publish them, the acquire loads on the other side make them visible
// to every later kernel of the waiting rank's stream.  Epochs only grow, so there is nothing to reset and a
// barrier can never be satisfied by a stale value.  Flag block: words 0..7 the peers' epochs, 8 this rank's epoch
// counter, 9 the ABORT word - sticky: once a wait timed out, every later barrier of every rank returns at once (the
// forward then finishes on garbage, quickly) until the host notices (rt_sp_status, once per step) and the group
// is re-synchronised (rt_sp_reset under a host barrier).
__global__ void sp_barrier_kernel(const BarrierParams p) {
  unsigned long long* mine = p.flags[p.rank];
  const int lane = threadIdx.x;
  // programmatic dependent launch: this kernel may be resident before its predecessor has finished; the peer
  // stores it publishes are only complete after this wait
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  unsigned long long epoch = 0, aborted = 0;
  if (lane == 0) {
    epoch = mine[8] + 1;
    mine[8] = epoch;
    aborted = ld_acquire_sys(mine + 9);  // sticky: set by an earlier time-out here, or published by a peer's
  }
  epoch = __shfl_sync(0xffffffffu, epoch, 0);
  aborted = __shfl_sync(0xffffffffu, aborted, 0);
  __threadfence_system();
  if (lane < p.world && lane != p.rank) st_release_sys(p.flags[lane] + p.rank, epoch);
  if (!aborted && lane < p.world && lane != p.rank) {
    const unsigned long long t0 = globaltimer_ns();
    unsigned spins = 0;
    while (ld_acquire_sys(mine + lane) < epoch) {
      __nanosleep(64);
      // a peer that gave up tells everybody (word 9 of every flag block), so the group fails within one time-out
      // instead of one time-out per barrier and rank
      if ((++spins & 255) == 0 && ld_acquire_sys(mine + 9)) break;
      if (globaltimer_ns() - t0 > 10000000000ull) {  // 10 s: a peer died; flag it instead of hanging the GPU
        for (int r = 0; r < p.world; ++r) st_release_sys(p.flags[r] + 9, 1ull);
        break;
      }
    }
  }
  __syncwarp();
  __threadfence_system();
}

}  // namespace

void launch_sp_barrier(const rt_sp_group& g, cudaStream_t stream) {
  RT_REQUIRE(g.world >= 2 && g.world <= RT_SP_MAX_RANKS && g.rank >= 0 && g.rank < g.world, "sp group: world / rank");
  BarrierParams p{};
  p.world = g.world;
  p.rank = g.rank;
  for (int i = 0; i < g.world; ++i) {
    RT_REQUIRE(g.peer_flags[i], "sp group: null flag pointer");
    p.flags[i] = g.peer_flags[i];
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(1);
  cfg.blockDim = dim3(32);
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = get_option("no_pdl") ? 0 : 1;
  RT_CHECK_CUDA(cudaLaunchKernelEx(&cfg, sp_barrier_kernel, p));
  count_launch();
}

// The in-kernel form's two steps as kernels, for launches that do not take the tcgen05 path (sp_sync.cuh): the barrier
// kernel draws the same epoch the announced word holds (word 8 + 1), the announcement is a one-thread kernel.
__global__ void sp_announce_kernel(unsigned long long* mine, int word) { mine[word] = mine[8] + 1; }

void launch_sp_sync_before(const SpSyncParams& s, cudaStream_t stream) {
  if (s.world == 0 || s.barrier_word == 0) return;
  rt_sp_group g{};
  g.world = s.world;
  g.rank = s.rank;
  for (int i = 0; i < s.world && i < RT_SP_MAX_RANKS; ++i) g.peer_flags[i] = s.flags[i];
  launch_sp_barrier(g, stream);
}
void launch_sp_sync_after(const SpSyncParams& s, cudaStream_t stream) {
  if (s.world == 0 || s.announce_word == 0) return;
  sp_announce_kernel<<<1, 1, 0, stream>>>(s.flags[s.rank], s.announce_word);
  RT_CHECK_CUDA(cudaGetLastError());
  count_launch();
}

}  // namespace rt

using namespace rt;

extern "C" {

int rt_sp_barrier(const rt_sp_group* g, void* stream) {
  return guarded([&] {
    RT_REQUIRE(g, "sp_barrier: null group");
    RT_REQUIRE(!g->lockstep, "sp_barrier: a lock-step group has no peers to wait for");
    launch_sp_barrier(*g, (cudaStream_t)stream);
  });
}

int rt_sp_status(const rt_sp_group* g, void* stream, int* timed_out) {
  return guarded([&] {
    RT_REQUIRE(g && timed_out && g->rank >= 0 && g->rank < RT_SP_MAX_RANKS && g->peer_flags[g->rank], "sp_status: bad argument");
    unsigned long long v = 0;
    RT_CHECK_CUDA(cudaMemcpyAsync(&v, g->peer_flags[g->rank] + 9, sizeof(v), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    RT_CHECK_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    *timed_out = v != 0;
  });
}

int rt_sp_reset(const rt_sp_group* g, void* stream) {
  // zero this rank's flag block (epochs, counter, abort word).  The caller brackets it with host barriers over
  // ALL ranks: nobody may be inside a forward, and nobody may start one before every rank has reset.
  return guarded([&] {
    RT_REQUIRE(g && g->rank >= 0 && g->rank < RT_SP_MAX_RANKS && g->peer_flags[g->rank], "sp_reset: bad argument");
    RT_CHECK_CUDA(cudaMemsetAsync(g->peer_flags[g->rank], 0, 16 * sizeof(unsigned long long), (cudaStream_t)stream));
    RT_CHECK_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  });
}

int rt_ipc_alloc(int64_t bytes, void** dev_ptr, unsigned char* handle64) {
  return guarded([&] {
    RT_REQUIRE(bytes > 0 && dev_ptr && handle64, "ipc_alloc: bad argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
    void* p = nullptr;
    RT_CHECK_CUDA(cudaMalloc(&p, (size_t)bytes));
    cudaError_t e = cudaMemset(p, 0, (size_t)bytes);
    cudaIpcMemHandle_t h;
    if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) {
      cudaFree(p);
      throw Error(RT_ERR_CUDA, std::string("ipc_alloc: ") + cudaGetErrorString(e));
    }
    memcpy(handle64, &h, 64);
    *dev_ptr = p;
  });
}

int rt_ipc_open(const unsigned char* handle64, void** dev_ptr) {
  return guarded([&] {
    RT_REQUIRE(handle64 && dev_ptr, "ipc_open: bad argument");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle64, 64);
    void* p = nullptr;
    RT_CHECK_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    *dev_ptr = p;
  });
}

int rt_ipc_close(void* dev_ptr) {
  return guarded([&] {
    if (dev_ptr) RT_CHECK_CUDA(cudaIpcCloseMemHandle(dev_ptr));
  });
}

int rt_ipc_free(void* dev_ptr) {
  return guarded([&] {
    if (dev_ptr) RT_CHECK_CUDA(cudaFree(dev_ptr));
  });
}

}  // extern "C"
