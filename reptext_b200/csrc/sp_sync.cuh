// Sequence-parallel phase barrier at the HEAD of the consuming kernel (rt_sp_group flag blocks, sp.cu).
//
// The stand-alone flag barrier (sp_barrier_kernel) is a kernel of its own between the kernel whose epilogue stores to
// the peers (QKV GEMM, attention) and the kernel that reads what the peers stored (attention, output projection): two
// kernel boundaries per hand-off instead of one, 124 hand-offs per step.  Here the same barrier runs inside the
// consuming kernel's TMA-producer warp, after griddepcontrol.wait (the producing grid has completed, so this rank's
// peer stores are performed - the guarantee the kernel form rests on as well) and before the first load of
// peer-written data:
//   * CTA 0 records the epoch in this rank's counter (word 8, so that the kernel form stays in step) and release-stores
//     it into every peer's flag block;
//   * every CTA acquire-spins until each peer's flag has reached the epoch; the loads that follow go through the async
//     proxy (TMA), hence the proxy fence.
// Every CTA of the kernel needs the same target epoch, and word 8 changes while they read it; so the target is ANNOUNCED
// one kernel earlier: CTA 0 of the kernel before (the QKV GEMM for the pre-attention barrier, the attention kernel for
// the post-attention one) writes `epoch + 1` into a spare word of the flag block (11 / 12, alternating so that a kernel
// that both reads its target and announces the next one does not overwrite what its own CTAs are reading).  That word
// was written by a grid that has completed when the barrier kernel's CTAs read it.  Purely device-side (CUDA-graph
// safe), same flag words and epochs as the kernel form - the two mix freely in one stream - and the same failure
// behaviour (10 s time-out -> sticky ABORT word in every rank's block).  A tried alternative - the PRODUCING kernel's last
// CTA publishing the epoch - needs a system-scope fence at the end of every CTA and cost more than it saved (98.2 vs
// 97.5 ms per cfg-5 step on two GPUs).
// Flag block: words 0..7 peers' epochs | 8 this rank's epoch counter | 9 ABORT | 11, 12 announced targets.
#pragma once
#include <cstdint>

#include "../../include/reptext_rt.h"

namespace rt {

struct SpSyncParams {  // by value inside the kernels' parameter blocks; world == 0: nothing to do
  unsigned long long* flags[RT_SP_MAX_RANKS];
  int world, rank;
  int barrier_word;   // 11 / 12: run the barrier at the head of this kernel, the target epoch is in that word; 0: none
  int announce_word;  // 11 / 12: CTA 0 writes the NEXT barrier's target there; 0: none
};

namespace spsync {

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
// the counter / announcement words: L2-coherent accesses (a line of the flag block may sit in an L1 from an earlier kernel)
__device__ __forceinline__ unsigned long long ld_gpu(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_gpu(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// Call from ONE whole warp of every CTA, after griddepcontrol.wait.
__device__ __forceinline__ void sp_barrier_head(const SpSyncParams& p, int lane, bool cta0) {
  if (p.world == 0) return;  // kernel-uniform
  unsigned long long* mine = p.flags[p.rank];
  if (p.barrier_word == 0) {
    if (cta0 && lane == 0 && p.announce_word) st_gpu(mine + p.announce_word, ld_gpu(mine + 8) + 1);
    return;
  }
  const unsigned long long epoch = ld_gpu(mine + p.barrier_word);  // announced by a grid that has completed
  if (cta0) {
    if (lane == 0) {
      st_gpu(mine + 8, epoch);
      if (p.announce_word) st_gpu(mine + p.announce_word, epoch + 1);
    }
    __threadfence_system();
    if (lane < p.world && lane != p.rank) st_release_sys(p.flags[lane] + p.rank, epoch);
  }
  if (lane < p.world && lane != p.rank && !ld_acquire_sys(mine + 9)) {
    const unsigned long long t0 = globaltimer_ns();
    unsigned spins = 0;
    while (ld_acquire_sys(mine + lane) < epoch) {
      __nanosleep(32);
      if ((++spins & 255) == 0) {
        if (ld_acquire_sys(mine + 9)) break;  // a peer gave up: the group fails within one time-out
        if (globaltimer_ns() - t0 > 10000000000ull) {  // 10 s: a peer died; flag it instead of hanging the GPU
          for (int r = 0; r < p.world; ++r) st_release_sys(p.flags[r] + 9, 1ull);
          break;
        }
      }
    }
  }
  __syncwarp();
  asm volatile("fence.proxy.async.global;" ::: "memory");
}

}  // namespace spsync
}  // namespace rt
