// kb_evict_kernels.cu — reclaim / preempt on the device (kb_evict.h holds the algorithm, shared with the CPU emulation).
//
// evict_kernel<PREEMPT>   cooperative grid of up to one CTA per SM.  Per preemptor task ALL CTAs sweep the node table together
//                         (one node per thread per iteration: K1 predicate, K2 score for preempt, the serial victim walk of
//                         the node), a grid-wide arg-max picks the node, thread 0 of the grid commits (evictions, Pipeline,
//                         Statement log) and runs the action's control flow; grid barriers in between.
//
// This translation unit is compiled with -Xptxas -dlcm=cg: every global load goes to L2.  The action mutates node records,
// job / queue accounting and the Running tasks' states from ONE thread while 147 other SMs read them in the next sweep; with
// L1-cached loads those SMs could see stale lines.  (The allocate kernels keep L1 caching: their mutable tables are only
// read by the SM that writes them.)
#include <cuda_runtime.h>

#include "kb_evict.h"
#include "kb_evict_launch.h"

namespace kb {

constexpr int EVICT_THREADS = 512;

__device__ __forceinline__ uint32_t ev_ld_acquire(const uint32_t* p) {
  uint32_t v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v;
}

struct GridExec {
  EvictCtl* g;               // barrier / broadcast / arg-max slots + the preemptor and its class
  uint64_t* red_smem;        // [32] per-CTA reduction scratch
  __device__ __forceinline__ int tid() const { return (int)(blockIdx.x * blockDim.x + threadIdx.x); }
  __device__ __forceinline__ int nthreads() const { return (int)(gridDim.x * blockDim.x); }
  // sense-reversing grid barrier (all CTAs are co-resident: cooperative launch)
  __device__ __forceinline__ void sync() {
    __syncthreads();
    if (threadIdx.x == 0 && gridDim.x > 1) {
      const uint32_t gen = ev_ld_acquire(&g->bar_gen);
      __threadfence();
      if (atomicAdd(&g->bar_count, 1u) == gridDim.x - 1) {
        g->bar_count = 0;
        __threadfence();
        atomicAdd(&g->bar_gen, 1u);
      } else {
        while (ev_ld_acquire(&g->bar_gen) == gen) __nanosleep(20);
      }
    }
    __syncthreads();
  }
  __device__ __forceinline__ uint32_t bcast(uint32_t v) {
    if (tid() == 0) g->bc = v;
    sync();
    const uint32_t r = *((volatile uint32_t*)&g->bc);
    sync();
    return r;
  }
  // arg-max over the grid: thread 0 cleared g->red before the barrier that precedes the sweep (try_preemptor)
  __device__ __forceinline__ uint64_t block_max(uint64_t v) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned hi = (unsigned)(v >> 32), lo = (unsigned)v;
    const unsigned mhi = __reduce_max_sync(0xFFFFFFFFu, hi);
    const unsigned mlo = __reduce_max_sync(0xFFFFFFFFu, hi == mhi ? lo : 0u);
    if (lane == 0) red_smem[warp] = ((uint64_t)mhi << 32) | mlo;
    __syncthreads();
    if (warp == 0) {
      uint64_t r = lane < (int)(blockDim.x >> 5) ? red_smem[lane] : 0ull;
      const unsigned rh = (unsigned)(r >> 32), rl = (unsigned)r;
      const unsigned xh = __reduce_max_sync(0xFFFFFFFFu, rh);
      const unsigned xl = __reduce_max_sync(0xFFFFFFFFu, rh == xh ? rl : 0u);
      if (lane == 0 && (xh | xl)) atomicMax(&g->red, ((unsigned long long)xh << 32) | xl);
    }
    sync();
    const uint64_t r = *((volatile unsigned long long*)&g->red);
    return r;
  }
  __device__ __forceinline__ void clear_max() { g->red = 0ull; }
  __device__ __forceinline__ ClassRec& cls() { return g->cls; }
  __device__ __forceinline__ Preemptor& pre() { return g->pre; }
};

template <int PREEMPT>
__global__ void __launch_bounds__(EVICT_THREADS, 1)
evict_kernel(const __grid_constant__ DevSession S, const __grid_constant__ EvictDev E) {
  __shared__ uint64_t s_red[32];
  GridExec x{E.ctl, s_red};
  if (PREEMPT) run_preempt(x, S, E);
  else run_reclaim(x, S, E);
}

cudaError_t launch_evict(const bool preempt, const DevSession& S, const EvictDev& E, const int sm_count, cudaStream_t stream) {
  // enough CTAs for one node per thread per sweep iteration, at most one CTA per SM (co-residency of the grid barrier)
  int grid = (int)((S.N + EVICT_THREADS - 1) / EVICT_THREADS);
  grid = grid < 1 ? 1 : (grid > sm_count ? sm_count : grid);
  DevSession s = S; EvictDev ev = E;
  void* args[] = {(void*)&s, (void*)&ev};
  const void* fn = preempt ? (const void*)evict_kernel<1> : (const void*)evict_kernel<0>;
  return cudaLaunchCooperativeKernel(fn, dim3(grid), dim3(EVICT_THREADS), args, 0, stream);
}

}  // namespace kb
