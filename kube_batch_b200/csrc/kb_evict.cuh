// kb_evict.cuh — reclaim / preempt on the device: ONE CTA runs the whole action (kb_evict.h holds the algorithm, shared
// with the CPU emulation).  Per preemptor task the CTA's threads sweep the node table (one node per thread per iteration:
// K1 predicate, K2 score for preempt, the serial victim walk of the node), a block arg-max picks the node, thread 0 commits.
#pragma once

#include "kb_evict.h"
#include "kb_kernels.cuh"

namespace kb {

constexpr int EVICT_THREADS = 1024;

struct GpuExec {
  ClassRec* c; Preemptor* p; uint32_t* bc; uint64_t* red;
  __device__ __forceinline__ int tid() const { return (int)threadIdx.x; }
  __device__ __forceinline__ int nthreads() const { return (int)blockDim.x; }
  __device__ __forceinline__ void sync() { __syncthreads(); }
  __device__ __forceinline__ uint32_t bcast(uint32_t v) {
    if (threadIdx.x == 0) *bc = v;
    __syncthreads();
    const uint32_t r = *bc;
    __syncthreads();
    return r;
  }
  __device__ __forceinline__ uint64_t block_max(uint64_t v) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    v = warp_max_u64(v);
    if (lane == 0) red[warp] = v;
    __syncthreads();
    uint64_t r = lane < (int)(blockDim.x >> 5) ? red[lane] : 0ull;
    r = warp_max_u64(r);
    __syncthreads();
    return r;
  }
  __device__ __forceinline__ ClassRec& cls() { return *c; }
  __device__ __forceinline__ Preemptor& pre() { return *p; }
};

template <int PREEMPT>
__global__ void __launch_bounds__(EVICT_THREADS, 1)
evict_kernel(const __grid_constant__ DevSession S, const __grid_constant__ EvictDev E) {
  __shared__ ClassRec s_cls;
  __shared__ Preemptor s_pre;
  __shared__ uint32_t s_bc;
  __shared__ uint64_t s_red[32];
  GpuExec x{&s_cls, &s_pre, &s_bc, s_red};
  if (PREEMPT) run_preempt(x, S, E);
  else run_reclaim(x, S, E);
}

}  // namespace kb
