// kb_bind.cu — the bind fan-out list: which (task, node) pairs reach cache.Bind this cycle, in the order
// ssn.dispatch would have issued them (framework/session.go:277-314: inside ssn.Allocate, when ssn.JobReady, every task of the job
// in TaskStatusIndex[Allocated] is dispatched; the reference iterates a Go map there — the deterministic rule is Allocate order).
// Device side: compact the dispatched decisions into (key = dispatch_step << 32 | step, task) pairs, radix-sort them
// (cub::DeviceRadixSort), gather the nodes.  One call per cycle; the caller hands the list to a batched Binder instead of one
// goroutine + one API call per task (cache/cache.go:491-535).
#include <cuda_runtime.h>
#include <cub/device/device_radix_sort.cuh>

#include "kb_bind.h"

namespace kb {

__global__ void bind_compact_kernel(const kb_decision* __restrict__ dec, const uint32_t T, unsigned long long* __restrict__ keys,
                                    uint32_t* __restrict__ tasks, uint32_t* __restrict__ count) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  bool d = false;
  kb_decision x{};
  if (t < T) { x = dec[t]; d = x.dispatched != 0 && x.node >= 0; }
  const unsigned m = __ballot_sync(0xFFFFFFFFu, d);
  if (!m) return;
  uint32_t base = 0;
  if (lane == __ffs(m) - 1) base = atomicAdd(count, (uint32_t)__popc(m));          // one atomic per warp
  base = __shfl_sync(0xFFFFFFFFu, base, __ffs(m) - 1);
  if (d) {
    const uint32_t i = base + (uint32_t)__popc(m & ((1u << lane) - 1u));
    keys[i] = ((unsigned long long)x.dispatch_step << 32) | x.step;
    tasks[i] = t;
  }
}

__global__ void bind_gather_kernel(const kb_decision* __restrict__ dec, const uint32_t* __restrict__ tasks, const uint32_t n, int32_t* __restrict__ nodes) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) nodes[i] = dec[tasks[i]].node;
}

// scratch: device buffer of at least bind_scratch_bytes(T) bytes.  Outputs (host): task[n], node[n] in bind order; returns n via *n_out.
size_t bind_scratch_bytes(uint32_t T) {
  size_t tmp = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, tmp, (const unsigned long long*)nullptr, (unsigned long long*)nullptr, (const uint32_t*)nullptr, (uint32_t*)nullptr, (int)T);
  const size_t t1 = (size_t)(T ? T : 1);
  return ((tmp + 255) & ~(size_t)255) + t1 * (8 + 8 + 4 + 4 + 4) + 256;
}

cudaError_t bind_list(const kb_decision* d_dec, uint32_t T, unsigned char* scratch, size_t scratch_bytes, uint32_t* h_task, int32_t* h_node,
                      uint32_t* n_out, cudaStream_t st) {
  *n_out = 0;
  if (T == 0) return cudaSuccess;
  size_t tmp = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, tmp, (const unsigned long long*)nullptr, (unsigned long long*)nullptr, (const uint32_t*)nullptr, (uint32_t*)nullptr, (int)T);
  tmp = (tmp + 255) & ~(size_t)255;
  unsigned char* p = scratch;
  void* d_tmp = p; p += tmp;
  unsigned long long* k0 = (unsigned long long*)p; p += (size_t)T * 8;
  unsigned long long* k1 = (unsigned long long*)p; p += (size_t)T * 8;
  uint32_t* v0 = (uint32_t*)p; p += (size_t)T * 4;
  uint32_t* v1 = (uint32_t*)p; p += (size_t)T * 4;
  int32_t* nd = (int32_t*)p; p += (size_t)T * 4;
  uint32_t* cnt = (uint32_t*)p; p += 256;
  if ((size_t)(p - scratch) > scratch_bytes) return cudaErrorInvalidValue;
  cudaError_t c = cudaMemsetAsync(cnt, 0, 4, st);
  if (c != cudaSuccess) return c;
  bind_compact_kernel<<<(T + 255) / 256, 256, 0, st>>>(d_dec, T, k0, v0, cnt);
  uint32_t n = 0;
  c = cudaMemcpyAsync(&n, cnt, 4, cudaMemcpyDeviceToHost, st);
  if (c == cudaSuccess) c = cudaStreamSynchronize(st);
  if (c != cudaSuccess) return c;
  if (n == 0) return cudaSuccess;
  size_t tmp2 = tmp;
  c = cub::DeviceRadixSort::SortPairs(d_tmp, tmp2, k0, k1, v0, v1, (int)n, 0, 64, st);
  if (c != cudaSuccess) return c;
  bind_gather_kernel<<<(n + 255) / 256, 256, 0, st>>>(d_dec, v1, n, nd);
  c = cudaMemcpyAsync(h_task, v1, (size_t)n * 4, cudaMemcpyDeviceToHost, st);
  if (c == cudaSuccess) c = cudaMemcpyAsync(h_node, nd, (size_t)n * 4, cudaMemcpyDeviceToHost, st);
  if (c == cudaSuccess) c = cudaStreamSynchronize(st);
  if (c == cudaSuccess) c = cudaGetLastError();
  *n_out = n;
  return c;
}

}  // namespace kb
