// Micro-benchmark of the replay step loop's building blocks on one warp (cycles per iteration, clock64).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o step_loop step_loop.cu && ./step_loop
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define FULL 0xFFFFFFFFu
__device__ __forceinline__ uint64_t warp_max_u64(uint64_t v) {
  const unsigned hi = (unsigned)(v >> 32), lo = (unsigned)v;
  const unsigned mhi = __reduce_max_sync(FULL, hi);
  const unsigned mlo = __reduce_max_sync(FULL, hi == mhi ? lo : 0u);
  return ((uint64_t)mhi << 32) | (uint64_t)mlo;
}
__device__ __forceinline__ uint64_t warp_max_shfl(uint64_t v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { const uint64_t x = __shfl_xor_sync(FULL, v, o); v = x > v ? x : v; }
  return v;
}
template <int VARIANT>
__global__ void k(uint64_t* out, long long* cyc, int iters, int spin_warps) {
  __shared__ uint64_t chain[8][32];
  __shared__ uint32_t fi[8];
  __shared__ uint32_t rec[32];
  __shared__ volatile uint32_t flag;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) flag = 0;
  for (int d = 0; d < 8; ++d) { chain[d][lane] = ((uint64_t)(1000 - d * 7 - (lane * 13) % 50) << 32) | (0xFFFFFFFFu - lane); if (lane == 0) fi[d] = 0xFFFF00FFu; }
  __syncthreads();
  if (warp != 0) {               // other warps of the CTA: spin like the pipeline's helper warps do
    if (warp <= spin_warps) while (!flag) __nanosleep(40);
    return;
  }
  uint64_t cur = chain[0][lane];
  uint32_t depth = 0, placed = 0, n_alloc = 0;
  bool cur_fi = true;
  uint64_t acc = 0;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    uint64_t best;
    if (VARIANT == 5) best = warp_max_shfl(cur); else best = warp_max_u64(cur);
    if (VARIANT == 0) { acc += best; cur = cur - (lane == (it & 31) ? (1ull << 32) : 0); continue; }
    const uint32_t owner = (uint32_t)__ffs(__ballot_sync(FULL, cur == best)) - 1u;
    if (VARIANT == 1) { acc += owner; if ((uint32_t)lane == owner) cur -= (1ull << 32); continue; }
    const bool fits = ((__ballot_sync(FULL, cur_fi) >> owner) & 1u) != 0;
    const bool own = (uint32_t)lane == owner;
    if (VARIANT >= 3 && __any_sync(FULL, own && depth + 1 >= 8)) { depth = 0; }
    if (own) { depth = (depth + 1) & 7; cur = chain[depth][lane] - ((uint64_t)it << 32); cur_fi = (fi[depth] >> lane) & 1u; }
    if (VARIANT >= 4 && lane == 0) rec[placed & 31] = owner | (fits ? 32u : 0u);
    placed += 1; n_alloc += fits ? 1u : 0u;
    acc += n_alloc;
  }
  const long long t1 = clock64();
  if (lane == 0) { out[0] = acc + cur + placed; cyc[0] = t1 - t0; flag = 1; }
}
int main() {
  uint64_t* out; long long* cyc; cudaMalloc(&out, 8); cudaMalloc(&cyc, 8);
  const int iters = 4096;
  const char* names[] = {"2xREDUX + update", "+ ballot/ffs owner", "+ fi ballot + chain LDS", "+ any_sync(extension)", "+ lane0 smem record", "shfl-tree max instead of REDUX (full)"};
  for (int spin = 0; spin <= 14; spin += 14)
    for (int v = 0; v < 6; ++v) {
      long long h = 0;
      for (int rep = 0; rep < 2; ++rep) {
        switch (v) {
          case 0: k<0><<<1, 512>>>(out, cyc, iters, spin); break; case 1: k<1><<<1, 512>>>(out, cyc, iters, spin); break;
          case 2: k<2><<<1, 512>>>(out, cyc, iters, spin); break; case 3: k<3><<<1, 512>>>(out, cyc, iters, spin); break;
          case 4: k<4><<<1, 512>>>(out, cyc, iters, spin); break; default: k<5><<<1, 512>>>(out, cyc, iters, spin); break;
        }
        cudaDeviceSynchronize();
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
      }
      printf("spin_warps %2d  variant %d  %-42s %7.1f cycles/iter\n", spin, v, names[v], (double)h / iters);
    }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
