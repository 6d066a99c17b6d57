// kb_kernels.cuh — sm_100a kernels of the allocate cycle.
//
//   visit_kernel        one launch = one SCAN of the node table for the class of the next run
//                       (K1 predicate bitmask + K2 fused score, node tiles staged into shared memory
//                       by TMA bulk copies) -> per-CTA top-KTOP candidate keys; the LAST CTA to finish
//                       (ticket) merges them (K3), then replays as many runs of that class as it can
//                       certify exactly (dirty-node re-evaluation, AddTask bookkeeping, gang stop rule)
//                       and runs the control plane (kb_ctl.h) to pick the next visit.  The host only
//                       pumps launches until Ctl.done — no host round trip inside the cycle.
//   gang_commit_kernel  K4: per-PodGroup inclusive prefix scan over the Allocated flags in processing
//                       order -> dispatched bit + dispatch step (framework/session.go:277-285).
//   matrix_kernel       full fit / score matrix for a task range (debug / parity, kb_predicate_score).
//   best_nodes_kernel   K1+K2+K3 over a task range x all nodes in ONE launch: per-task argmax via
//                       warp-shuffle max + one 64-bit atomicMax per warp (kb_best_nodes).
//
// No tensor cores anywhere: this is integer / FP64-compare work on an L2-resident table.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include "kb_ctl.h"

namespace kb {

constexpr int SCAN_THREADS = TILE_NODES;           // 128: one node per thread per tile
constexpr int MAXCOLS = 2 * KB_MAX_R + 6 + 3 * KB_MAX_W;   // 34
constexpr unsigned FULL = 0xFFFFFFFFu;

// ---------------------------------------------------------------------------------------------
// TMA (1-D bulk copy) + mbarrier helpers — SASS: UBLKCP / SYNCS
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  } while (!ok);
}

__device__ __forceinline__ uint64_t warp_max_u64(uint64_t v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    uint64_t t = __shfl_xor_sync(FULL, v, o);
    v = t > v ? t : v;
  }
  return v;
}

// Strided column accessor (shared-memory tile: stride TILE_NODES; dirty slots: stride DMAX)
struct ColAcc {
  const uint64_t* base; uint32_t i, stride, R, W;
  __device__ __forceinline__ uint64_t col(uint32_t c) const { return base[c * stride + i]; }
  __device__ __forceinline__ double idle(uint32_t r) const { return u64_as_double(col(col_idle(R, r))); }
  __device__ __forceinline__ double rel(uint32_t r) const { return u64_as_double(col(col_rel(R, r))); }
  __device__ __forceinline__ int64_t alloc_cpu() const { return (int64_t)col(col_alloc_cpu(R)); }
  __device__ __forceinline__ int64_t alloc_mem() const { return (int64_t)col(col_alloc_mem(R)); }
  __device__ __forceinline__ int64_t nz_cpu() const { return (int64_t)col(col_nz_cpu(R)); }
  __device__ __forceinline__ int64_t nz_mem() const { return (int64_t)col(col_nz_mem(R)); }
  __device__ __forceinline__ int32_t pods() const { return (int32_t)(uint32_t)(col(col_pods(R)) & 0xFFFFFFFFull); }
  __device__ __forceinline__ int32_t max_pods() const { return (int32_t)(uint32_t)(col(col_pods(R)) >> 32); }
  __device__ __forceinline__ uint32_t flags() const { return (uint32_t)col(col_flags(R)); }
  __device__ __forceinline__ uint64_t labels(uint32_t w) const { return col(col_labels(R, W, w)); }
  __device__ __forceinline__ uint64_t taints(uint32_t w) const { return col(col_taints(R, W, w)); }
  __device__ __forceinline__ uint64_t ports(uint32_t w) const { return col(col_ports(R, W, w)); }
};

// ---------------------------------------------------------------------------------------------
// CTA-wide exact top-KTOP merge by rank counting.
//   keys[0..KTOP)            current list, descending, 0-padded
//   keys[KTOP..KTOP+128)     128 fresh keys (one per thread, 0 = none)
// Non-zero keys are unique (distinct node indices), so ranks are unique.  All SCAN_THREADS call it.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void cta_topk_merge(uint64_t* keys, uint64_t* newl, uint64_t fresh, int tid) {
  keys[KTOP + tid] = fresh;
  if (tid < KTOP) newl[tid] = 0;
  __syncthreads();
  const bool beats = fresh > keys[KTOP - 1];
  if (!__syncthreads_or(beats ? 1 : 0)) return;        // nothing can enter the list
  if (fresh != 0) {
    int r = 0;
#pragma unroll 8
    for (int i = 0; i < KTOP + SCAN_THREADS; ++i) r += keys[i] > fresh ? 1 : 0;
    if (r < KTOP) newl[r] = fresh;
  }
  if (tid < KTOP) {
    const uint64_t old = keys[tid];
    if (old != 0) {
      int r = 0;
#pragma unroll 8
      for (int i = 0; i < KTOP + SCAN_THREADS; ++i) r += keys[i] > old ? 1 : 0;
      if (r < KTOP) newl[r] = old;
    }
  }
  __syncthreads();
  if (tid < KTOP) keys[tid] = newl[tid];
  __syncthreads();
}

struct DirtySlots {
  uint64_t col[MAXCOLS][DMAX];      // same column scheme as a tile, stride DMAX
  double   used_add[KB_MAX_R][DMAX];
  uint32_t node[DMAX];
};

struct VisitSmem {
  ClassRec cls;
  Ctl ctl;
  uint64_t keys[KTOP + SCAN_THREADS];
  uint64_t newl[KTOP];
  DirtySlots dirty;
  uint64_t mbar[2];
  uint32_t is_last;
};

// ---------------------------------------------------------------------------------------------
// visit_kernel
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(SCAN_THREADS)
visit_kernel(const __grid_constant__ DevSession S) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  // layout: [VisitSmem][pad to 128][tile buffer 0][tile buffer 1]
  VisitSmem& sm = *reinterpret_cast<VisitSmem*>(smem_raw);
  const uint32_t tile_u64 = S.ncols * TILE_NODES;
  const uint32_t tile_bytes = tile_u64 * 8u;
  uint64_t* tilebuf = reinterpret_cast<uint64_t*>(smem_raw + ((sizeof(VisitSmem) + 127) / 128) * 128);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  Ctl* gctl = S.ctl;
  if (*((volatile uint32_t*)&gctl->done)) return;
  const uint32_t cls_id = *((volatile uint32_t*)&gctl->cur_class);

  // class record -> shared memory (broadcast reads afterwards)
  {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(&S.classes[cls_id]);
    uint32_t* dst = reinterpret_cast<uint32_t*>(&sm.cls);
    for (uint32_t i = tid; i < sizeof(ClassRec) / 4; i += SCAN_THREADS) dst[i] = src[i];
  }
  if (tid == 0) { mbar_init(&sm.mbar[0], 1); mbar_init(&sm.mbar[1], 1); fence_mbar_init(); }
  if (tid < KTOP) sm.keys[tid] = 0;
  __syncthreads();

  // ---------------- scan: tiles blockIdx.x, +gridDim.x, ... double-buffered TMA ----------------
  const uint32_t first = blockIdx.x, stride = gridDim.x;
  uint32_t n_local = first < S.NT ? (S.NT - first + stride - 1) / stride : 0;
  if (tid == 0 && n_local > 0) {
    mbar_expect_tx(&sm.mbar[0], tile_bytes);
    tma_load_1d(tilebuf, S.tiles + (size_t)first * tile_u64, tile_bytes, &sm.mbar[0]);
  }
  for (uint32_t it = 0; it < n_local; ++it) {
    const uint32_t b = it & 1u;
    if (tid == 0 && it + 1 < n_local) {      // prefetch next tile into the other buffer (freed by the
      const uint32_t nb = b ^ 1u;            // __syncthreads at the end of the previous iteration)
      mbar_expect_tx(&sm.mbar[nb], tile_bytes);
      tma_load_1d(tilebuf + (size_t)nb * tile_u64, S.tiles + (size_t)(first + (it + 1) * stride) * tile_u64, tile_bytes, &sm.mbar[nb]);
    }
    mbar_wait(&sm.mbar[b], (it >> 1) & 1u);
    const uint32_t t = first + it * stride;
    const uint32_t node = t * TILE_NODES + tid;
    uint64_t key = 0;
    if (node < S.N) {
      ColAcc acc{tilebuf + (size_t)b * tile_u64, (uint32_t)tid, TILE_NODES, S.cf.R, S.cf.W};
      key = eval_pair(S.cf, sm.cls, acc, node, nullptr);
    }
    cta_topk_merge(sm.keys, sm.newl, key, tid);       // ends with __syncthreads: buffer b is free again
  }
  // publish this CTA's list
  if (tid < KTOP) S.cand[(size_t)blockIdx.x * KTOP + tid] = sm.keys[tid];
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    const uint32_t ticket = atomicAdd(&gctl->arrive, 1u);
    sm.is_last = (ticket == gridDim.x - 1) ? 1u : 0u;
  }
  __syncthreads();
  if (!sm.is_last) return;
  __threadfence();

  // ---------------- K3: merge the per-CTA lists (all 128 threads) ----------------
  if (gridDim.x > 1) {
    if (tid < KTOP) sm.keys[tid] = 0;
    __syncthreads();
    const uint32_t total = gridDim.x * KTOP;
    for (uint32_t base = 0; base < total; base += SCAN_THREADS) {
      const uint32_t i = base + tid;
      const uint64_t k = i < total ? __ldcg(&S.cand[i]) : 0ull;
      cta_topk_merge(sm.keys, sm.newl, k, tid);
    }
  }
  if (tid == 0) sm.ctl = *gctl;
  __syncthreads();
  if (warp != 0) return;

  // ---------------- exact replay + control: warp 0 only ----------------
  Ctl& c = sm.ctl;
  const uint32_t R = S.cf.R, W = S.cf.W, ncols = S.ncols;
  uint32_t nd = 0;        // dirty nodes
  uint32_t p = 0;         // first list entry not known to be dirty
  if (lane == 0) { c.scans += 1; c.pairs_scanned += (unsigned long long)S.N; }
  __syncwarp();

  for (;;) {              // runs
    if (c.done || c.cur_class != cls_id) break;
    const uint32_t j = (uint32_t)c.cur_job;
    const uint32_t jend = S.job_ord_off[j + 1];
    uint32_t run_left = c.cur_run;
    uint32_t placed = 0;
    uint32_t reason = STOP_RUN_DONE;
    while (run_left > 0) {      // steps: one pending task each
      if (nd == DMAX) { reason = STOP_RESCAN; break; }
      // advance p over entries that became dirty
      uint64_t clean_key = 0;
      while (p < (uint32_t)KTOP) {
        clean_key = sm.keys[p];
        if (clean_key == 0) break;
        const uint32_t pn = key_node(clean_key);
        const bool m = lane < nd && sm.dirty.node[lane] == pn;
        if (!__any_sync(FULL, m)) break;
        ++p;
      }
      if (p == (uint32_t)KTOP) { reason = STOP_RESCAN; break; }   // list exhausted while full: cannot certify
      // pop the task (allocate.go:130); lane 0 owns the job arrays
      const uint32_t pos = __shfl_sync(FULL, lane == 0 ? S.job_pos[j] : 0u, 0);
      // exact re-evaluation of every dirty node against its current state
      uint64_t my = 0;
      bool fi = false;
      if (lane < nd) {
        ColAcc acc{&sm.dirty.col[0][0], (uint32_t)lane, DMAX, R, W};
        my = eval_pair(S.cf, sm.cls, acc, sm.dirty.node[lane], &fi);
      }
      uint64_t best = warp_max_u64(my);
      best = clean_key > best ? clean_key : best;
      if (lane == 0) {
        S.job_pos[j] = pos + 1;
        c.tasks_processed += 1;
        c.pairs_logical += (unsigned long long)S.N;
        c.pairs_replayed += (unsigned long long)nd;
      }
      run_left -= 1;
      if (best == 0) { reason = STOP_NOFIT; break; }              // allocate.go:144-148
      const uint32_t bn = key_node(best);
      const unsigned hit = __ballot_sync(FULL, lane < nd && sm.dirty.node[lane] == bn);
      uint32_t slot;
      bool fits_idle;
      if (hit) {
        slot = (uint32_t)__ffs(hit) - 1u;
        fits_idle = __shfl_sync(FULL, fi ? 1 : 0, slot) != 0;
      } else {
        slot = nd;
        const uint64_t* gt = S.tiles + (size_t)(bn / TILE_NODES) * (ncols * TILE_NODES) + (bn % TILE_NODES);
        for (uint32_t cc = lane; cc < ncols; cc += 32) sm.dirty.col[cc][slot] = __ldcg(gt + (size_t)cc * TILE_NODES);
        if (lane < KB_MAX_R) sm.dirty.used_add[lane][slot] = 0.0;
        if (lane == 0) sm.dirty.node[slot] = bn;
        nd += 1;
        __syncwarp();
        ColAcc acc{&sm.dirty.col[0][0], slot, DMAX, R, W};
        fits_idle = res_less_equal(R, [&](uint32_t k) { return sm.cls.initreq[k]; }, [&](uint32_t k) { return acc.idle(k); });
      }
      // commit: ssn.Allocate (session.go:235) or ssn.Pipeline (session.go:194) -> NodeInfo.AddTask (node_info.go:172-212)
      if (lane == 0) {
        const uint32_t base_col = fits_idle ? col_idle(R, 0) : col_rel(R, 0);
        for (uint32_t k = 0; k < R; ++k) {
          const double cur = u64_as_double(sm.dirty.col[base_col + k][slot]);
          sm.dirty.col[base_col + k][slot] = double_as_u64(KB_DSUB(cur, sm.cls.resreq[k]));
          sm.dirty.used_add[k][slot] = KB_DADD(sm.dirty.used_add[k][slot], sm.cls.resreq[k]);
        }
        sm.dirty.col[col_nz_cpu(R)][slot] = (uint64_t)((int64_t)sm.dirty.col[col_nz_cpu(R)][slot] + sm.cls.nz_cpu);
        sm.dirty.col[col_nz_mem(R)][slot] = (uint64_t)((int64_t)sm.dirty.col[col_nz_mem(R)][slot] + sm.cls.nz_mem);
        sm.dirty.col[col_pods(R)][slot] += 1ull;                   // pods live in the low 32 bits
        for (uint32_t w = 0; w < W; ++w) sm.dirty.col[col_ports(R, W, w)][slot] |= sm.cls.port_own[w];
        kb_decision d;
        d.node = (int32_t)bn;
        d.kind = fits_idle ? KB_KIND_ALLOCATED : KB_KIND_PIPELINED;
        d.dispatched = 0; d.reserved = 0;
        d.step = c.step;
        d.dispatch_step = 0xFFFFFFFFu;
        S.dec[S.ord_task[pos]] = d;
        c.step += 1;
        if (fits_idle) { c.tasks_allocated += 1; S.job_ready[j] += 1; } else c.tasks_pipelined += 1;
        S.job_placed[j] += 1;
        on_allocate_event(S, j, sm.cls);
      }
      placed += 1;
      __syncwarp();
      // allocate.go:185-188: a ready job yields after every task while tasks remain
      const bool yield = __shfl_sync(FULL, (lane == 0 && ssn_job_ready(S, j) && (pos + 1 < jend)) ? 1 : 0, 0) != 0;
      if (yield) { reason = STOP_YIELD; break; }
    }
    if (lane == 0) {
      if (reason == STOP_RESCAN) c.rescans += 1;
      after_run(S, c, reason, placed);
    }
    __syncwarp();
    if (reason == STOP_RESCAN) break;
  }

  // write the dirty nodes back to the global table
  for (uint32_t s = 0; s < nd; ++s) {
    const uint32_t n = sm.dirty.node[s];
    uint64_t* gt = S.tiles + (size_t)(n / TILE_NODES) * (ncols * TILE_NODES) + (n % TILE_NODES);
    for (uint32_t cc = lane; cc < ncols; cc += 32) gt[(size_t)cc * TILE_NODES] = sm.dirty.col[cc][s];
    if (lane < R) S.node_used[(size_t)lane * S.N + n] = KB_DADD(S.node_used[(size_t)lane * S.N + n], sm.dirty.used_add[lane][s]);
  }
  __syncwarp();
  if (lane == 0) { c.arrive = 0; *gctl = c; }
}

// ---------------------------------------------------------------------------------------------
// K4: gang commit.  One warp per job: inclusive prefix scan of the Allocated flags over the job's
// tasks in processing order; e* = first Allocate at which ReadyTaskNum >= MinAvailable (always, when
// gang's JobReadyFn is not enabled); a task allocated at position i is dispatched at step[max(i, e*)].
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
gang_commit_kernel(const __grid_constant__ DevSession S, const int32_t* __restrict__ job_ready0) {
  const uint32_t warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp_global >= S.J) return;
  const uint32_t j = warp_global;
  const uint32_t lo = S.job_ord_off[j], hi = S.job_pos[j];    // processed slots
  if (lo >= hi) return;
  const int32_t need = S.gang_ready ? S.job_min_avail[j] - job_ready0[j] : 0;   // allocations required before JobReady
  // pass 1: find e* (slot index) and its step
  uint32_t estar = 0xFFFFFFFFu, estep = 0;
  int32_t carried = 0;
  for (uint32_t base = lo; base < hi && estar == 0xFFFFFFFFu; base += 32) {
    const uint32_t i = base + lane;
    uint32_t alloc = 0, step = 0;
    if (i < hi) { const kb_decision d = S.dec[S.ord_task[i]]; alloc = d.kind == KB_KIND_ALLOCATED; step = d.step; }
    int32_t x = (int32_t)alloc;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int32_t y = __shfl_up_sync(FULL, x, o); if (lane >= o) x += y; }
    const int32_t incl = carried + x;
    const unsigned m = __ballot_sync(FULL, alloc && incl >= need);
    if (m) {
      const int src = __ffs(m) - 1;
      estar = base + src;
      estep = __shfl_sync(FULL, step, src);
    }
    carried = __shfl_sync(FULL, incl, 31);
  }
  if (estar == 0xFFFFFFFFu) return;                   // never became ready: nothing is dispatched
  for (uint32_t i = lo + lane; i < hi; i += 32) {
    const uint32_t t = S.ord_task[i];
    kb_decision d = S.dec[t];
    if (d.kind != KB_KIND_ALLOCATED) continue;
    d.dispatched = 1;
    d.dispatch_step = i <= estar ? estep : d.step;
    S.dec[t] = d;
  }
}

// ---------------------------------------------------------------------------------------------
// Full matrix for tasks [task_lo, task_hi) x all nodes (parity / debug).
// grid = (NT, task chunks); each CTA stages one node tile with TMA and walks its task chunk.
// ---------------------------------------------------------------------------------------------
constexpr int MATRIX_TASKS_PER_CTA = 32;

__global__ void __launch_bounds__(SCAN_THREADS)
matrix_kernel(const __grid_constant__ DevSession S, const uint32_t* __restrict__ task_class, uint32_t task_lo, uint32_t task_hi,
              uint8_t* __restrict__ fit, double* __restrict__ score) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ ClassRec cls;
  __shared__ uint64_t mbar;
  uint64_t* tilebuf = reinterpret_cast<uint64_t*>(smem_raw);
  const int tid = threadIdx.x;
  const uint32_t tile_u64 = S.ncols * TILE_NODES;
  if (tid == 0) {
    mbar_init(&mbar, 1); fence_mbar_init();
    mbar_expect_tx(&mbar, tile_u64 * 8u);
    tma_load_1d(tilebuf, S.tiles + (size_t)blockIdx.x * tile_u64, tile_u64 * 8u, &mbar);
  }
  __syncthreads();
  mbar_wait(&mbar, 0);
  const uint32_t node = blockIdx.x * TILE_NODES + tid;
  const uint32_t t0 = task_lo + blockIdx.y * MATRIX_TASKS_PER_CTA;
  const uint32_t t1 = min(task_hi, t0 + MATRIX_TASKS_PER_CTA);
  ColAcc acc{tilebuf, (uint32_t)tid, TILE_NODES, S.cf.R, S.cf.W};
  for (uint32_t t = t0; t < t1; ++t) {
    __syncthreads();
    {
      const uint32_t* src = reinterpret_cast<const uint32_t*>(&S.classes[task_class[t]]);
      uint32_t* dst = reinterpret_cast<uint32_t*>(&cls);
      for (uint32_t i = tid; i < sizeof(ClassRec) / 4; i += SCAN_THREADS) dst[i] = src[i];
    }
    __syncthreads();
    if (node < S.N) {
      const uint64_t key = eval_pair(S.cf, cls, acc, node, nullptr);
      const size_t o = (size_t)(t - task_lo) * S.N + node;
      if (fit) fit[o] = key != 0;
      if (score) score[o] = key ? (double)(key_score(key) - S.cf.score_bias) : 0.0;
    }
  }
}

// K1+K2+K3 fused: per-task best packed key over ALL nodes in one launch.
__global__ void __launch_bounds__(SCAN_THREADS)
best_nodes_kernel(const __grid_constant__ DevSession S, const uint32_t* __restrict__ task_class, uint32_t task_lo, uint32_t task_hi,
                  unsigned long long* __restrict__ best_key) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ ClassRec cls;
  __shared__ uint64_t mbar;
  uint64_t* tilebuf = reinterpret_cast<uint64_t*>(smem_raw);
  const int tid = threadIdx.x, lane = tid & 31;
  const uint32_t tile_u64 = S.ncols * TILE_NODES;
  if (tid == 0) {
    mbar_init(&mbar, 1); fence_mbar_init();
    mbar_expect_tx(&mbar, tile_u64 * 8u);
    tma_load_1d(tilebuf, S.tiles + (size_t)blockIdx.x * tile_u64, tile_u64 * 8u, &mbar);
  }
  __syncthreads();
  mbar_wait(&mbar, 0);
  const uint32_t node = blockIdx.x * TILE_NODES + tid;
  const uint32_t chunk = (task_hi - task_lo + gridDim.y - 1) / gridDim.y;
  const uint32_t t0 = task_lo + blockIdx.y * chunk;
  const uint32_t t1 = min(task_hi, t0 + chunk);
  ColAcc acc{tilebuf, (uint32_t)tid, TILE_NODES, S.cf.R, S.cf.W};
  uint32_t cur_cls = 0xFFFFFFFFu;
  for (uint32_t t = t0; t < t1; ++t) {
    const uint32_t cid = task_class[t];
    if (cid != cur_cls) {            // tasks of a PodGroup share a class: reload only on change (uniform branch)
      __syncthreads();
      const uint32_t* src = reinterpret_cast<const uint32_t*>(&S.classes[cid]);
      uint32_t* dst = reinterpret_cast<uint32_t*>(&cls);
      for (uint32_t i = tid; i < sizeof(ClassRec) / 4; i += SCAN_THREADS) dst[i] = src[i];
      __syncthreads();
      cur_cls = cid;
    }
    uint64_t key = node < S.N ? eval_pair(S.cf, cls, acc, node, nullptr) : 0ull;
    key = warp_max_u64(key);
    if (lane == 0 && key) atomicMax(&best_key[t - task_lo], (unsigned long long)key);
  }
}

}  // namespace kb
