// kb_emu.cpp — CPU emulation of the DEVICE ALGORITHM (test infrastructure, never shipped).
//
// Compiles the product's host/device-shared headers (kb_core.h, kb_ctl.h, kb_build.h) with g++ and
// re-enacts the kernels step by step in one thread:
//   scan    (visit_kernel's scan + merge)  this rank's tiles -> exact top-KTOP keys + their node records
//   replay  (replay_epilogue)              merge the ranks' lists, lane-owned candidates with look-ahead,
//                                          certification rule, AddTask bookkeeping, control plane
//   finish  (gang_commit_kernel)           per-PodGroup prefix rule
// With world == 1 scan feeds replay directly; with world > 1 the caller all-gathers the send buffers
// (tests use torch.distributed/gloo) exactly like ncclAllGather does between the two kernels.
// It shares NO code with oracle/ and lets `-m "not gpu"` tests check the engine's logic (everything except
// the CUDA thread mechanics) against the oracle without a GPU.
#include <algorithm>
#include <cstring>
#include <string>
#include <vector>

#include "../../kube_batch_b200/csrc/kb_build.h"
#include "../../kube_batch_b200/csrc/kb_evict_build.h"

using namespace kb;

namespace {
thread_local std::string g_err;

constexpr uint32_t MAXC = 2 * KB_MAX_R + 6 + 3 * KB_MAX_W;
struct Slot { uint64_t col[MAXC]; };
struct SlotAcc {
  const Slot* s; uint32_t R, W;
  uint64_t col(uint32_t c) const { return s->col[c]; }
  double idle(uint32_t r) const { return u64_as_double(col(col_idle(R, r))); }
  double rel(uint32_t r) const { return u64_as_double(col(col_rel(R, r))); }
  int64_t alloc_cpu() const { return (int64_t)col(col_alloc_cpu(R)); }
  int64_t alloc_mem() const { return (int64_t)col(col_alloc_mem(R)); }
  int64_t nz_cpu() const { return (int64_t)col(col_nz_cpu(R)); }
  int64_t nz_mem() const { return (int64_t)col(col_nz_mem(R)); }
  int32_t pods() const { return (int32_t)(uint32_t)(col(col_pods(R)) & 0xFFFFFFFFull); }
  int32_t max_pods() const { return (int32_t)(uint32_t)(col(col_pods(R)) >> 32); }
  uint32_t flags() const { return (uint32_t)col(col_flags(R)); }
  uint64_t labels(uint32_t w) const { return col(col_labels(R, W, w)); }
  uint64_t taints(uint32_t w) const { return col(col_taints(R, W, w)); }
  uint64_t ports(uint32_t w) const { return col(col_ports(R, W, w)); }
};

struct Emu {
  BuiltSession B;
  DevSession S{};            // allocate view
  DevSession Sbf{};          // backfill view (kb_backfill)
  DevSession* cur = nullptr; // view the launches run on
  uint32_t launches = 0;
  // prototype of the two-pass scan for preferred node affinity (a12): pass 1 of a launch leaves the max count over the
  // feasible nodes and how many feasible nodes reach it; the replay of that launch consumes them
  int64_t pref_max = 0;
  uint32_t pref_nmax = 0;
  // persistent pipeline (cycle_kernel): modification log with the records, the scanners' lagging view of the table
  bool pipe = false;
  std::vector<uint32_t> log_node;
  std::vector<std::vector<uint64_t>> log_rec;
  std::vector<uint64_t> view;          // scanner-side copy of the node tiles
  uint32_t view_pos = 0;               // log entries applied to it
  uint64_t rng = 0x9E3779B97F4A7C15ull;
  bool need_fresh = false;             // the last visit stopped for a rescan
  uint32_t next() { rng = rng * 6364136223846793005ull + 1442695040888963407ull; return (uint32_t)(rng >> 33); }
};

struct PrefCtx { const ClassPref* cp; int64_t w_nodeaff; int64_t max; uint32_t* nmax; };
// key of a feasible node with the NormalizeReduce(10) term added to the (biased) score half (kb_core.h add_pref_term)
inline uint64_t add_pref_term(uint64_t key, const PrefCtx* pc, int64_t count) {
  if (!pc) return key;
  return kb::add_pref_term(key, pc->w_nodeaff, count, pc->max);
}

// visit_kernel, scan + merge + (sharded) pack: keys[32] then columns [ncols][32]
void emu_scan(Emu& E, uint64_t* sendbuf) {
  const DevSession& S = *E.cur;
  const Ctl& c = *S.ctl;
  const size_t cnt = (size_t)xchg_u64(S.ncols);
  std::fill(sendbuf, sendbuf + cnt, 0ull);
  if (c.done) return;
  const ClassRec& cls = S.classes[c.cur_class];
  const uint32_t R = S.cf.R, W = S.cf.W;
  const size_t tile_u64 = (size_t)S.ncols * TILE_NODES;
  // pass 1 (only for a class with preferred terms): max count over the FEASIBLE nodes + how many feasible nodes reach it
  const ClassPref* cp = (E.B.has_pref && S.cf.nodeorder && !S.backfill && E.B.class_pref[c.cur_class].n) ? &E.B.class_pref[c.cur_class] : nullptr;
  E.pref_max = 0; E.pref_nmax = 0;
  // inter-pod affinity (kb_aff.h): step 10 joins the plugin predicates; a class with a weight list gets the priority's passes
  // first (aff_prepass_kernel): pass 1 = weight per topology domain over the FEASIBLE nodes, pass 2 = min / max count
  const AffDev& A = S.aff;
  const ClassAff* ca = A.on ? &A.cls[c.cur_class] : nullptr;
  const bool ipa = ca && ca->w_cnt && S.cf.nodeorder;
  auto feasible_key = [&](uint32_t n, bool* pok) {
    TileAcc acc{S.tiles + (size_t)(n / TILE_NODES) * tile_u64, n % TILE_NODES, R, W};
    uint64_t k = eval_pair(S.cf, cls, acc, n, nullptr, pok);
    if (ca && S.cf.predicates && !aff_pred(A, *ca, S.N, n)) { k = 0; if (pok) *pok = false; }
    if (S.backfill && c.pred_dead) { k = 0; if (pok) *pok = false; }      // Ctl.pred_dead: every predicate fails once a task is Allocated on no node
    return k;
  };
  if (cp)                       // aff_prepass_kernel<2>: max preferred count over the FEASIBLE nodes (incl. predicate step 10)
    for (uint32_t n = 0; n < S.N; ++n) {
      if (!feasible_key(n, nullptr)) continue;
      TileAcc acc{S.tiles + (size_t)(n / TILE_NODES) * tile_u64, n % TILE_NODES, R, W};
      const int64_t cnt = pref_count(*cp, acc, W);
      if (cnt > E.pref_max) { E.pref_max = cnt; E.pref_nmax = 1; }
      else if (cnt == E.pref_max) E.pref_nmax += 1;
    }
  uint32_t nmax_dummy = 0;
  PrefCtx pc{cp, E.B.hc.w_nodeaff, E.pref_max, &nmax_dummy};
  if (ipa) {
    for (uint32_t i = 0; i < A.dom_total; ++i) A.dom_sum[i] = 0;
    A.minmax[0] = 0; A.minmax[1] = 0;
    for (uint32_t n = 0; n < S.N; ++n)
      if (feasible_key(n, nullptr)) aff_pass1_node(A, *ca, S.N, n, [&](uint32_t slot, long long v) { A.dom_sum[slot] += v; });
    for (uint32_t n = 0; n < S.N; ++n)
      if (feasible_key(n, nullptr)) {
        const long long cnt = aff_count_node(A, *ca, S.N, n);
        if (cnt < A.minmax[0]) A.minmax[0] = cnt;
        if (cnt > A.minmax[1]) A.minmax[1] = cnt;
      }
  }
  std::vector<uint64_t> keys;
  for (uint32_t t = S.tile_lo; t < S.tile_hi; ++t)
    for (uint32_t i = 0; i < TILE_NODES; ++i) {
      const uint32_t n = t * TILE_NODES + i;
      if (n >= S.N) break;
      TileAcc acc{S.tiles + (size_t)t * tile_u64, i, R, W};
      bool pok = false;
      uint64_t k = feasible_key(n, &pok);
      if (S.backfill && pok && !k) sendbuf[(size_t)(1 + S.ncols) * 32] = 1;      // flag row: passes ssn.PredicateFn, no Idle for Resreq
      if (k && cp) k = add_pref_term(k, &pc, pref_count(*cp, acc, W));
      if (k && ipa) k = aff_add_score(k, A.w_podaff, aff_score(aff_count_node(A, *ca, S.N, n), A.minmax[0], A.minmax[1]));
      if (k) keys.push_back(k);
    }
  std::sort(keys.begin(), keys.end(), [](uint64_t a, uint64_t b) { return a > b; });
  for (uint32_t l = 0; l < 32 && l < keys.size(); ++l) {
    sendbuf[l] = keys[l];
    const uint32_t n = key_node(keys[l]);
    const uint64_t* gt = S.tiles + (size_t)(n / TILE_NODES) * tile_u64 + (n % TILE_NODES);
    for (uint32_t cc = 0; cc < S.ncols; ++cc) sendbuf[(size_t)(1 + cc) * 32 + l] = gt[(size_t)cc * TILE_NODES];
  }
}

struct Cand { bool have = false, cur_fi = false, next_fi = false, next_valid = false, modified = false;
              uint64_t cur_key = 0, next_key = 0; uint32_t node = 0, cnt = 0; Slot st[2]; int which = 0; int64_t pref = 0; };

// replay_epilogue's core: look-ahead refresh + certified steps + control plane, for candidates already loaded
bool replay_core(const DevSession& S, Ctl& c, const uint32_t cls_id, std::vector<Cand>& cand, const uint64_t floor_key,
                 const PrefCtx* pc = nullptr, const bool pred_any_outside = false, const bool pref_stop_each = false) {
  const ClassRec& cls = S.classes[cls_id];
  const uint32_t R = S.cf.R, W = S.cf.W, ncols = S.ncols;
  const ClassAff* ca = S.aff.on ? &S.aff.cls[cls_id] : nullptr;
  const bool ca_reads = ca && aff_stop_each(*ca, S.cf.nodeorder != 0);
  bool aff_stale = false;       // a placement of a class whose keys read the inter-pod counters: the list is used once
  auto refresh = [&]() {
    for (auto& cd : cand) {
      if (!(cd.have && cd.cur_key != 0 && !cd.next_valid)) continue;
      const Slot& src = cd.st[cd.which];
      Slot& dst = cd.st[cd.which ^ 1];
      for (uint32_t cc = 0; cc < ncols; ++cc) dst.col[cc] = src.col[cc];
      const uint32_t base_col = cd.cur_fi ? col_idle(R, 0) : col_rel(R, 0);
      for (uint32_t k = 0; k < R; ++k) dst.col[base_col + k] = double_as_u64(u64_as_double(src.col[base_col + k]) - cls.resreq[k]);
      dst.col[col_nz_cpu(R)] = (uint64_t)((int64_t)src.col[col_nz_cpu(R)] + cls.nz_cpu);
      dst.col[col_nz_mem(R)] = (uint64_t)((int64_t)src.col[col_nz_mem(R)] + cls.nz_mem);
      dst.col[col_pods(R)] = src.col[col_pods(R)] + 1ull;
      for (uint32_t w = 0; w < W; ++w) dst.col[col_ports(R, W, w)] = src.col[col_ports(R, W, w)] | cls.port_own[w] | (cd.cur_fi ? cls.aff_own[w] : 0ull);
      SlotAcc acc{&dst, R, W};
      bool f = false;
      cd.next_key = add_pref_term(eval_pair(S.cf, cls, acc, cd.node, &f), pc, cd.pref);
      if (ca && cd.cur_fi && aff_self_blocks(S.aff, *ca, S.N, cd.node)) cd.next_key = 0;       // one replica per host: the class's own pod, once ALLOCATED, forbids the node
      cd.next_fi = f; cd.next_valid = true;
      c.pairs_replayed += 1;
    }
  };
  refresh();
  bool pref_stale = false;      // the last feasible max-count node filled up: every key of this launch used a stale normalisation
  for (;;) {
    if (c.done || c.cur_class != cls_id || pref_stale || aff_stale) break;
    const uint32_t j = (uint32_t)c.cur_job;
    const uint32_t jend = S.job_ord_off[j + 1];
    uint32_t run_left = c.cur_run, placed = 0, reason = STOP_RUN_DONE;
    while (run_left > 0) {
      uint64_t best = 0; int owner = -1;
      for (int l = 0; l < KTOP; ++l) if (cand[l].cur_key > best) { best = cand[l].cur_key; owner = l; }
      if (best < floor_key) { reason = STOP_RESCAN; break; }
      const uint32_t pos = S.job_pos[j];
      S.job_pos[j] = pos + 1;
      c.tasks_processed += 1; c.pairs_logical += S.N;
      run_left -= 1;
      if (best == 0) {
        if (S.backfill) {
          // phantom Allocated (session.go:241-262): some node passes ssn.PredicateFn but none has Idle for Resreq
          bool any = pred_any_outside;
          for (auto& cd : cand) {
            if (!cd.have) continue;
            SlotAcc acc{&cd.st[cd.which], R, W};
            bool pok = false;
            eval_pair(S.cf, cls, acc, cd.node, nullptr, &pok);
            if (ca && S.cf.predicates && !aff_pred(S.aff, *ca, S.N, cd.node)) pok = false;
            any = any || pok;
          }
          if (any && S.cf.fit_mode != 2) {
            kb_decision dd;
            dd.node = -1; dd.kind = KB_KIND_ALLOCATED; dd.dispatched = 0; dd.reserved = 0; dd.step = 0xFFFFFFFFu; dd.dispatch_step = 0xFFFFFFFFu;
            S.dec[S.ord_task[pos]] = dd;
            S.job_ready[j] += 1;
            c.phantoms += 1;
          }
        }
        reason = STOP_NOFIT; break;
      }
      Cand& cd = cand[owner];
      if (S.cf.fit_mode == 2 && !cd.cur_fi) {
        // backfill with the predicates plugin: the FIRST node that passes ssn.PredicateFn is tried; node.AddTask refuses
        // (Resreq > Idle) -> the task stays Allocated on no node and every later predicate fails (Ctl.pred_dead)
        kb_decision dd;
        dd.node = -1; dd.kind = KB_KIND_ALLOCATED; dd.dispatched = 0; dd.reserved = 0; dd.step = 0xFFFFFFFFu; dd.dispatch_step = 0xFFFFFFFFu;
        S.dec[S.ord_task[pos]] = dd;
        S.job_ready[j] += 1;
        c.phantoms += 1;
        c.pred_dead = 1;
        for (auto& x : cand) { x.cur_key = 0; x.next_key = 0; x.next_valid = true; }
        reason = STOP_NOFIT; break;
      }
      if (!cd.next_valid) refresh();
      const bool fits_idle = cd.cur_fi;
      cd.cnt += 1;
      cd.which ^= 1; cd.cur_key = cd.next_key; cd.cur_fi = cd.next_fi; cd.next_valid = false; cd.modified = true;
      kb_decision dd;
      dd.node = (int32_t)key_node(best); dd.kind = fits_idle ? KB_KIND_ALLOCATED : KB_KIND_PIPELINED; dd.dispatched = 0; dd.reserved = 0;
      dd.step = c.step; dd.dispatch_step = 0xFFFFFFFFu;
      S.dec[S.ord_task[pos]] = dd;
      c.step += 1;
      if (fits_idle) { c.tasks_allocated += 1; S.job_ready[j] += 1; } else c.tasks_pipelined += 1;
      S.job_placed[j] += 1;
      on_allocate_event(S, j, cls);
      if (ca) { aff_commit(S.aff, *ca, S.N, cd.node, fits_idle); aff_stale = ca_reads; }
      if (pc && pref_stop_each) aff_stale = true;       // visit_kernel<.,1>: a class with preferred node-affinity terms is scanned afresh per task
      placed += 1;
      if (pc && pc->max > 0 && cd.cur_key == 0 && cd.pref == pc->max) {            // a max-count node left the feasible set
        *pc->nmax -= 1;
        if (*pc->nmax == 0) pref_stale = true;
      }
      if (ssn_job_ready(S, j) && (pos + 1 < jend) && !S.backfill) { reason = STOP_YIELD; break; }
      if (pref_stale || aff_stale) { if (run_left > 0) reason = STOP_RESCAN; break; }
    }
    if (reason == STOP_RESCAN) c.rescans += 1;
    after_run(S, c, reason, placed);
    if (reason == STOP_RESCAN) return true;
  }
  return false;
}

void write_back(const DevSession& S, const ClassRec& cls, std::vector<Cand>& cand) {
  const uint32_t R = S.cf.R, ncols = S.ncols;
  const size_t tile_u64 = (size_t)ncols * TILE_NODES;
  for (auto& cd : cand) {
    if (!cd.modified) continue;
    const Slot& src = cd.st[cd.which];
    uint64_t* gt = S.tiles + (size_t)(cd.node / TILE_NODES) * tile_u64 + (cd.node % TILE_NODES);
    for (uint32_t cc = 0; cc < ncols; ++cc) gt[(size_t)cc * TILE_NODES] = src.col[cc];
    for (uint32_t r = 0; r < R; ++r)
      for (uint32_t i = 0; i < cd.cnt; ++i) S.node_used[(size_t)r * S.N + cd.node] += cls.resreq[r];
  }
}

// replay_kernel / replay_epilogue
void emu_replay(Emu& E, const uint64_t* recvbuf) {
  const DevSession& S = *E.cur;
  Ctl& c = *S.ctl;
  E.launches += 1;
  if (c.done) return;
  const uint32_t cls_id = c.cur_class;
  const ClassRec& cls = S.classes[cls_id];
  const uint32_t R = S.cf.R, W = S.cf.W, ncols = S.ncols;
  const size_t rank_u64 = (size_t)xchg_u64(ncols);
  bool pred_any = false;
  for (uint32_t r = 0; r < S.world; ++r) pred_any = pred_any || recvbuf[r * rank_u64 + (size_t)(1 + ncols) * 32] != 0;
  // merge the ranks' lists
  struct Src { uint64_t key; uint32_t rank, idx; };
  std::vector<Src> all;
  for (uint32_t r = 0; r < S.world; ++r)
    for (uint32_t i = 0; i < 32; ++i) {
      uint64_t k = recvbuf[r * rank_u64 + i];
      if (k) all.push_back({k, r, i});
    }
  std::sort(all.begin(), all.end(), [](const Src& a, const Src& b) { return a.key > b.key; });
  c.scans += 1; c.pairs_scanned += S.N;

  std::vector<Cand> cand(KTOP);
  const uint64_t floor_key = all.size() >= (size_t)KTOP ? all[KTOP - 1].key : 0ull;
  for (int l = 0; l < KTOP && l < (int)all.size(); ++l) {
    Cand& cd = cand[l];
    cd.cur_key = all[l].key; cd.have = true;
    cd.node = key_node(cd.cur_key);
    const uint64_t* rec = recvbuf + all[l].rank * rank_u64 + 32 + all[l].idx;
    for (uint32_t cc = 0; cc < ncols; ++cc) cd.st[0].col[cc] = rec[(size_t)cc * 32];
    SlotAcc acc{&cd.st[0], R, W};
    cd.cur_fi = res_less_equal(R, [&](uint32_t k) { return cls.initreq[k]; }, [&](uint32_t k) { return acc.idle(k); });
  }
  const ClassPref* cp = (E.B.has_pref && S.cf.nodeorder && !S.backfill && E.B.class_pref[cls_id].n) ? &E.B.class_pref[cls_id] : nullptr;
  PrefCtx pc{cp, E.B.hc.w_nodeaff, E.pref_max, &E.pref_nmax};
  if (cp) for (auto& cd : cand) if (cd.have) { SlotAcc acc{&cd.st[0], R, W}; cd.pref = pref_count(*cp, acc, W); }
  replay_core(S, c, cls_id, cand, floor_key, cp ? &pc : nullptr, pred_any, S.aff.on != 0);      // sharded prototype (no counter path): countdown rule
  write_back(S, cls, cand);          // every replica writes every modified candidate back
}

// One launch of visit_kernel in OVERLAP mode (single GPU): the scanner CTAs evaluate Ctl.scan_class (the predicted class of
// the visit after the current one) skipping Ctl.excl, while the replayer CTA consumes Ctl.list; the last CTA merges the scan
// lists with the replayer's patch keys and publishes list / scan_class / excl for the next launch.
void emu_launch_overlap(Emu& E) {
  const DevSession& S = *E.cur;
  Ctl& c = *S.ctl;
  E.launches += 1;
  if (c.done) return;
  const uint32_t R = S.cf.R, W = S.cf.W, ncols = S.ncols;
  const size_t tile_u64 = (size_t)ncols * TILE_NODES;
  const uint32_t scan_class = c.scan_class, n_excl = c.n_excl;
  uint32_t excl[32];
  for (uint32_t i = 0; i < 32; ++i) excl[i] = c.excl[i];

  // ---- replayer CTA ----
  std::vector<Cand> cand(KTOP);
  bool replayed = false;
  if (c.list_valid && c.list_class == c.cur_class) {
    replayed = true;
    const uint32_t cls_id = c.list_class;
    const ClassRec& cls = S.classes[cls_id];
    c.list_valid = 0;                                  // consumed
    const uint64_t floor_key = c.list[KTOP - 1];
    for (int l = 0; l < KTOP; ++l) {
      Cand& cd = cand[l];
      cd.cur_key = c.list[l]; cd.have = cd.cur_key != 0;
      if (!cd.have) continue;
      cd.node = key_node(cd.cur_key);
      const uint64_t* gt = S.tiles + (size_t)(cd.node / TILE_NODES) * tile_u64 + (cd.node % TILE_NODES);
      for (uint32_t cc = 0; cc < ncols; ++cc) cd.st[0].col[cc] = gt[(size_t)cc * TILE_NODES];
      SlotAcc acc{&cd.st[0], R, W};
      cd.cur_fi = res_less_equal(R, [&](uint32_t k) { return cls.initreq[k]; }, [&](uint32_t k) { return acc.idle(k); });
    }
    replay_core(S, c, cls_id, cand, floor_key);
    write_back(S, cls, cand);
  }
  c.patch_valid = 0;
  if (replayed && !c.done && n_excl > 0 && c.cur_class == scan_class) {      // prediction hit: fresh keys of the excluded nodes
    const ClassRec& pc = S.classes[scan_class];
    std::vector<uint64_t> pk;
    for (auto& cd : cand) {
      if (!cd.have) continue;
      SlotAcc acc{&cd.st[cd.which], R, W};
      uint64_t k = eval_pair(S.cf, pc, acc, cd.node, nullptr);
      c.pairs_replayed += 1;
      if (k) pk.push_back(k);
    }
    std::sort(pk.begin(), pk.end(), [](uint64_t a, uint64_t b) { return a > b; });
    for (int l = 0; l < KTOP; ++l) c.patch[l] = l < (int)pk.size() ? pk[l] : 0ull;
    c.patch_valid = 1;
  }

  // ---- scanner CTAs (concurrent on the device; equivalent in any order because they skip the excluded nodes) ----
  std::vector<uint64_t> keys;
  {
    const ClassRec& sc = S.classes[scan_class];
    for (uint32_t n = 0; n < S.N; ++n) {
      bool skip = false;
      for (uint32_t i = 0; i < n_excl; ++i) skip = skip || excl[i] == n;
      if (skip) continue;
      TileAcc acc{S.tiles + (size_t)(n / TILE_NODES) * tile_u64, n % TILE_NODES, R, W};
      uint64_t k = eval_pair(S.cf, sc, acc, n, nullptr);
      if (k) keys.push_back(k);
    }
  }
  // ---- merger (last CTA) ----
  if (c.patch_valid) for (int l = 0; l < KTOP; ++l) if (c.patch[l]) keys.push_back(c.patch[l]);
  std::sort(keys.begin(), keys.end(), [](uint64_t a, uint64_t b) { return a > b; });
  for (int l = 0; l < KTOP; ++l) c.list[l] = l < (int)keys.size() ? keys[l] : 0ull;
  c.list_class = scan_class;
  c.list_valid = (n_excl == 0 || c.patch_valid) ? 1u : 0u;
  c.scans += 1; c.pairs_scanned += S.N;
  if (n_excl > 0) { c.predictions += 1; if (!c.patch_valid) c.mispredictions += 1; }
  if (!c.done) {
    if (c.list_valid && c.list_class == c.cur_class) {
      const uint32_t pk = S.ord_peek[S.job_pos[(uint32_t)c.cur_job]];
      c.scan_class = pk != 0xFFFFFFFFu ? pk : c.cur_class;
      uint32_t n = 0;
      for (int l = 0; l < KTOP; ++l) if (c.list[l]) c.excl[n++] = key_node(c.list[l]);
      c.n_excl = n;
    } else {
      c.scan_class = c.cur_class;
      c.n_excl = 0;
    }
  }
}

// One launch of visit_chain_kernel<K> (single GPU): ONE pass over the table evaluates cur_class and the predicted classes
// of the following visits (Ctl.chain); the first visit is replayed as usual, every following visit whose class has a
// look-ahead list is replayed in the same launch after the list was PATCHED: entries of nodes modified since the scan are
// dropped, those nodes are re-evaluated against their current records, and the certification floor rises to the largest
// key that no longer fits into the 32 lanes.
void emu_launch_chain(Emu& E) {
  const DevSession& S = *E.cur;
  Ctl& c = *S.ctl;
  E.launches += 1;
  if (c.done) return;
  const uint32_t R = S.cf.R, W = S.cf.W, ncols = S.ncols;
  const size_t tile_u64 = (size_t)ncols * TILE_NODES;
  uint32_t cls[KB_CHAIN_MAX]; uint32_t nK = 1;
  cls[0] = c.cur_class;
  for (uint32_t k = 0; k + 1 < S.kchain && k + 1 < KB_CHAIN_MAX; ++k) { if (c.chain[k] == 0xFFFFFFFFu) break; cls[nK++] = c.chain[k]; }
  // ---- scan: every class against the table as it is at launch start ----
  std::vector<std::vector<uint64_t>> lists(nK);
  std::vector<uint64_t> floor(nK, 0);
  for (uint32_t k = 0; k < nK; ++k) {
    std::vector<uint64_t> keys;
    for (uint32_t n = 0; n < S.N; ++n) {
      TileAcc acc{S.tiles + (size_t)(n / TILE_NODES) * tile_u64, n % TILE_NODES, R, W};
      uint64_t key = eval_pair(S.cf, S.classes[cls[k]], acc, n, nullptr);
      if (key) keys.push_back(key);
    }
    std::sort(keys.begin(), keys.end(), [](uint64_t a, uint64_t b) { return a > b; });
    keys.resize(KTOP, 0ull);
    lists[k] = keys; floor[k] = keys[KTOP - 1];
  }
  std::vector<uint32_t> mod;          // nodes modified in this launch, no duplicates
  uint32_t used = 0;
  uint32_t k = 0;
  for (;;) {
    used |= 1u << k;
    const ClassRec& cr = S.classes[cls[k]];
    c.scans += 1; c.pairs_scanned += S.N;
    std::vector<Cand> cand(KTOP);
    for (int l = 0; l < KTOP; ++l) {
      Cand& cd = cand[l];
      cd.cur_key = lists[k][l]; cd.have = cd.cur_key != 0;
      if (!cd.have) continue;
      cd.node = key_node(cd.cur_key);
      const uint64_t* gt = S.tiles + (size_t)(cd.node / TILE_NODES) * tile_u64 + (cd.node % TILE_NODES);
      for (uint32_t cc = 0; cc < ncols; ++cc) cd.st[0].col[cc] = gt[(size_t)cc * TILE_NODES];
      SlotAcc acc{&cd.st[0], R, W};
      cd.cur_fi = res_less_equal(R, [&](uint32_t q) { return cr.initreq[q]; }, [&](uint32_t q) { return acc.idle(q); });
    }
    const bool rescan = replay_core(S, c, cls[k], cand, floor[k]);
    write_back(S, cr, cand);
    for (auto& cd : cand) if (cd.modified && std::find(mod.begin(), mod.end(), cd.node) == mod.end()) mod.push_back(cd.node);
    if (c.done || rescan) break;
    // next visit: is there an unused look-ahead list of its class?
    uint32_t nk = 0xFFFFFFFFu;
    for (uint32_t z = 0; z < nK; ++z) if (!((used >> z) & 1u) && cls[z] == c.cur_class) { nk = z; break; }
    if (nk == 0xFFFFFFFFu) break;
    if (mod.size() + KTOP > (size_t)KTOP * KB_CHAIN_MAX) break;            // capacity of the device's modified-node list
    // ---- patch list nk ----
    std::vector<uint64_t> keep, fresh;
    for (uint64_t key : lists[nk]) if (key && std::find(mod.begin(), mod.end(), key_node(key)) == mod.end()) keep.push_back(key);
    uint64_t fl = floor[nk];
    for (size_t base = 0; base < mod.size(); base += 32) {                  // the device merges 32 fresh keys at a time
      std::vector<uint64_t> all = keep;
      for (size_t i = base; i < mod.size() && i < base + 32; ++i) {
        const uint32_t n = mod[i];
        TileAcc acc{S.tiles + (size_t)(n / TILE_NODES) * tile_u64, n % TILE_NODES, R, W};
        const uint64_t key = eval_pair(S.cf, S.classes[cls[nk]], acc, n, nullptr);
        c.pairs_replayed += 1;
        if (key) all.push_back(key);
      }
      std::sort(all.begin(), all.end(), [](uint64_t a, uint64_t b) { return a > b; });
      if (all.size() > (size_t)KTOP) { fl = std::max(fl, all[KTOP]); all.resize(KTOP); }
      keep = all;
    }
    keep.resize(KTOP, 0ull);
    lists[nk] = keep; floor[nk] = fl;
    c.chain_hits += 1;
    k = nk;
  }
  publish_chain(S, c);
}

// One visit chain of cycle_kernel (kb_pipe.cuh).  The scanners answered the request for cur_class from THEIR copy of the
// table as of an earlier log position (`stamp`, random lag of up to PIPE_PATCH entries; the copy only moves forward); nodes
// with entries in log[stamp, head) may have been read torn, so their scanned keys are ARBITRARY (old state, current state,
// zero, or an absurdly good score).  The replayer drops those nodes from the list, re-evaluates them on their current
// records (hot ring), keeps the 32 best, raises the floor to the best key dropped, and replays.
void emu_launch_pipe(Emu& E) {
  const DevSession& S = *E.cur;
  Ctl& c = *S.ctl;
  E.launches += 1;
  if (c.done) return;
  const uint32_t R = S.cf.R, W = S.cf.W, ncols = S.ncols;
  const size_t tile_u64 = (size_t)ncols * TILE_NODES;
  const uint32_t cls_id = c.cur_class;
  const ClassRec& cls = S.classes[cls_id];
  const uint32_t head = (uint32_t)E.log_node.size();
  if (E.view.empty()) { E.view.assign(S.tiles, S.tiles + (size_t)std::max(1u, S.NT) * tile_u64); E.view_pos = head; }
  // a12: a class with preferred node-affinity terms only ever uses a list of the CURRENT table state (kb_pipe.cuh)
  const ClassPref* cp = (E.B.has_pref && S.cf.nodeorder && E.B.class_pref[cls_id].n) ? &E.B.class_pref[cls_id] : nullptr;
  uint32_t lag = (E.need_fresh || cp) ? 0u : E.next() % (PIPE_PATCH + 1);
  if (E.next() % 4 == 0) lag = 0;
  uint32_t stamp = head > lag ? head - lag : 0;
  if (stamp < E.view_pos) stamp = E.view_pos;                  // the applier never goes back
  E.need_fresh = false;
  for (uint32_t i = E.view_pos; i < stamp; ++i) {              // applier: entries below the stamp are in the resident tiles
    const uint32_t n = E.log_node[i];
    uint64_t* vt = E.view.data() + (size_t)(n / TILE_NODES) * tile_u64 + (n % TILE_NODES);
    for (uint32_t cc = 0; cc < ncols; ++cc) vt[(size_t)cc * TILE_NODES] = E.log_rec[i][cc];
  }
  E.view_pos = stamp;
  std::vector<uint32_t> inflight;                              // distinct nodes of log[stamp, head)
  for (uint32_t i = stamp; i < head; ++i) if (std::find(inflight.begin(), inflight.end(), E.log_node[i]) == inflight.end()) inflight.push_back(E.log_node[i]);
  // ---- scanners ----
  // pass 1 (preferred terms): max count over the feasible nodes, nodes reaching it
  E.pref_max = 0; E.pref_nmax = 0;
  if (cp)
    for (uint32_t n = 0; n < S.N; ++n) {
      TileAcc acc{E.view.data() + (size_t)(n / TILE_NODES) * tile_u64, n % TILE_NODES, R, W};
      if (!eval_pair(S.cf, cls, acc, n, nullptr)) continue;
      const int64_t cnt = pref_count(*cp, acc, W);
      if (cnt > E.pref_max) { E.pref_max = cnt; E.pref_nmax = 1; }
      else if (cnt == E.pref_max) E.pref_nmax += 1;
    }
  PrefCtx pc{cp, E.B.hc.w_nodeaff, E.pref_max, &E.pref_nmax};
  std::vector<uint64_t> keys;
  for (uint32_t n = 0; n < S.N; ++n) {
    uint64_t k;
    if (std::find(inflight.begin(), inflight.end(), n) != inflight.end()) {
      const uint32_t how = E.next() % 4;
      if (how == 0) { TileAcc acc{E.view.data() + (size_t)(n / TILE_NODES) * tile_u64, n % TILE_NODES, R, W}; k = eval_pair(S.cf, cls, acc, n, nullptr); }
      else if (how == 1) { TileAcc acc{S.tiles + (size_t)(n / TILE_NODES) * tile_u64, n % TILE_NODES, R, W}; k = eval_pair(S.cf, cls, acc, n, nullptr); }
      else if (how == 2) k = 0;
      else k = pack_key(S.cf.score_bias + 1000 + (int64_t)(E.next() % 7), n);        // garbage from a torn read
    } else {
      TileAcc acc{E.view.data() + (size_t)(n / TILE_NODES) * tile_u64, n % TILE_NODES, R, W};
      k = eval_pair(S.cf, cls, acc, n, nullptr);
      if (k && cp) k = add_pref_term(k, &pc, pref_count(*cp, acc, W));
    }
    if (k) keys.push_back(k);
  }
  std::sort(keys.begin(), keys.end(), [](uint64_t a, uint64_t b) { return a > b; });
  keys.resize(KTOP, 0ull);
  uint64_t fl = keys[KTOP - 1];
  // ---- replayer: patch ----
  std::vector<uint64_t> pool;
  for (uint64_t k : keys) if (k && std::find(inflight.begin(), inflight.end(), key_node(k)) == inflight.end()) pool.push_back(k);
  for (uint32_t n : inflight) {
    TileAcc acc{S.tiles + (size_t)(n / TILE_NODES) * tile_u64, n % TILE_NODES, R, W};
    const uint64_t k = eval_pair(S.cf, cls, acc, n, nullptr);
    c.pairs_replayed += 1;
    if (k) pool.push_back(k);
  }
  std::sort(pool.begin(), pool.end(), [](uint64_t a, uint64_t b) { return a > b; });
  if (pool.size() > (size_t)KTOP) { fl = std::max(fl, pool[KTOP]); pool.resize(KTOP); }
  c.scans += 1; c.pairs_scanned += S.N;
  std::vector<Cand> cand(KTOP);
  for (int l = 0; l < KTOP && l < (int)pool.size(); ++l) {
    Cand& cd = cand[l];
    cd.cur_key = pool[l]; cd.have = true;
    cd.node = key_node(cd.cur_key);
    const uint64_t* gt = S.tiles + (size_t)(cd.node / TILE_NODES) * tile_u64 + (cd.node % TILE_NODES);
    for (uint32_t cc = 0; cc < ncols; ++cc) cd.st[0].col[cc] = gt[(size_t)cc * TILE_NODES];
    SlotAcc acc{&cd.st[0], R, W};
    cd.cur_fi = res_less_equal(R, [&](uint32_t k) { return cls.initreq[k]; }, [&](uint32_t k) { return acc.idle(k); });
    if (cp) cd.pref = pref_count(*cp, acc, W);
  }
  E.need_fresh = replay_core(S, c, cls_id, cand, fl, cp ? &pc : nullptr);
  write_back(S, cls, cand);
  for (auto& cd : cand) {
    if (!cd.modified) continue;
    E.log_node.push_back(cd.node);
    E.log_rec.emplace_back(cd.st[cd.which].col, cd.st[cd.which].col + ncols);
  }
}

// gang_commit_kernel, serially: a job's processed slots of the allocate view, then of the backfill view
void emu_gang_commit(const DevSession& S, const DevSession& Sbf, const int32_t* ready0) {
  for (uint32_t j = 0; j < S.J; ++j) {
    std::vector<uint32_t> tasks;
    for (uint32_t i = S.job_ord_off[j]; i < S.job_pos[j]; ++i) tasks.push_back(S.ord_task[i]);
    for (uint32_t i = Sbf.job_ord_off[j]; i < Sbf.job_pos[j]; ++i) {
      const uint32_t t = Sbf.ord_task[i];
      if (S.dec[t].kind == KB_KIND_SKIPPED) S.dec[t].kind = KB_KIND_NONE;
      tasks.push_back(t);
    }
    const int32_t need = S.gang_ready ? S.job_min_avail[j] - ready0[j] : 0;
    int32_t incl = 0;
    size_t estar = (size_t)-1; uint32_t estep = 0;
    for (size_t i = 0; i < tasks.size(); ++i) {
      const kb_decision& d = S.dec[tasks[i]];
      if (d.kind != KB_KIND_ALLOCATED) continue;
      incl += 1;                                                 // phantoms (step none) count, but only a real Allocate dispatches
      if (d.step != 0xFFFFFFFFu && incl >= need) { estar = i; estep = d.step; break; }
    }
    if (estar == (size_t)-1) continue;
    for (size_t i = 0; i < tasks.size(); ++i) {
      kb_decision& d = S.dec[tasks[i]];
      if (d.kind != KB_KIND_ALLOCATED) continue;
      if (i <= estar) { d.dispatched = 1; d.dispatch_step = estep; }
      else if (d.step != 0xFFFFFFFFu) { d.dispatched = 1; d.dispatch_step = d.step; }
      else
        for (size_t k = i + 1; k < tasks.size(); ++k) {
          const kb_decision& dk = S.dec[tasks[k]];
          if (dk.kind == KB_KIND_ALLOCATED && dk.step != 0xFFFFFFFFu) { d.dispatched = 1; d.dispatch_step = dk.step; break; }
        }
    }
  }
}

// seed_backfill_kernel + the switch kb_backfill makes
void emu_begin_backfill(Emu& E, bool carry) {
  const Ctl& m = *E.S.ctl;
  Ctl& b = *E.Sbf.ctl;
  if (!b.bf_seeded && !carry) b.bf_seeded = 1;
  if (!b.bf_seeded) {
    b.bf_seeded = 1;
    b.step = m.step;
    b.tasks_processed = m.tasks_processed; b.tasks_allocated = m.tasks_allocated; b.tasks_pipelined = m.tasks_pipelined;
    b.visits = m.visits; b.scans = m.scans; b.rescans = m.rescans;
    b.pairs_logical = m.pairs_logical; b.pairs_scanned = m.pairs_scanned; b.pairs_replayed = m.pairs_replayed;
    b.predictions = m.predictions; b.mispredictions = m.mispredictions;
  }
  E.cur = &E.Sbf;
}
}  // namespace

extern "C" {

// the engine's exact-arithmetic shortcuts next to the forms they replace (tests/test_emu_parity.py pins them on each other)
int kbemu_le(double l, double r, double diff) { return le_func(l, r, diff) ? 1 : 0; }
int kbemu_le_reference_form(double l, double r, double diff) { return le_func_reference_form(l, r, diff) ? 1 : 0; }
long long kbemu_div_0_to_10(long long a, long long b) { return div_0_to_10(a, b); }

uint32_t kbemu_buf_u64(void* h);
const char* kbemu_last_error(void) { return g_err.c_str(); }

// mode (single rank): 0 = scan/replay overlap protocol, 1 = plain one-class launches, 2 / 4 = chained visits (visit_chain_kernel<K>),
// 5 = persistent pipeline (cycle_kernel: stale look-ahead lists + patch; only the R = 3 / W = 2 geometry runs it, other sessions fall back to mode 1)
void* kbemu_create2(const kb_snapshot* snap, const kb_plugin_conf* conf, uint32_t rank, uint32_t world, uint32_t mode) {
  Emu* E = new Emu();
  BuildErr be;
  // mode 1 (plain launches) also accepts preferred node-affinity terms: the emulation prototypes the two-pass scan (a12)
  if (build_session(snap, conf, 148, E->B, &be, rank, world, mode == 0 ? 1 : 0, (mode == 2 || mode == 4) ? mode : 1, mode == 1, mode == 5 ? 1 : 0)) { g_err = be.msg; delete E; return nullptr; }
  E->pipe = mode == 5 && E->B.pipe;
  E->B.bind(E->S, E->B.mut.host.data(), E->B.imm.host.data());
  E->B.bind_backfill(E->Sbf, E->B.mut.host.data(), E->B.imm.host.data());
  E->cur = &E->S;
  return E;
}
// the default emulation exercises the overlap protocol on one rank (the device enables it by size)
void* kbemu_create(const kb_snapshot* snap, const kb_plugin_conf* conf, uint32_t rank, uint32_t world) { return kbemu_create2(snap, conf, rank, world, 0); }
// kb_session_load on an engine that already holds a session: the BuiltSession (and its slabs' capacity) is reused
int kbemu_reload(void* h, const kb_snapshot* snap, const kb_plugin_conf* conf, uint32_t mode) {
  Emu* E = (Emu*)h;
  BuildErr be;
  if (int rc = build_session(snap, conf, 148, E->B, &be, 0, 1, mode == 0 ? 1 : 0, (mode == 2 || mode == 4) ? mode : 1, false, mode == 5 ? 1 : 0)) { g_err = be.msg; return rc; }
  E->pipe = mode == 5 && E->B.pipe; E->log_node.clear(); E->log_rec.clear(); E->view.clear(); E->view_pos = 0; E->need_fresh = false;
  E->B.bind(E->S, E->B.mut.host.data(), E->B.imm.host.data());
  E->B.bind_backfill(E->Sbf, E->B.mut.host.data(), E->B.imm.host.data());
  E->cur = &E->S;
  E->launches = 0;
  return KB_OK;
}
int kbemu_run(void* h, uint32_t actions) {
  Emu* E = (Emu*)h;
  std::vector<uint64_t> buf(kbemu_buf_u64(E));
  const uint64_t guard = 4ull * ((uint64_t)E->B.J + E->B.To + E->B.Tb) + 1024;
  for (uint32_t pass = 0; pass < 2; ++pass) {
    if (!((actions ? actions : 1u) & (1u << pass))) continue;
    if (pass == 1) emu_begin_backfill(*E, ((actions ? actions : 1u) & 1u) != 0);
    while (!E->cur->ctl->done) {
      if (E->pipe && !E->cur->backfill) emu_launch_pipe(*E);
      else if (E->cur->overlap) emu_launch_overlap(*E);
      else if (E->cur->kchain > 1) emu_launch_chain(*E);
      else { emu_scan(*E, buf.data()); emu_replay(*E, buf.data()); }
      if (E->launches > guard) { g_err = "emulated cycle did not terminate"; return KB_E_STATE; }
    }
  }
  return KB_OK;
}
void kbemu_destroy(void* h) { delete (Emu*)h; }
int kbemu_done(void* h) { return ((Emu*)h)->cur->ctl->done ? 1 : 0; }
void kbemu_begin_backfill(void* h, int allocate_ran) { emu_begin_backfill(*(Emu*)h, allocate_ran != 0); }
uint32_t kbemu_buf_u64(void* h) { return xchg_u64(((Emu*)h)->S.ncols); }
void kbemu_scan(void* h, uint64_t* sendbuf) { emu_scan(*(Emu*)h, sendbuf); }
void kbemu_replay(void* h, const uint64_t* recvbuf) { emu_replay(*(Emu*)h, recvbuf); }

int kbemu_finish(void* h, kb_decision* out, kb_stats* stats,
                 double* node_idle, double* node_releasing, double* node_used, int32_t* node_pods,
                 int64_t* node_nz_cpu, int64_t* node_nz_mem, uint64_t* node_ports,
                 double* job_share, int32_t* job_ready, double* queue_share, double* queue_deserved, double* queue_allocated) {
  Emu& E = *(Emu*)h;
  const DevSession& S = E.S;
  const BuiltSession& B = E.B;
  emu_gang_commit(S, E.Sbf, (const int32_t*)(B.imm.host.data() + B.oi.job_ready0));
  const uint32_t R = B.R, W = B.W, N = B.N, T = B.T, J = B.J, Q = B.Q;
  if (out) memcpy(out, S.dec, (size_t)T * sizeof(kb_decision));
  const size_t tile_u64 = (size_t)B.ncols * TILE_NODES;
  for (uint32_t n = 0; n < N; ++n) {
    TileAcc a{S.tiles + (size_t)(n / TILE_NODES) * tile_u64, n % TILE_NODES, R, W};
    for (uint32_t r = 0; r < R; ++r) {
      if (node_idle) node_idle[(size_t)r * N + n] = a.idle(r);
      if (node_releasing) node_releasing[(size_t)r * N + n] = a.rel(r);
      if (node_used) node_used[(size_t)r * N + n] = S.node_used[(size_t)r * N + n];
    }
    if (node_pods) node_pods[n] = a.pods();
    if (node_nz_cpu) node_nz_cpu[n] = a.nz_cpu();
    if (node_nz_mem) node_nz_mem[n] = a.nz_mem();
    if (node_ports) for (uint32_t w = 0; w < W; ++w) node_ports[(size_t)w * N + n] = a.ports(w) & ~E.B.aff_atom_mask[w];
  }
  for (uint32_t j = 0; j < J; ++j) { if (job_share) job_share[j] = S.job_share[j]; if (job_ready) job_ready[j] = S.job_ready[j]; }
  for (uint32_t q = 0; q < Q; ++q) {
    if (queue_share) queue_share[q] = S.q_share[q];
    for (uint32_t r = 0; r < R; ++r) {
      if (queue_deserved) queue_deserved[(size_t)r * Q + q] = S.q_deserved[(size_t)r * Q + q];
      if (queue_allocated) queue_allocated[(size_t)r * Q + q] = S.q_allocated[(size_t)r * Q + q];
    }
  }
  if (stats) {
    const Ctl& c = *E.cur->ctl;
    memset(stats, 0, sizeof *stats);
    stats->pairs_logical = c.pairs_logical; stats->pairs_scanned = c.pairs_scanned; stats->pairs_replayed = c.pairs_replayed;
    stats->tasks_processed = c.tasks_processed; stats->tasks_allocated = c.tasks_allocated; stats->tasks_pipelined = c.tasks_pipelined;
    stats->visits = c.visits; stats->kernel_launches = E.launches; stats->n_classes = B.C;
    stats->scans = c.scans; stats->rescans = c.rescans; stats->predictions = c.predictions; stats->mispredictions = c.mispredictions;
    stats->chain_hits = c.chain_hits;
    uint32_t jr = 0;
    for (uint32_t j = 0; j < J; ++j) if (S.job_placed[j] && ssn_job_ready(S, j)) ++jr;
    stats->jobs_ready = jr;
  }
  return KB_OK;
}

// single-rank convenience: the whole cycle
int kbemu_allocate(const kb_snapshot* snap, const kb_plugin_conf* conf, uint32_t actions /* bit0 allocate, bit1 backfill */,
                   uint32_t mode /* kbemu_create2 */, kb_decision* out, kb_stats* stats,
                   double* node_idle, double* node_releasing, double* node_used, int32_t* node_pods,
                   int64_t* node_nz_cpu, int64_t* node_nz_mem, uint64_t* node_ports,
                   double* job_share, int32_t* job_ready, double* queue_share, double* queue_deserved, double* queue_allocated) {
  Emu* E = (Emu*)kbemu_create2(snap, conf, 0, 1, mode);
  if (!E) return KB_E_BADARG;
  std::vector<uint64_t> buf(kbemu_buf_u64(E));
  const uint64_t guard = 4ull * ((uint64_t)E->B.J + E->B.To + E->B.Tb) + 1024;
  for (uint32_t pass = 0; pass < 2; ++pass) {
    if (!((actions ? actions : 1u) & (1u << pass))) continue;
    if (pass == 1) emu_begin_backfill(*E, ((actions ? actions : 1u) & 1u) != 0);
    while (!E->cur->ctl->done) {
      if (E->pipe && !E->cur->backfill) emu_launch_pipe(*E);
      else if (E->cur->overlap) emu_launch_overlap(*E);
      else if (E->cur->kchain > 1) emu_launch_chain(*E);
      else { emu_scan(*E, buf.data()); emu_replay(*E, buf.data()); }
      if (E->launches > guard) { g_err = "emulated cycle did not terminate"; delete E; return KB_E_STATE; }
    }
  }
  int rc = kbemu_finish(E, out, stats, node_idle, node_releasing, node_used, node_pods, node_nz_cpu, node_nz_mem, node_ports,
                        job_share, job_ready, queue_share, queue_deserved, queue_allocated);
  delete E;
  return rc;
}

// scheduler.go:88-101 on the emulation: the action list on ONE session (kb_cycle).  actions: 0 reclaim, 1 allocate, 2 backfill, 3 preempt.
int kbemu_cycle(const kb_snapshot* snap, const kb_running* running, const kb_plugin_conf* conf, const uint8_t* actions, uint32_t n_actions, uint32_t mode,
                kb_decision* out, uint8_t* evicted, uint32_t* evict_order, kb_stats* stats,
                double* node_idle, double* node_releasing, double* node_used, int32_t* node_pods,
                int64_t* node_nz_cpu, int64_t* node_nz_mem, uint64_t* node_ports,
                double* job_share, int32_t* job_ready, double* queue_share, double* queue_deserved, double* queue_allocated) {
  for (uint32_t i = 0; i + 1 < n_actions; ++i)
    if (actions[i] == 3 && actions[i + 1] != 3) { g_err = "an action after preempt is outside this build (kb_cycle)"; return KB_E_UNSUPPORTED_FEATURE; }
  Emu* E = (Emu*)kbemu_create2(snap, conf, 0, 1, mode);
  if (!E) return KB_E_BADARG;
  if (E->B.aff_session && !E->B.aff_evict_ok && running && running->n)
    for (uint32_t i = 0; i < n_actions; ++i)
      if (actions[i] == 0 || actions[i] == 3) { g_err = "reclaim / preempt in this session with inter-pod affinity are outside this build"; delete E; return KB_E_UNSUPPORTED_FEATURE; }
  EvictBuilt EB;
  BuildErr be;
  if (int rc = build_evict(snap, running, E->B, E->S, EB, &be)) { g_err = be.msg; delete E; return rc; }
  EvictDev D{};
  EB.bind(D, EB.imm.host.data(), EB.mut.host.data());
  CpuExec x;
  std::vector<uint64_t> buf(kbemu_buf_u64(E));
  const uint64_t guard = 4ull * ((uint64_t)E->B.J + E->B.To + E->B.Tb) + 1024;
  const uint32_t J = E->B.J;
  std::vector<int32_t> ready_start((const int32_t*)(E->B.imm.host.data() + E->B.oi.job_ready0), (const int32_t*)(E->B.imm.host.data() + E->B.oi.job_ready0) + J);
  bool dirty = false, ready_taken = false, have_latest = false, alloc_ran = false;
  uint32_t latest = 0;
  auto pump = [&]() -> int {
    while (!E->cur->ctl->done) {
      if (E->pipe && !E->cur->backfill) emu_launch_pipe(*E);
      else if (E->cur->overlap) emu_launch_overlap(*E);
      else if (E->cur->kchain > 1) emu_launch_chain(*E);
      else { emu_scan(*E, buf.data()); emu_replay(*E, buf.data()); }
      if (E->launches > guard) { g_err = "emulated cycle did not terminate"; return KB_E_STATE; }
    }
    return KB_OK;
  };
  for (uint32_t i = 0; i < n_actions; ++i) {
    const uint8_t a = actions[i];
    // Ctl.pred_dead: backfill left a task Allocated on no node, ssn.PredicateFn fails for every pair from then on — the
    // remaining actions find no node for anybody (kb_cycle skips them the same way)
    if (E->Sbf.ctl->pred_dead) continue;
    if (a == 0 || a == 3) {
      if (have_latest) D.ctl->step = latest;
      if (a == 3) run_preempt(x, E->S, D); else run_reclaim(x, E->S, D);
      if (D.ctl->error == 3) { g_err = "a member of an inter-pod affinity counter group was evicted: the outcome of this cycle is withheld"; delete E; return KB_E_UNSUPPORTED_FEATURE; }
      if (D.ctl->error) { g_err = D.ctl->error == 2 ? "victim overflow" : "the reference would panic (Resource.Sub)"; delete E; return KB_E_STATE; }
      latest = D.ctl->step; have_latest = true; dirty = true;
    } else {
      if (!ready_taken) { for (uint32_t j = 0; j < J; ++j) ready_start[j] = E->S.job_ready[j]; ready_taken = true; }
      if (a == 1) {
        if (dirty || have_latest) { prep_task_lists(E->S); prep_allocate(E->S, *E->S.ctl, have_latest ? latest : 0u); }
        E->cur = &E->S;
        if (int rc = pump()) { delete E; return rc; }
        latest = E->S.ctl->step; alloc_ran = true;
      } else {
        if (dirty) { prep_backfill(E->Sbf, *E->Sbf.ctl, have_latest ? latest : 0u); E->cur = &E->Sbf; }
        else emu_begin_backfill(*E, alloc_ran);
        if (int rc = pump()) { delete E; return rc; }
        latest = E->Sbf.ctl->step;
      }
      have_latest = true;
    }
  }
  const uint32_t n = EB.n_run;
  for (uint32_t k = 0; k < n; ++k) {
    const uint32_t i = D.r_orig[k];
    if (evicted) evicted[i] = D.evict_order[k] != 0xFFFFFFFFu ? 1 : 0;
    if (evict_order) evict_order[i] = D.evict_order[k];
  }
  memcpy(E->B.imm.host.data() + E->B.oi.job_ready0, ready_start.data(), (size_t)J * 4);      // what the gang commit counts from
  int rc = kbemu_finish(E, out, stats, node_idle, node_releasing, node_used, node_pods, node_nz_cpu, node_nz_mem, node_ports,
                        job_share, job_ready, queue_share, queue_deserved, queue_allocated);
  if (stats) { stats->evictions = D.ctl->n_evicted; stats->evict_sweeps = D.ctl->scans; stats->tasks_pipelined += D.ctl->n_pipelined; stats->tasks_processed += D.ctl->tasks_processed; }
  delete E;
  return rc;
}

// reclaim (action 0) / preempt (action 1) from the as-loaded state: the product's kb_evict.h run by one CPU thread
int kbemu_evict(const kb_snapshot* snap, const kb_running* running, const kb_plugin_conf* conf, uint32_t action,
                kb_decision* out, uint8_t* evicted, uint32_t* evict_order, kb_stats* stats,
                double* node_idle, double* node_releasing, double* node_used, int32_t* node_pods,
                int64_t* node_nz_cpu, int64_t* node_nz_mem, uint64_t* node_ports,
                double* job_share, int32_t* job_ready, double* queue_share, double* queue_deserved, double* queue_allocated) {
  Emu* E = (Emu*)kbemu_create2(snap, conf, 0, 1, 1);
  if (!E) return KB_E_BADARG;
  EvictBuilt EB;
  BuildErr be;
  if (int rc = build_evict(snap, running, E->B, E->S, EB, &be)) { g_err = be.msg; delete E; return rc; }
  EvictDev D{};
  EB.bind(D, EB.imm.host.data(), EB.mut.host.data());
  CpuExec x;
  if (action) run_preempt(x, E->S, D); else run_reclaim(x, E->S, D);
  const EvictCtl& ctl = *D.ctl;
  if (ctl.error == 3) { g_err = "a member of an inter-pod affinity counter group was evicted: the outcome of this action is withheld"; delete E; return KB_E_UNSUPPORTED_FEATURE; }
  if (ctl.error) { g_err = ctl.error == 2 ? "victim overflow" : "the reference would panic (Resource.Sub)"; delete E; return KB_E_STATE; }
  const uint32_t n = EB.n_run, T = E->B.T;
  for (uint32_t k = 0; k < n; ++k) {
    const uint32_t i = D.r_orig[k];
    if (evicted) evicted[i] = D.evict_order[k] != 0xFFFFFFFFu ? 1 : 0;
    if (evict_order) evict_order[i] = D.evict_order[k];
  }
  const BuiltSession& B = E->B; const DevSession& S = E->S;
  const uint32_t R = B.R, W = B.W, N = B.N, J = B.J, Q = B.Q;
  if (out) memcpy(out, S.dec, (size_t)T * sizeof(kb_decision));
  const size_t tile_u64 = (size_t)B.ncols * TILE_NODES;
  for (uint32_t nn = 0; nn < N; ++nn) {
    TileAcc a{S.tiles + (size_t)(nn / TILE_NODES) * tile_u64, nn % TILE_NODES, R, W};
    for (uint32_t r = 0; r < R; ++r) {
      if (node_idle) node_idle[(size_t)r * N + nn] = a.idle(r);
      if (node_releasing) node_releasing[(size_t)r * N + nn] = a.rel(r);
      if (node_used) node_used[(size_t)r * N + nn] = S.node_used[(size_t)r * N + nn];
    }
    if (node_pods) node_pods[nn] = a.pods();
    if (node_nz_cpu) node_nz_cpu[nn] = a.nz_cpu();
    if (node_nz_mem) node_nz_mem[nn] = a.nz_mem();
    if (node_ports) for (uint32_t w = 0; w < W; ++w) node_ports[(size_t)w * N + nn] = a.ports(w);
  }
  for (uint32_t j = 0; j < J; ++j) { if (job_share) job_share[j] = S.job_share[j]; if (job_ready) job_ready[j] = S.job_ready[j]; }
  for (uint32_t q = 0; q < Q; ++q) {
    if (queue_share) queue_share[q] = S.q_share[q];
    for (uint32_t r = 0; r < R; ++r) {
      if (queue_deserved) queue_deserved[(size_t)r * Q + q] = S.q_deserved[(size_t)r * Q + q];
      if (queue_allocated) queue_allocated[(size_t)r * Q + q] = S.q_allocated[(size_t)r * Q + q];
    }
  }
  if (stats) {
    memset(stats, 0, sizeof *stats);
    stats->pairs_logical = ctl.pairs_logical; stats->pairs_scanned = (uint64_t)ctl.scans * N;
    stats->tasks_processed = ctl.tasks_processed; stats->tasks_pipelined = ctl.n_pipelined;
    stats->evictions = ctl.n_evicted; stats->evict_sweeps = ctl.scans; stats->n_classes = B.C;
  }
  delete E;
  return KB_OK;
}

}  // extern "C"
