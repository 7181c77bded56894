// pm_nlcc.cuh — non-local constraint checking (NLCC) as a batched GPU token queue.
//
// Replaces the reference's asynchronous token visitors
//   nem_1  tppm_visitor      (token_passing_pattern_matching_nonunique_nem_1.hpp:98-303, 311-861)
//   TDS    tppm_visitor_tds  (token_passing_pattern_matching_nonunique_tds_batch_1.hpp:122-335, 347-919)
// and the driver's post-processing of token_source_map (src/run_pattern_matching_beta.cpp:956-1062).
// Paths relative to /root/reference (headers under include/havoqgt/).
//
// Formulation.  Nothing a token reads (vertex_active, template_vertices,
// vertex_active_edges_map keys) is written while a constraint runs, so the walk
// can be advanced one hop per kernel: level h of the token pool holds the tokens
// ACCEPTED at hop h; k_*_expand walks E_v of every token's vertex v and applies
// the receiver's pre_visit tests of hop h+1 at emission time, so only accepted
// tokens are ever stored.
//   nem_1: token = (vertex, source); per-(vertex, source) work aggregation
//          (vertex_token_source_set, nem_1.hpp:131-139, 270-285) is an
//          open-addressing device hash set.  The final hop stores nothing: a path
//          constraint acknowledges the source (ok[s] = 1, nem_1.hpp:683-727,
//          326-342), a cycle constraint additionally flags the edge the token
//          came back on (nem_1.hpp:764-770) in bit 31 of the working adjacency.
//          The reference never forwards a token to the parent it came from
//          (nem_1.hpp:836-838); that parent is always rejected downstream when
//          interior hop labels are pairwise distinct and P[h-1] != P[h+1]
//          (SURVEY A.6 #7), which pm_pattern_load_dir verifies per constraint —
//          otherwise the reference result is arrival-order dependent.
//   TDS:   aggregation off (enable_vertex_token_source_cache = false,
//          tds_batch_1.hpp:11); token = (index of parent token, vertex), the
//          visited history (tds_batch_1.hpp:964) is recovered by walking parent
//          links, and the enumeration-index rule (tds_batch_1.hpp:284-302,
//          622-639, 808-886) is tested against it.  Completed walks are appended
//          to a match list and materialised for the subgraph files on request.
#pragma once

#include "pm_common.cuh"
#include "pm_lcc.cuh"

namespace pm {

__constant__ NlcConst c_nlc;

#define PM_HSET_EMPTY 0xFFFFFFFFFFFFFFFFull

// every vertex below is named by its COMPACT ID (see pm_lcc.cuh); `rowblk` is the row start by local compact id
// A neighbour u can take hop h iff bit I[h] of S[u] is set: S[u] is always a subset of labelmask(label[u]),
// so the bit test implies the label test of the reference (nem_1.hpp:557-581) and no class gather is needed.
struct NlcArgs {
  const uint32_t* rowblk;
  uint32_t* colw;
  const uint16_t* S;
  const uint32_t* adeg;
  const uint8_t* cls;
  uint8_t* ok;
  uint32_t* src_list;
  unsigned long long* hset;
  uint64_t hset_mask;
  uint2* pool;
  uint64_t pool_cap;
  uint2* matches;      // TDS: (token index at the penultimate level, final vertex)
  uint64_t match_cap;
  DevCounters* cnt;
  uint32_t base;       // first compact id of this rank (rank-local arrays rowblk / adeg: index cid - base)
  const uint32_t* vid; // compact id -> slot
  int par;             // multi-GPU: token inbox written in this hop (the previous hop's is par ^ 1)
  const StepMsg* all;  // multi-GPU: everyone's StepMsg of the previous hop
};

__device__ __forceinline__ bool hop_ok(uint32_t su, uint32_t cu, int h) {
  // active + label + template bit of hop h (nem_1.hpp:101,186-210; tds_batch_1.hpp:125,207-233)
  return cu == c_nlc.cls[h] && ((su >> c_nlc.I[h]) & 1u);
}

__device__ __forceinline__ unsigned long long mix64(unsigned long long x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
  return x;
}

// true iff (u, s) was not in the set before
__device__ __forceinline__ bool hset_insert(const NlcArgs& a, uint32_t u, uint32_t s) {
  const unsigned long long key = ((unsigned long long)u << 32) | s;
  uint64_t h = mix64(key) & a.hset_mask;
  for (int probe = 0; probe < 256; ++probe) {
    const unsigned long long prev = atomicCAS(&a.hset[h], PM_HSET_EMPTY, key);
    if (prev == PM_HSET_EMPTY) return true;
    if (prev == key) return false;
    h = (h + 1) & a.hset_mask;
  }
  a.cnt->overflow = 1u;
  return false;
}

// Reserves `n` consecutive slots per lane behind *counter with one atomic per warp;
// returns this lane's first slot.  All 32 lanes must call it.
__device__ __forceinline__ unsigned long long warp_reserve(unsigned long long* counter, uint32_t n) {
  const uint32_t lane = threadIdx.x & 31;
  uint32_t incl = n;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= (uint32_t)o) incl += t;
  }
  const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
  unsigned long long base = 0;
  if (total) {
    if (lane == 31) base = atomicAdd(counter, (unsigned long long)total);
    base = __shfl_sync(0xffffffffu, base, 31);
  }
  return base + incl - n;
}


// ---------------------------------------------------------------------------
// Work aggregation for token emission.  A returning atomic on ONE address retires
// about one per clock chip-wide, so reserving pool slots per warp and pass would
// serialise the expand kernels on the pool counter.  Each warp instead stages its
// tokens in a private shared-memory buffer and reserves pool slots once per
// PM_STAGE_FLUSH or more tokens (one atomic, then coalesced 8-byte stores).
// ---------------------------------------------------------------------------
#define PM_STAGE_FLUSH 64u
#define PM_STAGE_CAP (PM_STAGE_FLUSH + 128u)  // a pass appends at most 32 lanes x 4 slots

struct WarpStage {
  uint2* buf;     // this warp's PM_STAGE_CAP entries
  uint32_t fill;  // same value in every lane
};

__device__ __forceinline__ void stage_flush(WarpStage& w, uint2* __restrict__ dst, unsigned long long cap,
                                            unsigned long long* counter, uint32_t* overflow) {
  if (w.fill == 0) return;
  const uint32_t lane = threadIdx.x & 31;
  unsigned long long base = 0;
  if (lane == 0) base = atomicAdd(counter, (unsigned long long)w.fill);
  base = __shfl_sync(0xffffffffu, base, 0);
  __syncwarp();
  for (uint32_t i = lane; i < w.fill; i += 32) {
    if (dst && base + i < cap) dst[base + i] = w.buf[i];
    else if (dst) *overflow = 1u;
  }
  __syncwarp();
  w.fill = 0;
}

// all 32 lanes call; each lane appends its flagged values (order: lane, then k)
__device__ __forceinline__ void stage_push(WarpStage& w, const bool (&flag)[4], const uint2 (&val)[4],
                                           uint2* __restrict__ dst, unsigned long long cap,
                                           unsigned long long* counter, uint32_t* overflow) {
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t n = (uint32_t)flag[0] + flag[1] + flag[2] + flag[3];
  uint32_t incl = n;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= (uint32_t)o) incl += t;
  }
  const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
  if (total == 0) return;
  uint32_t pos = w.fill + incl - n;
#pragma unroll
  for (int k = 0; k < 4; ++k)
    if (flag[k]) w.buf[pos++] = val[k];
  w.fill += total;
  if (w.fill > PM_STAGE_FLUSH) stage_flush(w, dst, cap, counter, overflow);
}

// all 32 lanes call; a lane appends at most one value (order: lane)
__device__ __forceinline__ void stage_push1(WarpStage& w, bool flag, uint2 val, uint2* __restrict__ dst,
                                            unsigned long long cap, unsigned long long* counter, uint32_t* overflow) {
  const uint32_t m = __ballot_sync(0xffffffffu, flag);
  if (m == 0u) return;
  if (flag) w.buf[w.fill + __popc(m & lanemask_lt())] = val;
  w.fill += __popc(m);
  if (w.fill > PM_STAGE_FLUSH) stage_flush(w, dst, cap, counter, overflow);
}

// ---------------------------------------------------------------------------
// Dealing token rows to lanes.  A warp takes 32 tokens (one per lane), lays the rows E_v of their vertices end
// to end and deals the SLOTS to its lanes, 32 per pass: every lane of every pass carries one neighbour whatever
// the row lengths are (rows of pruned graphs hold a handful of slots: a fixed group of lanes per token idles most
// of them).  Token parameters travel through shared memory; a lane finds its token by a 5-step search over the
// inclusive prefix of the row lengths.
// ---------------------------------------------------------------------------
struct TokBatch {
  uint32_t (*cum)[32];  // [warp][lane] inclusive prefix of the row lengths
  uint4 (*tok)[32];     // [warp][lane] {vertex, source, row start (sectors), row length}
};

// returns the number of slots of the batch; all 32 lanes call
__device__ __forceinline__ uint32_t deal_begin(const TokBatch& b, uint32_t wid, uint32_t lane, uint32_t v, uint32_t s,
                                               uint32_t row, uint32_t d) {
  uint32_t cum = d;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, cum, o);
    if (lane >= (uint32_t)o) cum += t;
  }
  __syncwarp();
  b.cum[wid][lane] = cum;
  b.tok[wid][lane] = make_uint4(v, s, row, d);
  __syncwarp();
  return __shfl_sync(0xffffffffu, cum, 31);
}

// slot g of the batch (g < total): its token and the position inside the token's row
__device__ __forceinline__ uint4 deal_slot(const TokBatch& b, uint32_t wid, uint32_t g, uint32_t& j) {
  uint32_t idx = 0;  // number of tokens whose rows end at or before g
#pragma unroll
  for (int k = 16; k; k >>= 1)
    if (b.cum[wid][idx + k - 1] <= g) idx += k;
  const uint4 t = b.tok[wid][idx];
  j = g - (b.cum[wid][idx] - t.w);
  return t;
}

// ---------------------------------------------------------------------------
// token sources (nem_1.hpp:387-527; tds_batch_1.hpp:1067-1135, 425-512)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock) k_nlcc_sources(NlcArgs a, const uint4* __restrict__ l0,
                                                          const uint4* __restrict__ l1, int cur, int tds) {
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t c0 = a.cnt->fr_n[cur][0], c1 = a.cnt->fr_n[cur][1];
  const uint32_t total = c0 + c1;
  const bool multi = c_peer.G > 1;
  uint32_t i0 = (blockIdx.x * blockDim.x + threadIdx.x) & ~31u;
  for (; i0 < total; i0 += gridDim.x * blockDim.x) {
    const uint32_t i = i0 + lane;
    bool is_src = false;
    uint32_t v = 0;
    if (i < total) {
      const uint4 e = i < c0 ? l0[i] : l1[i - c0];
      v = e.x;  // compact id
      const uint32_t T = e.y != PM_TOMB ? (uint32_t)a.S[v] : 0u;
      is_src = T != 0 && hop_ok(T, a.cls[v], 0);
      // path checking starts only from vertices that match BOTH end points (nem_1.hpp:447-451)
      if (is_src && !tds && !c_nlc.valid_cycle) is_src = (T >> c_nlc.I[c_nlc.n - 1]) & 1u;
    }
    const uint32_t m = __ballot_sync(0xffffffffu, is_src);
    if (m) {
      uint32_t base = 0;
      if (lane == 0) base = atomicAdd(&a.cnt->n_src, (uint32_t)__popc(m));
      base = __shfl_sync(0xffffffffu, base, 0);
      if (is_src) {
        const uint32_t pos = base + __popc(m & lanemask_lt());
        a.src_list[pos] = v;
        a.ok[v] = 0;
        if (!multi) {
          // level 0 of the token pool (pm_nlcc sizes the pool for at least the vertices still in the map; the
          // guard covers a pool sized by an earlier, smaller search)
          if (pos < a.pool_cap) a.pool[pos] = tds ? make_uint2(0xFFFFFFFFu, v) : make_uint2(v, v);
          else a.cnt->overflow = 1u;
        } else if (!tds) {
          // level 0 = my own region of my token inbox
          if (pos < c_peer.tcap) c_peer.tin[a.par][c_peer.rank][(unsigned long long)c_peer.rank * c_peer.tcap + pos] = make_uint2(v, v);
          else a.cnt->overflow = 1u;
        } else {
          const unsigned long long n = (unsigned long long)c_nlc.n;
          if (pos < c_peer.tcap * 2ull / n)
            reinterpret_cast<uint32_t*>(c_peer.tin[a.par][c_peer.rank])[(unsigned long long)c_peer.rank * c_peer.tcap * 2ull + pos * n] = v;
          else a.cnt->overflow = 1u;
        }
      }
    }
  }
}

// flag the edge (s -> parent) a successful cycle token came back on (nem_1.hpp:764-770)
__device__ __forceinline__ void mark_edge(const NlcArgs& a, uint32_t s, uint32_t parent) {
  const uint64_t row = (uint64_t)a.rowblk[s] * 8;
  uint32_t lo = 0, hi = a.adeg[s];
  while (lo < hi) {  // rows stay ascending: compaction is stable
    const uint32_t mid = (lo + hi) >> 1;
    const uint32_t x = a.colw[row + mid] & PM_IDMASK;
    if (x < parent) lo = mid + 1; else hi = mid;
  }
  if (lo < a.adeg[s] && (a.colw[row + lo] & PM_IDMASK) == parent) atomicOr(&a.colw[row + lo], 0x80000000u);
}

// level bookkeeping on the device (no host round trip per hop): level 0 = the sources
__global__ void k_nlcc_begin(DevCounters* cnt) {
  cnt->lvl[0] = 0;
  cnt->lvl[1] = cnt->n_src;
  cnt->pool_n = cnt->n_src;
  if (c_peer.G > 1) {  // the sources are level 0 of my own inbox region
    cnt->pool_n = 0;
    cnt->out_n[c_peer.rank] = cnt->n_src;
  }
}
// after the expand kernel that produced level h
__global__ void k_nlcc_close_level(DevCounters* cnt, int h, unsigned long long pool_cap) {
  const unsigned long long n = cnt->pool_n;
  cnt->lvl[h + 1] = n < pool_cap ? n : pool_cap;
}

// ---------------------------------------------------------------------------
// nem_1, final hop of a CYCLE constraint (max_itr_count == itr_count,
// nem_1.hpp:661-773): the token at v (hop C) would be forwarded along E_v and only
// the copy arriving at the source s can succeed, so instead of walking E_v the row
// is binary-searched for s.  Success acknowledges the source and flags the edge
// E_s[v] the token came back on (nem_1.hpp:764-770).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock) k_nem1_final_cycle(NlcArgs a, int hlevel, int hn) {
  const unsigned long long lo = a.cnt->lvl[hlevel], hi = a.cnt->lvl[hlevel + 1];
  unsigned long long fan = 0;
  for (unsigned long long t = lo + (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; t < hi;
       t += (unsigned long long)gridDim.x * blockDim.x) {
    const uint2 tk = a.pool[t];
    const uint32_t v = tk.x, s = tk.y;
    const uint32_t ss = a.S[s];
    if (ss == 0 || !hop_ok(ss, a.cls[s], hn)) continue;  // receiver tests at the source (nem_1.hpp:557-581)
    const uint64_t row = (uint64_t)a.rowblk[v] * 8;
    uint32_t b = 0, e = a.adeg[v];
    fan += e;
    while (b < e) {  // rows stay ascending: compaction is stable
      const uint32_t mid = (b + e) >> 1;
      const uint32_t x = a.colw[row + mid] & PM_IDMASK;
      if (x < s) b = mid + 1; else e = mid;
    }
    if (b < a.adeg[v] && (a.colw[row + b] & PM_IDMASK) == s) {
      a.ok[s] = 1;
      a.cnt->found = 1u;
      mark_edge(a, s, v);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) fan += __shfl_xor_sync(0xffffffffu, fan, o);
  if ((threadIdx.x & 31) == 0 && fan) atomicAdd(&a.cnt->fanout, fan);
}

// ---------------------------------------------------------------------------
// nem_1: advance the tokens of level hlevel (accepted at hop hn-1) to hop hn
// ---------------------------------------------------------------------------
// dedupe: apply the (vertex, source) aggregation.  Off where duplicates cannot occur (hop 1: the neighbours of
// a source are distinct) or cannot matter (the level that feeds k_nem1_close_cycle, whose effects are idempotent).
template <bool FINAL>
__global__ void __launch_bounds__(kBlock) k_nem1_expand(NlcArgs a, int hlevel, int hn, int dedupe) {
  __shared__ uint2 s_stage[FINAL ? 1 : (kBlock / 32) * PM_STAGE_CAP];
  __shared__ uint32_t s_cum[kBlock / 32][32];
  __shared__ uint4 s_tok[kBlock / 32][32];
  const TokBatch tb{s_cum, s_tok};
  WarpStage stage{s_stage + (FINAL ? 0 : (threadIdx.x >> 5) * PM_STAGE_CAP), 0u};
  const uint64_t lo = a.cnt->lvl[hlevel], hi = a.cnt->lvl[hlevel + 1];
  const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  const uint32_t ibit = c_nlc.I[hn];
  unsigned long long fan = 0;
  for (uint64_t base = lo + warp * 32; base < hi; base += nwarps * 32) {
    const uint64_t t = base + lane;
    uint32_t v = 0, s = 0, d = 0, row = 0;
    if (t < hi) {
      const uint2 tk = a.pool[t];
      v = tk.x;
      s = tk.y;
      d = a.adeg[v];
      row = a.rowblk[v];
      // a path constraint needs ONE completed walk per source (ack_success just sets
      // token_source_map[s] = 1, nem_1.hpp:326-342): later tokens of an acknowledged source are moot
      if (FINAL && a.ok[s]) d = 0;
    }
    fan += d;
    const uint32_t total = deal_begin(tb, wid, lane, v, s, row, d);
    for (uint32_t g0 = 0; g0 < total; g0 += 32) {
      const uint32_t g = g0 + lane;
      bool pass = false;
      uint32_t u = 0, ts = 0, tv = 0;
      if (g < total) {
        uint32_t j;
        const uint4 tk = deal_slot(tb, wid, g, j);
        tv = tk.x;
        ts = tk.y;
        u = a.colw[(uint64_t)tk.z * 8 + j] & PM_IDMASK;
        const uint32_t su = a.S[u];
        pass = su != 0 && ((su >> ibit) & 1u);
      }
      if (FINAL) {
        // max_itr_count == itr_count (nem_1.hpp:661-791)
        if (pass && (c_nlc.valid_cycle ? u == ts : u != ts)) {
          a.ok[ts] = 1;
          a.cnt->found = 1u;
          if (c_nlc.valid_cycle) mark_edge(a, ts, tv);
        }
      } else {
        // interior hop: the source cannot relay (nem_1.hpp:174-177), one token per
        // (vertex, source) (nem_1.hpp:131-139, 270-285)
        bool ins = pass && u != ts;
        if (ins && dedupe) ins = hset_insert(a, u, ts);
        stage_push1(stage, ins, make_uint2(u, ts), a.pool, a.pool_cap, &a.cnt->pool_n, &a.cnt->overflow);
      }
    }
  }
  if (!FINAL) stage_flush(stage, a.pool, a.pool_cap, &a.cnt->pool_n, &a.cnt->overflow);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) fan += __shfl_xor_sync(0xffffffffu, fan, o);
  if (lane == 0 && fan) atomicAdd(&a.cnt->fanout, fan);
}

// ---------------------------------------------------------------------------
// nem_1, the last TWO hops of a CYCLE constraint in one kernel.  A token (v, s)
// accepted at hop C-1 would be relayed to every u in E_v that passes the tests of
// hop C (the last interior hop), and u would relay it along E_u where only the copy
// arriving at the source can succeed (max_itr_count == itr_count, nem_1.hpp:661-773).
// So the walk closes iff some u passes the hop-C tests and lies in E_v AND E_s
// (edge maps are symmetric between live vertices once an LCC call of >= 2 supersteps
// has run, see pm_lcc.cuh; pm_nlcc checks that precondition): the slots of E_v are dealt
// to the lanes, a neighbour that passes the hop test (one mask gather) is looked up in the
// (short, cached) row of s.  No level-C tokens are stored or deduplicated.
// Success acknowledges the source and flags the edge E_s[u] the token would have
// come back on (nem_1.hpp:764-770).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock) k_nem1_close_cycle(NlcArgs a, int hlevel, int hn) {
  __shared__ uint32_t s_cum[kBlock / 32][32];
  __shared__ uint4 s_tok[kBlock / 32][32];
  __shared__ uint2 s_src[kBlock / 32][32];  // per token: {row start of E_s (sectors), |E_s|}
  const TokBatch tb{s_cum, s_tok};
  const uint64_t lo = a.cnt->lvl[hlevel], hi = a.cnt->lvl[hlevel + 1];
  const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  const uint32_t ibit = c_nlc.I[hn];
  unsigned long long fan = 0;
  for (uint64_t base = lo + warp * 32; base < hi; base += nwarps * 32) {
    const uint64_t t = base + lane;
    uint32_t s = 0, d = 0, row = 0, ds = 0, rs = 0;
    if (t < hi) {
      const uint2 tk = a.pool[t];
      s = tk.y;
      const uint32_t ss = a.S[s];
      // receiver tests of the closing hop at the source (nem_1.hpp:557-581)
      if (ss != 0 && hop_ok(ss, a.cls[s], hn + 1)) {
        d = a.adeg[tk.x];
        row = a.rowblk[tk.x];
        ds = a.adeg[s];
        rs = a.rowblk[s];
      }
    }
    fan += d;
    __syncwarp();
    s_src[wid][lane] = make_uint2(rs, ds);
    const uint32_t total = deal_begin(tb, wid, lane, lane, s, row, d);  // .x carries the token's lane: indexes s_src
    for (uint32_t g0 = 0; g0 < total; g0 += 32) {
      const uint32_t g = g0 + lane;
      if (g >= total) continue;
      uint32_t j;
      const uint4 tk = deal_slot(tb, wid, g, j);
      const uint32_t u = a.colw[(uint64_t)tk.z * 8 + j] & PM_IDMASK;
      if (u == tk.y) continue;  // the source cannot relay (nem_1.hpp:174-177)
      const uint32_t su = a.S[u];
      if (su == 0 || !((su >> ibit) & 1u)) continue;
      const uint2 sr = s_src[wid][tk.x];
      const uint64_t rsrc = (uint64_t)sr.x * 8;
      uint32_t b = 0, e = sr.y;
      while (b < e) {  // rows stay ascending: compaction is stable
        const uint32_t mid = (b + e) >> 1;
        const uint32_t x = a.colw[rsrc + mid] & PM_IDMASK;
        if (x < u) b = mid + 1; else e = mid;
      }
      if (b < sr.y && (a.colw[rsrc + b] & PM_IDMASK) == u) {
        a.ok[tk.y] = 1;
        a.cnt->found = 1u;
        atomicOr(&a.colw[rsrc + b], 0x80000000u);
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) fan += __shfl_xor_sync(0xffffffffu, fan, o);
  if (lane == 0 && fan) atomicAdd(&a.cnt->fanout, fan);
}

// ---------------------------------------------------------------------------
// TDS: advance tokens [lo, hi) of level h = hn-1 to hop hn.
// ---------------------------------------------------------------------------
__device__ __forceinline__ bool hist_rule(const uint32_t (&hist)[16], int hp, uint32_t x) {
  const int e = c_nlc.e[hp];
  bool dup = false, eq = false;
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    if (i < hp && hist[i] == x) dup = true;
    if (i == e && hist[i] == x) eq = true;
  }
  if (e == hp) return !dup;   // a new vertex must differ from everything visited
  if (e < hp) return eq;      // a revisit must equal visited[e]
  return false;               // "invalid value" branches drop the token
}

template <bool FINAL>
__global__ void __launch_bounds__(kBlock) k_tds_expand(NlcArgs a, int hlevel, int hn) {
  __shared__ uint2 s_stage[(kBlock / 32) * PM_STAGE_CAP];
  WarpStage stage{s_stage + (threadIdx.x >> 5) * PM_STAGE_CAP, 0u};
  const uint64_t lo = a.cnt->lvl[hlevel], hi = a.cnt->lvl[hlevel + 1];
  constexpr int GROUP = 8;
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t gl = lane % GROUP, gw = lane / GROUP;
  const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  const int h = hn - 1;
  unsigned long long fan = 0;
  for (uint64_t base = lo + warp * 4; base < hi; base += nwarps * 4) {
    const uint64_t t = base + gw;
    const bool has = t < hi;
    uint32_t hist[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) hist[i] = 0xFFFFFFFFu;
    uint32_t v = 0, d = 0;
    if (has) {
      uint2 tk = a.pool[t];
      v = tk.y;
      // walk the parent links: hist[h] = v, hist[h-1] = parent's vertex, ...
#pragma unroll
      for (int i = 15; i >= 0; --i) {
        if (i == h) hist[i] = v;
        if (i < h) {
          tk = a.pool[tk.x];
          hist[i] = tk.y;
        }
      }
      d = a.adeg[v];
    }
    const uint32_t s = hist[0];
    const uint64_t row = has ? (uint64_t)a.rowblk[v] * 8 : 0;
    const uint32_t passes = (d + GROUP * 4 - 1) / (GROUP * 4);
    const uint32_t maxp = __reduce_max_sync(0xffffffffu, passes);
    for (uint32_t p = 0; p < maxp; ++p) {
      const uint32_t j0 = p * GROUP * 4 + gl * 4;
      uint4 q = make_uint4(0, 0, 0, 0);
      if (j0 < d) {
        q = *reinterpret_cast<const uint4*>(a.colw + row + j0);
      }
      const uint32_t u[4] = {q.x & PM_IDMASK, q.y & PM_IDMASK, q.z & PM_IDMASK, q.w & PM_IDMASK};
      bool acc[4];
      uint2 tok[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        acc[k] = false;
        tok[k] = make_uint2((uint32_t)t, u[k]);
        bool may = j0 + k < d;
        if (may) {
          const uint32_t su = a.S[u[k]];
          acc[k] = su != 0 && ((su >> c_nlc.I[hn]) & 1u);
          if (acc[k]) {
            if (FINAL)  // penultimate-hop filter of the sender (tds_batch_1.hpp:808-845)
              acc[k] = c_nlc.valid_cycle ? (u[k] == s) : (u[k] != s && hist_rule(hist, hn, u[k]));
            else        // tds_batch_1.hpp:284-302 (receiver) == :846-886 (sender)
              acc[k] = hist_rule(hist, hn, u[k]);
          }
        }
        if (FINAL && acc[k]) {  // walk completed (tds_batch_1.hpp:664-694, 699-750)
          a.ok[s] = 1;
          a.cnt->found = 1u;
        }
      }
      // completed walks go to the match list (a count only when it is not kept: the host
      // compares `matches` with match_cap itself), relayed tokens to the next pool level
      if (FINAL) stage_push(stage, acc, tok, a.matches, a.match_cap, &a.cnt->matches, &a.cnt->match_drop);
      else stage_push(stage, acc, tok, a.pool, a.pool_cap, &a.cnt->pool_n, &a.cnt->overflow);
    }
    if (has && gl == 0) fan += d;
  }
  if (FINAL) stage_flush(stage, a.matches, a.match_cap, &a.cnt->matches, &a.cnt->match_drop);
  else stage_flush(stage, a.pool, a.pool_cap, &a.cnt->pool_n, &a.cnt->overflow);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) fan += __shfl_xor_sync(0xffffffffu, fan, o);
  if (lane == 0 && fan) atomicAdd(&a.cnt->fanout, fan);
}

// rows_out[i * width + j] = j-th vertex of completed walk i (subgraph file rows, tds_batch_1.hpp:685-689)
__global__ void k_tds_materialize(const uint2* __restrict__ pool, const uint2* __restrict__ matches,
                                  uint64_t n, int width, const uint32_t* __restrict__ vid,
                                  uint32_t* __restrict__ rows_out) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    const uint2 mt = matches[i];
    uint32_t* r = rows_out + i * width;
    r[width - 1] = vid[mt.y];
    uint32_t t = mt.x;
    for (int j = width - 2; j >= 0; --j) {
      const uint2 tk = pool[t];
      r[j] = vid[tk.y];
      t = tk.x;
    }
  }
}

// ---------------------------------------------------------------------------
// post-processing of token_source_map (beta.cpp:964-1005, 1043-1062): a source
// whose walk never completed loses bit pattern_indices[0] in template_vertices
// (T_arr ONLY — vertex_state.template_vertices keeps it); with no
// bit left it is deactivated and leaves the vertex_state_map (S == 0).
// ---------------------------------------------------------------------------
// pool_cap / match_cap: a walk that ran out of token pool, key table or match list (match_cap = 0: none kept)
// leaves the state untouched — the host retries the constraint with larger buffers.
__global__ void __launch_bounds__(kBlock) k_nlcc_apply(uint16_t* __restrict__ S, const uint8_t* __restrict__ ok,
                                                        const uint32_t* __restrict__ src_list,
                                                        DevCounters* cnt, int par, unsigned long long pool_cap,
                                                        unsigned long long match_cap) {
  if (cnt->overflow || cnt->pool_n > pool_cap || (match_cap && (cnt->matches > match_cap || cnt->match_drop))) return;
  const uint32_t n = cnt->n_src;
  const uint32_t nr = (n + 31u) & ~31u;  // whole warps: publish_mask is warp collective
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nr; i += gridDim.x * blockDim.x) {
    bool changed = false;
    uint32_t s = 0, T2 = 0;
    if (i < n) {
      s = src_list[i];
      const uint32_t T = S[s];
      if (!ok[s] && T != 0) {
        T2 = T & ~(1u << c_nlc.I[0]);
        S[s] = (uint16_t)T2;
        cnt->deleted = 1u;
        changed = T2 != T;
      }
    }
    if (c_peer.G > 1) publish_mask(changed, s, T2, cnt, par);
  }
}

}  // namespace pm
