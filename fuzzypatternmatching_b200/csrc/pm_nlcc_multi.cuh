// pm_nlcc_multi.cuh — NLCC token walks across GPUs (1-D partition, see PeerTab in pm_common.cuh).
//
// The reference routes every token visitor through the mailbox to the rank that owns the target
// vertex, where pre_visit applies the acceptance tests and the work-aggregation set
// (token_passing_pattern_matching_nonunique_nem_1.hpp:98-303; visitor_queue.hpp:395-434).  Here a hop
// is one kernel per GPU: the sender applies every test that needs only replicated state (label stream,
// template bit of the replicated mask array S) and stores the surviving token STRAIGHT into its region
// of the owner's token inbox over NVLink; the owner applies the (vertex, source) aggregation when it
// picks the token up at the start of the next hop — the order the reference uses.  Region fill counts
// are local counters that travel in the per-hop StepMsg all-gather, so there are no remote atomics.
// Acknowledgements (token_source_map[s] = 1, nem_1.hpp:326-342) are single-byte stores into the
// owner's `ok` array; the edge flag of a successful cycle (nem_1.hpp:764-770) is a remote atomicOr.
//   TDS (tds_batch_1.hpp): a token record carries its visited history (the reference's
//   visited_vertices array, :964) — `n` words per record — because parent links cannot cross GPUs.
#pragma once

#include "pm_nlcc.cuh"

namespace pm {

struct TokSrc {  // the tokens that arrived for this rank in the previous hop: G regions of the inbox
  unsigned long long n[PM_MAX_RANKS];
  unsigned long long total;
};

// region_cap: records a region holds.  A sender that ran out of room kept counting (and raised `overflow`, which
// makes the host retry the constraint with larger inboxes) but stored nothing past the region: never read there.
__device__ __forceinline__ TokSrc tok_src(const NlcArgs& a, unsigned long long region_cap) {
  TokSrc t;
  t.total = 0;
#pragma unroll
  for (int r = 0; r < PM_MAX_RANKS; ++r) {
    unsigned long long n = r < c_peer.G ? a.all[r].out_n[c_peer.rank] : 0ull;
    if (n > region_cap) n = region_cap;
    t.n[r] = n;
    t.total += n;
  }
  return t;
}

// index within the concatenation of the regions -> element index in the inbox (in units of records)
__device__ __forceinline__ unsigned long long tok_locate(const TokSrc& ts, unsigned long long t, unsigned long long region_cap) {
  int r = 0;
#pragma unroll
  for (int q = 0; q < PM_MAX_RANKS - 1; ++q)
    if (r == q && t >= ts.n[q]) { t -= ts.n[q]; r = q + 1; }
  return (unsigned long long)r * region_cap + t;
}

__device__ __forceinline__ void ack_source(const NlcArgs& a, uint32_t s) {
  if (a.ok[s]) return;  // own source, or already acknowledged from this GPU
  a.ok[s] = 1;
  const uint32_t o = cid_owner(s);
  if ((int)o != c_peer.rank) c_peer.ok[o][s] = 1;
}

// flag E_s[parent] at the owner of s (nem_1.hpp:764-770)
__device__ __forceinline__ void mark_edge_m(uint32_t s, uint32_t parent) {
  const uint32_t o = cid_owner(s), li = s - c_peer.off[o];
  const uint64_t row = (uint64_t)c_peer.rowblk[o][li] * 8;
  const uint32_t d = c_peer.adeg[o][li];
  uint32_t* __restrict__ cw = c_peer.colw[o];
  uint32_t lo = 0, hi = d;
  while (lo < hi) {
    const uint32_t mid = (lo + hi) >> 1;
    const uint32_t x = cw[row + mid] & PM_IDMASK;
    if (x < parent) lo = mid + 1; else hi = mid;
  }
  if (lo < d && (cw[row + lo] & PM_IDMASK) == parent) atomicOr(&cw[row + lo], 0x80000000u);
}

// Routing of accepted tokens to the inbox regions of their owners.  Every warp stages its tokens per destination
// in shared memory and only when a destination's buffer runs full reserves inbox room — ONE atomicAdd on this rank's
// counter cnt->out_n[g] for up to kRouteCap tokens (a counter per destination is a single address: reserving per
// warp pass serialised several hundred thousand atomics per hop in the L2) — and stores the batch as whole
// 256-byte lines over NVLink.  All 32 lanes call; route_finish() at the end of the kernel drains the buffers.
// DEST 0: the records go to the owner of the compact id u; 1: to the owner of the SOURCE s (closing requests,
// k_close_check_m); 2: u names a SLOT (run_fuzzy path), owner = u / nlmax
constexpr int kRouteCap = 64;
struct RouteStage {
  uint2 buf[kBlock / 32][PM_MAX_RANKS][kRouteCap];
  uint32_t fill[kBlock / 32][PM_MAX_RANKS];
};

__device__ __forceinline__ void route_init(RouteStage& st) {
  const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane < PM_MAX_RANKS) st.fill[wid][lane] = 0u;
  __syncwarp();
}

__device__ __forceinline__ void route_flush(const NlcArgs& a, RouteStage& st, int g) {
  const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const uint32_t n = st.fill[wid][g];
  if (n == 0u) return;  // the same value in every lane
  unsigned long long basep = 0;
  if (lane == 0) basep = atomicAdd(&a.cnt->out_n[g], (unsigned long long)n);
  basep = __shfl_sync(0xffffffffu, basep, 0);
  uint2* __restrict__ dst = c_peer.tin[a.par][g] + (unsigned long long)c_peer.rank * c_peer.tcap;
  for (uint32_t i = lane; i < n; i += 32) {
    const unsigned long long pos = basep + i;
    if (pos < c_peer.tcap) dst[pos] = st.buf[wid][g][i];
    else a.cnt->overflow = 1u;
  }
  __syncwarp();
  if (lane == 0) st.fill[wid][g] = 0u;
  __syncwarp();
}

template <int DEST = 0>
__device__ __forceinline__ void route_tokens(const NlcArgs& a, RouteStage& st, const bool (&flag)[4], const uint32_t (&u)[4], uint32_t s) {
  const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const uint32_t lt = (1u << lane) - 1u;
  const uint32_t so = DEST == 1 ? cid_owner(s) : 0u;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (__ballot_sync(0xffffffffu, flag[k]) == 0u) continue;
    const uint32_t dest = flag[k] ? (DEST == 1 ? so : DEST == 2 ? u[k] / c_peer.nlmax : cid_owner(u[k])) : 0xFFFFFFFFu;
    for (int g = 0; g < c_peer.G; ++g) {
      const uint32_t m = __ballot_sync(0xffffffffu, dest == (uint32_t)g);
      if (m == 0u) continue;
      const uint32_t add = __popc(m);
      uint32_t fill = st.fill[wid][g];
      if (fill + add > (uint32_t)kRouteCap) {
        route_flush(a, st, g);
        fill = 0u;
      }
      if (dest == (uint32_t)g) st.buf[wid][g][fill + __popc(m & lt)] = make_uint2(u[k], s);
      __syncwarp();
      if (lane == 0) st.fill[wid][g] = fill + add;
      __syncwarp();
    }
  }
}

__device__ __forceinline__ void route_finish(const NlcArgs& a, RouteStage& st) {
  for (int g = 0; g < c_peer.G; ++g) route_flush(a, st, g);
}

// ---------------------------------------------------------------------------
// Closing-edge set of a cycle constraint.  The closing two hops (see k_nem1_close_cycle) need
// "is u a neighbour of the source s that passes the tests of the last interior hop C?" at the GPU that
// owns the token's vertex, but E_s lives at the owner of s, and peer loads (about 2 us each, never cached
// in L2) make a remote binary search per candidate far too slow.  Every rank therefore publishes the
// qualifying (s, u) pairs of its own sources to ALL ranks once per constraint — a broadcast through the
// token inboxes, a few MB — and every rank files them in its hash set under a tag bit.  The closing
// kernel then does one local probe per candidate.
// ---------------------------------------------------------------------------
#define PM_CE_TAG 0x8000000000000000ull  // vertex slots are < 2^31, so (vertex, source) keys never carry bit 63

__global__ void __launch_bounds__(kBlock) k_close_keys_m(NlcArgs a, const uint4* __restrict__ l0,
                                                          const uint4* __restrict__ l1, int cur, int hn) {
  const uint32_t c0 = a.cnt->fr_n[cur][0], c1 = a.cnt->fr_n[cur][1];
  const uint32_t total = c0 + c1;
  constexpr int GROUP = 8;
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t gl = lane % GROUP, gw = lane / GROUP;
  const uint32_t lt = (1u << lane) - 1u;
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t base = warp * 4; base < total; base += nwarps * 4) {
    const uint32_t i = base + gw;
    uint32_t s = 0, d = 0;
    uint64_t row = 0;
    if (i < total) {
      const uint4 e = i < c0 ? l0[i] : l1[i - c0];
      s = e.x;  // compact id
      const uint32_t li = s - a.base;
      const uint32_t T = e.y != PM_TOMB ? (uint32_t)a.S[s] : 0u;
      // a source of the constraint that can also receive the closing hop (nem_1.hpp:428-451, 557-581)
      if (T != 0 && hop_ok(T, a.cls[s], 0) && hop_ok(T, a.cls[s], hn + 1)) {
        d = a.adeg[li];
        row = (uint64_t)a.rowblk[li] * 8;
      }
    }
    const uint32_t passes = (d + GROUP * 4 - 1) / (GROUP * 4);
    const uint32_t maxp = __reduce_max_sync(0xffffffffu, passes);
    for (uint32_t p = 0; p < maxp; ++p) {
      const uint32_t j0 = p * GROUP * 4 + gl * 4;
      uint4 q = make_uint4(0, 0, 0, 0);
      if (j0 < d) {
        q = *reinterpret_cast<const uint4*>(a.colw + row + j0);
      }
      const uint32_t u[4] = {q.x & PM_IDMASK, q.y & PM_IDMASK, q.z & PM_IDMASK, q.w & PM_IDMASK};
      bool ok[4];
      uint32_t n = 0;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        bool may = j0 + k < d && u[k] != s;
        ok[k] = false;
        if (may) {
          const uint32_t su = a.S[u[k]];
          ok[k] = su != 0 && ((su >> c_nlc.I[hn]) & 1u);
        }
        n += ok[k];
      }
      // one reservation per warp and pass
      uint32_t incl = n;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= (uint32_t)o) incl += t;
      }
      const uint32_t tot = __shfl_sync(0xffffffffu, incl, 31);
      if (tot == 0) continue;
      unsigned long long pos = 0;
      if (lane == 31) pos = atomicAdd(&a.cnt->out_n[0], (unsigned long long)tot);
      pos = __shfl_sync(0xffffffffu, pos, 31) + incl - n;
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (ok[k]) {
          if (pos < c_peer.tcap) {
            const uint2 key = make_uint2(u[k], s);
            for (int g = 0; g < c_peer.G; ++g) c_peer.tin[a.par][g][(unsigned long long)c_peer.rank * c_peer.tcap + pos] = key;
          } else {
            a.cnt->overflow = 1u;
          }
          ++pos;
        }
    }
  }
  (void)lt;
}

// the same count goes to every rank
__global__ void k_close_keys_count_m(DevCounters* cnt) {
  const unsigned long long n = cnt->out_n[0];
  for (int g = 1; g < c_peer.G; ++g) cnt->out_n[g] = n;
}

__global__ void __launch_bounds__(kBlock) k_close_ingest_m(NlcArgs a) {
  const TokSrc src = tok_src(a, c_peer.tcap);
  const uint2* __restrict__ in = c_peer.tin[a.par ^ 1][c_peer.rank];
  for (unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; t < src.total;
       t += (unsigned long long)gridDim.x * blockDim.x) {
    const uint2 k = in[tok_locate(src, t, c_peer.tcap)];
    const unsigned long long key = PM_CE_TAG | ((unsigned long long)k.y << 32) | k.x;  // (source, neighbour)
    uint64_t h = mix64(key) & a.hset_mask;
    int probe = 0;
    for (; probe < 256; ++probe) {
      const unsigned long long prev = atomicCAS(&a.hset[h], PM_HSET_EMPTY, key);
      if (prev == PM_HSET_EMPTY || prev == key) break;
      h = (h + 1) & a.hset_mask;
    }
    if (probe == 256) a.cnt->overflow = 1u;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) a.cnt->ce_n = src.total;
}

__device__ __forceinline__ bool close_edge_known(const NlcArgs& a, uint32_t s, uint32_t u) {
  const unsigned long long key = PM_CE_TAG | ((unsigned long long)s << 32) | u;
  uint64_t h = mix64(key) & a.hset_mask;
  for (int probe = 0; probe < 256; ++probe) {
    const unsigned long long x = a.hset[h];
    if (x == key) return true;
    if (x == PM_HSET_EMPTY) return false;
    h = (h + 1) & a.hset_mask;
  }
  return false;
}

// ---------------------------------------------------------------------------
// nem_1 across GPUs: tokens of the previous hop (inbox `par ^ 1`) -> hop hn (inbox `par` of the owners)
//   MODE 0: interior hop, 1: final hop of a path or (generic) cycle constraint, 2: closing two hops of a cycle
//   first: the tokens are the sources themselves (no aggregation test)
// ---------------------------------------------------------------------------
// first: no aggregation test on arrival — the tokens are the sources themselves, or were accepted at a hop where
//   duplicates cannot occur (hop 1) or cannot matter (the level feeding the closing mode, up to hop 2)
template <int MODE>
__global__ void __launch_bounds__(kBlock) k_nem1_hop_m(NlcArgs a, int hn, int first) {
  __shared__ RouteStage st;
  route_init(st);
  const TokSrc src = tok_src(a, c_peer.tcap);
  const uint2* __restrict__ in = c_peer.tin[a.par ^ 1][c_peer.rank];
  constexpr int GROUP = 4;         // lanes per token: 16 slots of its row per pass (the pruned rows are short)
  constexpr int TPW = 32 / GROUP;  // tokens per warp and iteration
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t gl = lane % GROUP, gw = lane / GROUP;
  const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  unsigned long long fan = 0, accepted = 0;
  for (uint64_t base = warp * 32; base < src.total; base += nwarps * 32) {
    // every lane picks up ONE token: 32 independent chains token -> aggregation -> row header are in flight per warp
    // (a group of lanes per token would serialise them eight at a time); the rows are then walked TPW tokens at a time
    const uint64_t tl = base + lane;
    uint32_t v_l = 0, s_l = 0, d_l = 0, rb_l = 0;
    if (tl < src.total) {
      const uint2 tk = in[tok_locate(src, tl, c_peer.tcap)];
      // work aggregation at the receiver: one token per (vertex, source) (nem_1.hpp:131-139, 270-285)
      if (first || hset_insert(a, tk.x, tk.y)) {
        v_l = tk.x;
        s_l = tk.y;
        d_l = a.adeg[v_l - a.base];
        rb_l = a.rowblk[v_l - a.base];
        accepted++;
        if (MODE == 1 && !c_nlc.valid_cycle && a.ok[s_l]) d_l = 0;  // acknowledged path source: later tokens are moot
        if (MODE == 2 || MODE == 3) {
          const uint32_t ss = a.S[s_l];
          if (ss == 0 || !hop_ok(ss, a.cls[s_l], hn + 1)) d_l = 0;  // receiver tests of the closing hop at the source
        }
        fan += d_l;
      }
    }
    for (int sub = 0; sub < GROUP; ++sub) {
      const int from = sub * TPW + (int)gw;
      const uint32_t v = __shfl_sync(0xffffffffu, v_l, from), s = __shfl_sync(0xffffffffu, s_l, from);
      const uint32_t d = __shfl_sync(0xffffffffu, d_l, from);
      const uint64_t row = (uint64_t)__shfl_sync(0xffffffffu, rb_l, from) * 8;
      const uint32_t passes = (d + GROUP * 4 - 1) / (GROUP * 4);
      const uint32_t maxp = __reduce_max_sync(0xffffffffu, passes);
      for (uint32_t p = 0; p < maxp; ++p) {
        const uint32_t j0 = p * GROUP * 4 + gl * 4;
        uint4 q = make_uint4(0, 0, 0, 0);
        if (j0 < d) {
          q = *reinterpret_cast<const uint4*>(a.colw + row + j0);
        }
        const uint32_t u[4] = {q.x & PM_IDMASK, q.y & PM_IDMASK, q.z & PM_IDMASK, q.w & PM_IDMASK};
        bool pass[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          pass[k] = false;
          bool may = j0 + k < d;
          if (MODE == 2) may = may && u[k] != s;
          if (!may) continue;
          if (MODE == 2) {
            // u must be a qualifying neighbour of the source: one probe of the closing-edge set (which
            // already holds the label / template-bit tests of hop C, evaluated by the owner of s); the template
            // bit of u is tested first — a gather from the L2 resident mask replica that spares most probes
            const uint32_t su = a.S[u[k]];
            pass[k] = su != 0 && ((su >> c_nlc.I[hn]) & 1u) && close_edge_known(a, s, u[k]);
            if (pass[k]) {
              ack_source(a, s);
              a.cnt->found = 1u;
              mark_edge_m(s, u[k]);  // rare: only completed cycles get here
            }
            continue;
          }
          const uint32_t su = a.S[u[k]];
          pass[k] = su != 0 && ((su >> c_nlc.I[hn]) & 1u);
        }
        if (MODE == 3) {
          // closing two hops of a cycle, routed: every neighbour u of v that passes the tests of the last interior hop
          // becomes a closing request (u, s) for the owner of s, who looks u up in E_s (k_close_check_m)
          bool req[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) req[k] = pass[k] && u[k] != s;  // the source cannot relay (nem_1.hpp:174-177)
          route_tokens<1>(a, st, req, u, s);
        }
        if (MODE == 1) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const bool succ = pass[k] && (c_nlc.valid_cycle ? u[k] == s : u[k] != s);
            if (succ) {
              ack_source(a, s);
              a.cnt->found = 1u;
              if (c_nlc.valid_cycle) mark_edge_m(s, v);
            }
          }
        } else if (MODE == 0) {
          bool ins[4];
#pragma unroll
          for (int k = 0; k < 4; ++k) ins[k] = pass[k] && u[k] != s;  // the source cannot relay (nem_1.hpp:174-177)
          // one source per group: lanes of a group share s, lanes of different groups do not
          route_tokens(a, st, ins, u, s);
        }
      }
    }
  }
  if (MODE == 0 || MODE == 3) route_finish(a, st);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    fan += __shfl_xor_sync(0xffffffffu, fan, o);
    accepted += __shfl_xor_sync(0xffffffffu, accepted, o);
  }
  if (lane == 0 && fan) atomicAdd(&a.cnt->fanout, fan);
  if (lane == 0 && accepted) atomicAdd(&a.cnt->pool_n, accepted);
}

// Closing requests (u, s) that arrived for the sources this rank owns: the cycle closes iff u is a key of E_s
// (the edge maps are symmetric between live vertices, see k_nem1_close_cycle): a binary search in the short,
// local row of s.  Success acknowledges the source and flags E_s[u] (nem_1.hpp:749-770) — all local, no remote
// atomics, no key broadcast.
__global__ void __launch_bounds__(kBlock) k_close_check_m(NlcArgs a) {
  const TokSrc src = tok_src(a, c_peer.tcap);
  const uint2* __restrict__ in = c_peer.tin[a.par ^ 1][c_peer.rank];
  for (unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; t < src.total;
       t += (unsigned long long)gridDim.x * blockDim.x) {
    const uint2 rq = in[tok_locate(src, t, c_peer.tcap)];
    const uint32_t u = rq.x, s = rq.y, li = s - a.base;
    const uint64_t row = (uint64_t)a.rowblk[li] * 8;
    const uint32_t d = a.adeg[li];
    uint32_t b = 0, e = d;
    while (b < e) {  // rows stay ascending: compaction is stable
      const uint32_t mid = (b + e) >> 1;
      const uint32_t x = a.colw[row + mid] & PM_IDMASK;
      if (x < u) b = mid + 1; else e = mid;
    }
    if (b < d && (a.colw[row + b] & PM_IDMASK) == u) {
      a.ok[s] = 1;
      a.cnt->found = 1u;
      atomicOr(&a.colw[row + b], 0x80000000u);
    }
  }
}

// ---------------------------------------------------------------------------
// TDS across GPUs.  A record is `n` words: hist[0..h] = the walk so far (h = hn - 1).
// Completed walks (FINAL) are records of n words routed to the owner of the last vertex, which is
// where the reference writes the subgraph line (tds_batch_1.hpp:684-693).
// ---------------------------------------------------------------------------
template <bool FINAL>
__global__ void __launch_bounds__(kBlock) k_tds_hop_m(NlcArgs a, int hn) {
  const int n = c_nlc.n;
  const unsigned long long rcap = c_peer.tcap * 2ull / (unsigned long long)n;  // records per region
  const TokSrc src = tok_src(a, rcap);
  const uint32_t* __restrict__ in = reinterpret_cast<const uint32_t*>(c_peer.tin[a.par ^ 1][c_peer.rank]);
  constexpr int GROUP = 8;
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t gl = lane % GROUP, gw = lane / GROUP;
  const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  const int h = hn - 1;
  unsigned long long fan = 0, accepted = 0;
  for (uint64_t base = warp * 4; base < src.total; base += nwarps * 4) {
    const uint64_t t = base + gw;
    const bool has = t < src.total;
    uint32_t hist[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) hist[i] = 0xFFFFFFFFu;
    uint32_t v = 0, d = 0;
    if (has) {
      // region r starts at byte offset r * tcap * 8 whatever the record width
      TokSrc tmp = src;
      unsigned long long tt = t;
      int r = 0;
#pragma unroll
      for (int q = 0; q < PM_MAX_RANKS - 1; ++q)
        if (r == q && tt >= tmp.n[q]) { tt -= tmp.n[q]; r = q + 1; }
      const uint32_t* rec = in + (unsigned long long)r * c_peer.tcap * 2ull + tt * (unsigned long long)n;
#pragma unroll
      for (int i = 0; i < 16; ++i)
        if (i <= h) hist[i] = rec[i];
#pragma unroll
      for (int i = 0; i < 16; ++i)
        if (i == h) v = hist[i];
      d = a.adeg[v - a.base];
      if (gl == 0) accepted++;
    }
    const uint32_t s = hist[0];
    const uint64_t row = has ? (uint64_t)a.rowblk[v - a.base] * 8 : 0;
    const uint32_t passes = (d + GROUP * 4 - 1) / (GROUP * 4);
    const uint32_t maxp = __reduce_max_sync(0xffffffffu, passes);
    for (uint32_t p = 0; p < maxp; ++p) {
      const uint32_t j0 = p * GROUP * 4 + gl * 4;
      uint4 q = make_uint4(0, 0, 0, 0);
      if (j0 < d) {
        q = *reinterpret_cast<const uint4*>(a.colw + row + j0);
      }
      const uint32_t u[4] = {q.x & PM_IDMASK, q.y & PM_IDMASK, q.z & PM_IDMASK, q.w & PM_IDMASK};
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        bool acc = false;
        bool may = j0 + k < d;
        if (may) {
          const uint32_t su = a.S[u[k]];
          acc = su != 0 && ((su >> c_nlc.I[hn]) & 1u);
          if (acc) {
            if (FINAL) acc = c_nlc.valid_cycle ? (u[k] == s) : (u[k] != s && hist_rule(hist, hn, u[k]));
            else acc = hist_rule(hist, hn, u[k]);
          }
        }
        // records are rare compared with nem_1 tokens: one reservation per record
        if (acc) {
          const uint32_t g = cid_owner(u[k]);
          const unsigned long long pos = atomicAdd(&a.cnt->out_n[g], 1ull);
          if (pos < rcap) {
            uint32_t* out = reinterpret_cast<uint32_t*>(c_peer.tin[a.par][g]) +
                            (unsigned long long)c_peer.rank * c_peer.tcap * 2ull + pos * (unsigned long long)n;
#pragma unroll
            for (int i = 0; i < 16; ++i)
              if (i <= h) out[i] = hist[i];
            out[hn] = u[k];
          } else {
            a.cnt->overflow = 1u;
          }
          if (FINAL) {  // walk completed (tds_batch_1.hpp:664-694, 699-750)
            ack_source(a, s);
            a.cnt->found = 1u;
          }
        }
      }
    }
    if (has && gl == 0) fan += d;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    fan += __shfl_xor_sync(0xffffffffu, fan, o);
    accepted += __shfl_xor_sync(0xffffffffu, accepted, o);
  }
  if (lane == 0 && fan) atomicAdd(&a.cnt->fanout, fan);
  if (lane == 0 && accepted) atomicAdd(&a.cnt->pool_n, accepted);
}

// gathers the completed walks that arrived for this rank into one dense row array
__global__ void k_tds_collect_m(NlcArgs a, int n, uint32_t* __restrict__ rows_out) {
  const TokSrc src = tok_src(a, c_peer.tcap * 2ull / (unsigned long long)n);
  const uint32_t* __restrict__ in = reinterpret_cast<const uint32_t*>(c_peer.tin[a.par ^ 1][c_peer.rank]);
  for (unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; t < src.total;
       t += (unsigned long long)gridDim.x * blockDim.x) {
    unsigned long long tt = t;
    int r = 0;
#pragma unroll
    for (int q = 0; q < PM_MAX_RANKS - 1; ++q)
      if (r == q && tt >= src.n[q]) { tt -= src.n[q]; r = q + 1; }
    const uint32_t* rec = in + (unsigned long long)r * c_peer.tcap * 2ull + tt * (unsigned long long)n;
    for (int i = 0; i < n; ++i) rows_out[t * n + i] = a.vid[rec[i]];
  }
}

}  // namespace pm
