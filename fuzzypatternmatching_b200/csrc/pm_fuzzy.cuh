// pm_fuzzy.cuh — the run_fuzzy_pattern_matching path on the GPU.
//
// Replaces
//   LCC   label_propagation_pattern_matching_bsp.hpp:66-310 (lppm_visitor), :317-520 (per-message
//         state), :527-593 (post step), :598-699 (superstep loop)
//   NLCC  token_passing_pattern_matching.hpp:98-335 (tppm_visitor)
//   loop  src/run_pattern_matching.cpp:340-722 (the compiling twin of the stale
//         src/run_fuzzy_pattern_matching.cpp:287-557)
// Paths relative to /root/reference (headers under include/havoqgt/).
//
// Formulation.  In this path a vertex stands for ONE template vertex, q0(v) = the first template
// vertex carrying its label (pre_visit only ever tests that one, bsp.hpp:113-141, and it is the one
// recorded as vertex_pattern_index, :360-370); a sender announces every template index of its label
// to ALL graph neighbours every superstep (no edge elimination, :207-225), and v stays in the
// vertex_state_map iff within one superstep it hears every template neighbour of q0(v)
// (:438-466, :527-593).  Since template indices of one label always travel together, "heard index p"
// is "some sending neighbour has label(p)": v survives iff
//     req[q0(v)]  is a subset of  { label(u) : u in adj(v), u sends }      (req: PatConst)
// which is a pull over the pristine adjacency with the neighbour labels streamed next to the ids;
// the mask gather S[u] != 0 ("u sends") is only needed for neighbours whose label is required.
// In the very first superstep every label-matching vertex sends, so the test reads the neighbour
// label signature sig[v] and walks no row at all.
// State: S[v] = 1 << q0(v) while v is in the map, else 0.
#pragma once

#include "pm_lcc.cuh"
#include "pm_nlcc.cuh"
#include "pm_nlcc_multi.cuh"

namespace pm {

struct FzArgs {
  const uint32_t* rowblk;
  const uint32_t* deg;
  const uint32_t* col0;
  const uint8_t* lab0;
  const uint8_t* lab8;
  uint16_t* S;
  DevCounters* cnt;
  RowStat* row;
  uint32_t idmask;  // id bits of a col0 slot (packed labels ride above them)
  uint32_t base;    // first slot of this rank: S / lab8 are indexed by slot, rowblk / deg / sig by slot - base
  int par;          // several ranks: delta inbox of the current step
};

__device__ __forceinline__ int fz_q0(uint32_t lm) { return __ffs(lm) - 1; }  // first template vertex of the label

// per-pattern initialisation + the first superstep of the first LCC call
__global__ void __launch_bounds__(kBlock) k_fz_init(FzArgs a, const unsigned long long* __restrict__ sig, uint64_t V,
                                                     uint4* fr, int buf) {
  __shared__ uint8_t s_cl[64];
  __shared__ uint16_t s_lm[17];
  if (threadIdx.x < 64) s_cl[threadIdx.x] = c_pat.cls_of_label[threadIdx.x];
  if (threadIdx.x < 17) s_lm[threadIdx.x] = c_pat.LMc[threadIdx.x];
  __syncthreads();
  constexpr int IT = 4;
  const uint64_t tile = (uint64_t)blockDim.x * IT;
  unsigned long long nv = 0;
  bool removed = false;
  for (uint64_t base = (uint64_t)blockIdx.x * tile; base < V; base += (uint64_t)gridDim.x * tile) {
    bool alive[IT];
    int bin[IT];
    uint4 val[IT];
#pragma unroll
    for (int k = 0; k < IT; ++k) {
      const uint64_t v = base + (uint64_t)k * blockDim.x + threadIdx.x;
      alive[k] = false;
      bin[k] = 0;
      val[k] = make_uint4(0, 0, 0, 0);
      if (v < V) {
        const uint32_t lm = s_lm[s_cl[a.lab8[a.base + v] & 63]];
        uint32_t s = 0;
        if (lm) {
          const int q0 = fz_q0(lm);
          const unsigned long long rq = c_pat.req[q0], sg = sig[v];
          alive[k] = rq != 0ull && (sg & rq) == rq;                 // heard every template neighbour
          removed = removed || ((sg & rq) != 0ull && !alive[k]);   // entered the map and left it (bsp.hpp:566-579)
          if (alive[k]) s = 1u << q0;
        }
        a.S[a.base + v] = (uint16_t)s;
        if (alive[k]) { nv++; val[k] = make_uint4(a.base + (uint32_t)v, a.rowblk[v], 0u, 1u); }
      }
    }
    block_append2<IT>(alive, bin, val, fr, fr, &a.cnt->fr_n[buf][0]);
  }
  if (removed) a.cnt->nf = 1u;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) nv += __shfl_xor_sync(0xffffffffu, nv, o);
  if ((threadIdx.x & 31) == 0 && nv) atomicAdd(&a.row->nv, nv);
}

// one superstep: every vertex in the map pulls the labels of its sending neighbours
__global__ void __launch_bounds__(kBlock) k_fz_scan(FzArgs a, uint4* __restrict__ list, const uint32_t* __restrict__ n_ptr) {
  constexpr int GROUP = 8;
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t gl = lane % GROUP, gw = lane / GROUP;
  const uint32_t n = *n_ptr;
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
  unsigned long long scanned = 0, verts = 0;
  for (uint32_t base = warp * 4; base < n; base += nwarps * 4) {
    const uint32_t idx = base + gw;
    const bool has = idx < n;
    uint4 e = make_uint4(0, 0, 0, 0);
    uint32_t Sv = 0, d = 0;
    if (has) {
      e = list[idx];
      Sv = a.S[e.x];
      if (Sv) d = a.deg[e.x - a.base];  // erased by token passing since the last commit: nothing to do
    }
    const unsigned long long rq = Sv ? c_pat.req[fz_q0(Sv)] : 0ull;
    const uint64_t row = (uint64_t)e.y * 8;
    const uint32_t passes = (d + GROUP * 4 - 1) / (GROUP * 4);
    const uint32_t maxp = __reduce_max_sync(0xffffffffu, passes);
    unsigned long long heard = 0;
    for (uint32_t p = 0; p < maxp; ++p) {
      const uint32_t j0 = p * GROUP * 4 + gl * 4;
      if (j0 < d) {
        const uint32_t l4 = *reinterpret_cast<const uint32_t*>(a.lab0 + row + j0);
        // only neighbours whose label is still missing can matter: fetch their ids and masks
        bool need = false;
#pragma unroll
        for (int k = 0; k < 4; ++k) need = need || (j0 + k < d && ((rq & ~heard) >> ((l4 >> (8 * k)) & 63u)) & 1ull);
        if (need) {
          const uint4 q = *reinterpret_cast<const uint4*>(a.col0 + row + j0);
          const uint32_t u[4] = {q.x & a.idmask, q.y & a.idmask, q.z & a.idmask, q.w & a.idmask};
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint32_t lab = (l4 >> (8 * k)) & 63u;
            if (j0 + k < d && (((rq & ~heard) >> lab) & 1ull) && a.S[u[k]] != 0) heard |= 1ull << lab;
          }
        }
      }
    }
    heard |= __shfl_xor_sync(0xffffffffu, heard, 1);
    heard |= __shfl_xor_sync(0xffffffffu, heard, 2);
    heard |= __shfl_xor_sync(0xffffffffu, heard, 4);
    if (has && gl == 0) {
      e.w = (Sv != 0 && rq != 0ull && (heard & rq) == rq) ? 1u : 0u;
      list[idx] = e;
      scanned += d;
      verts += Sv != 0;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    scanned += __shfl_xor_sync(0xffffffffu, scanned, o);
    verts += __shfl_xor_sync(0xffffffffu, verts, o);
  }
  if (lane == 0 && verts) {
    atomicAdd(&a.row->scanned[0], scanned);
    atomicAdd(&a.row->verts[0], verts);
  }
}

// post step (bsp.hpp:527-593): vertices that did not hear everything leave the map
__global__ void __launch_bounds__(kBlock) k_fz_commit(FzArgs a, const uint4* __restrict__ l0, uint4* n0, int cur, int nxt) {
  const uint32_t total = a.cnt->fr_n[cur][0];
  unsigned long long nv = 0;
  constexpr int IT = 4;
  const uint32_t tile = blockDim.x * IT;
  for (uint32_t base = blockIdx.x * tile; base < total; base += gridDim.x * tile) {
    bool alive[IT];
    int bin[IT];
    uint4 val[IT];
#pragma unroll
    for (int k = 0; k < IT; ++k) {
      const uint32_t i = base + k * blockDim.x + threadIdx.x;
      alive[k] = false;
      bin[k] = 0;
      val[k] = make_uint4(0, 0, 0, 0);
      bool changed = false;
      uint32_t cslot = 0;
      if (i < total) {
        const uint4 e = l0[i];
        const bool was = a.S[e.x] != 0;
        alive[k] = was && e.w != 0;
        if (was && !alive[k]) { a.S[e.x] = 0; a.cnt->nf = 1u; changed = true; cslot = e.x; }
        if (alive[k]) nv++;
        val[k] = e;
      }
      if (c_peer.G > 1) publish_mask(changed, cslot, 0u, a.cnt, a.par);  // peers drop the vertex from their replicas
    }
    block_append2<IT>(alive, bin, val, n0, n0, &a.cnt->fr_n[nxt][0]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) nv += __shfl_xor_sync(0xffffffffu, nv, o);
  if ((threadIdx.x & 31) == 0 && nv) atomicAdd(&a.row->nv, nv);
}

__global__ void __launch_bounds__(kBlock) k_fz_count(FzArgs a, const uint4* __restrict__ l0, int cur) {
  const uint32_t total = a.cnt->fr_n[cur][0];
  unsigned long long nv = 0;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x)
    if (a.S[l0[i].x]) nv++;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) nv += __shfl_xor_sync(0xffffffffu, nv, o);
  if ((threadIdx.x & 31) == 0 && nv) atomicAdd(&a.row->nv, nv);
}

// ---- token passing over the unpruned adjacency (token_passing_pattern_matching.hpp) -------------------
struct FzTok {
  uint8_t lab[18];  // label of hop h (P[h])
  uint8_t I[18];    // template vertex of hop h
  int C;            // pattern_cycle_length
  int valid_cycle;
};
__constant__ FzTok c_fz;

// sources: vertices in the map whose vertex_pattern_index is I[0] (tp.hpp:208-228)
__global__ void __launch_bounds__(kBlock) k_fz_sources(FzArgs a, const uint4* __restrict__ l0, int cur, uint8_t* ok,
                                                        uint32_t* src_list, uint2* pool, unsigned long long pool_cap) {
  const uint32_t total = a.cnt->fr_n[cur][0];
  const uint32_t lane = threadIdx.x & 31;
  uint32_t i0 = (blockIdx.x * blockDim.x + threadIdx.x) & ~31u;
  for (; i0 < total; i0 += gridDim.x * blockDim.x) {
    const uint32_t i = i0 + lane;
    bool is_src = false;
    uint32_t v = 0;
    if (i < total) {
      v = l0[i].x;
      const uint32_t s = a.S[v];
      is_src = s != 0 && fz_q0(s) == (int)c_fz.I[0] && a.lab8[v] == c_fz.lab[0];
    }
    const uint32_t m = __ballot_sync(0xffffffffu, is_src);
    if (m) {
      uint32_t base = 0;
      if (lane == 0) base = atomicAdd(&a.cnt->n_src, (uint32_t)__popc(m));
      base = __shfl_sync(0xffffffffu, base, 0);
      if (is_src) {
        const uint32_t pos = base + __popc(m & lanemask_lt());
        src_list[pos] = v;
        ok[v] = 0;
        if (c_peer.G > 1) {  // level 0 = my own region of my token inbox
          if (pos < c_peer.tcap) c_peer.tin[a.par][c_peer.rank][(unsigned long long)c_peer.rank * c_peer.tcap + pos] = make_uint2(v, v);
          else a.cnt->overflow = 1u;
        } else if (pos < pool_cap) pool[pos] = make_uint2(v, v);
        else a.cnt->overflow = 1u;
      }
    }
  }
}

// tokens accepted at hop hn-1 (level hlevel) -> hop hn; interior hops only (tp.hpp:98-170, 289-295)
__global__ void __launch_bounds__(kBlock) k_fz_expand(FzArgs a, NlcArgs t, int hlevel, int hn) {
  __shared__ uint2 s_stage[(kBlock / 32) * PM_STAGE_CAP];
  WarpStage stage{s_stage + (threadIdx.x >> 5) * PM_STAGE_CAP, 0u};
  const uint64_t lo = t.cnt->lvl[hlevel], hi = t.cnt->lvl[hlevel + 1];
  constexpr int GROUP = 8;
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t gl = lane % GROUP, gw = lane / GROUP;
  const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  const uint32_t want_lab = c_fz.lab[hn];
  const uint32_t want_s = 1u << c_fz.I[hn];
  unsigned long long fan = 0;
  for (uint64_t base = lo + warp * 4; base < hi; base += nwarps * 4) {
    const uint64_t ti = base + gw;
    const bool has = ti < hi;
    uint32_t v = 0, s = 0, d = 0;
    if (has) {
      const uint2 tk = t.pool[ti];
      v = tk.x;
      s = tk.y;
      d = a.deg[v];
    }
    const uint64_t row = has ? (uint64_t)a.rowblk[v] * 8 : 0;
    const uint32_t passes = (d + GROUP * 4 - 1) / (GROUP * 4);
    const uint32_t maxp = __reduce_max_sync(0xffffffffu, passes);
    for (uint32_t p = 0; p < maxp; ++p) {
      const uint32_t j0 = p * GROUP * 4 + gl * 4;
      uint4 q = make_uint4(0, 0, 0, 0);
      uint32_t l4 = 0;
      if (j0 < d) {
        q = *reinterpret_cast<const uint4*>(a.col0 + row + j0);
        l4 = *reinterpret_cast<const uint32_t*>(a.lab0 + row + j0);
      }
      const uint32_t u[4] = {q.x & a.idmask, q.y & a.idmask, q.z & a.idmask, q.w & a.idmask};
      bool ins[4];
      uint2 tok[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        // in the map with the expected vertex_pattern_index (S[u] == 1 << I[hn]) and label (tp.hpp:127-131);
        // one token per (vertex, source) at interior hops (:104-109, 137-149) — the source itself may relay
        ins[k] = j0 + k < d && ((l4 >> (8 * k)) & 0xffu) == want_lab && a.S[u[k]] == want_s;
        if (ins[k]) ins[k] = hset_insert(t, u[k], s);
        tok[k] = make_uint2(u[k], s);
      }
      stage_push(stage, ins, tok, t.pool, t.pool_cap, &t.cnt->pool_n, &t.cnt->overflow);
    }
    if (has && gl == 0) fan += d;
  }
  stage_flush(stage, t.pool, t.pool_cap, &t.cnt->pool_n, &t.cnt->overflow);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) fan += __shfl_xor_sync(0xffffffffu, fan, o);
  if (lane == 0 && fan) atomicAdd(&t.cnt->fanout, fan);
}

// final hop (max_itr_count == itr_count, tp.hpp:236-262): only the copy arriving at the source itself
// can complete a cycle, so the (ascending) row of v is searched for s
__global__ void __launch_bounds__(kBlock) k_fz_final(FzArgs a, NlcArgs t, int hlevel, int hn) {
  const unsigned long long lo = t.cnt->lvl[hlevel], hi = t.cnt->lvl[hlevel + 1];
  for (unsigned long long ti = lo + (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; ti < hi;
       ti += (unsigned long long)gridDim.x * blockDim.x) {
    const uint2 tk = t.pool[ti];
    const uint32_t v = tk.x, s = tk.y;
    if (a.S[s] != (1u << c_fz.I[hn]) || a.lab8[s] != c_fz.lab[hn]) continue;
    const uint64_t row = (uint64_t)a.rowblk[v] * 8;
    uint32_t b = 0, e = a.deg[v];
    const uint32_t d = e;
    while (b < e) {
      const uint32_t mid = (b + e) >> 1;
      const uint32_t x = a.col0[row + mid] & a.idmask;
      if (x < s) b = mid + 1; else e = mid;
    }
    if (b < d && (a.col0[row + b] & a.idmask) == s) {
      t.ok[s] = 1;
      t.cnt->found = 1u;
    }
  }
}

// TP_ORIG post-processing (run_pattern_matching.cpp:583-629): failed sources leave the map
__global__ void __launch_bounds__(kBlock) k_fz_apply(uint16_t* __restrict__ S, const uint8_t* __restrict__ ok,
                                                      const uint32_t* __restrict__ src_list, DevCounters* cnt, int par) {
  const uint32_t n = cnt->n_src;
  const uint32_t nr = (n + 31u) & ~31u;  // whole warps: publish_mask is warp collective
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nr; i += gridDim.x * blockDim.x) {
    bool changed = false;
    uint32_t s = 0;
    if (i < n) {
      s = src_list[i];
      if (!ok[s] && S[s] != 0) { S[s] = 0; cnt->deleted = 1u; changed = true; }
    }
    if (c_peer.G > 1) publish_mask(changed, s, 0u, cnt, par);
  }
}

// ---- several ranks: tokens travel through the owners' inboxes (see pm_nlcc_multi.cuh); vertices are SLOTS here ----
// tokens of the previous hop (inbox par ^ 1) -> hop hn (inbox par of the owners); first: the tokens are the sources
__global__ void __launch_bounds__(kBlock) k_fz_expand_m(FzArgs a, NlcArgs t, int hn, int first) {
  __shared__ RouteStage st;
  route_init(st);
  const TokSrc src = tok_src(t, c_peer.tcap);
  const uint2* __restrict__ in = c_peer.tin[t.par ^ 1][c_peer.rank];
  constexpr int GROUP = 8;
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t gl = lane % GROUP, gw = lane / GROUP;
  const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  const uint32_t want_lab = c_fz.lab[hn];
  const uint32_t want_s = 1u << c_fz.I[hn];
  unsigned long long fan = 0, accepted = 0;
  for (uint64_t base = warp * 4; base < src.total; base += nwarps * 4) {
    const uint64_t ti = base + gw;
    bool has = ti < src.total;
    uint32_t v = 0, s = 0, d = 0, fresh = 1;
    if (has) {
      const uint2 tk = in[tok_locate(src, ti, c_peer.tcap)];
      v = tk.x;
      s = tk.y;
      // one token per (vertex, source) at interior hops (tp.hpp:104-109, 137-149), applied where the token arrives
      if (!first && gl == 0) fresh = hset_insert(t, v, s) ? 1u : 0u;
    }
    fresh = __shfl_sync(0xffffffffu, fresh, gw * GROUP);
    if (!fresh) has = false;
    if (has) {
      d = a.deg[v - a.base];
      if (gl == 0) accepted++;
    }
    const uint64_t row = has ? (uint64_t)a.rowblk[v - a.base] * 8 : 0;
    const uint32_t passes = (d + GROUP * 4 - 1) / (GROUP * 4);
    const uint32_t maxp = __reduce_max_sync(0xffffffffu, passes);
    for (uint32_t p = 0; p < maxp; ++p) {
      const uint32_t j0 = p * GROUP * 4 + gl * 4;
      uint4 q = make_uint4(0, 0, 0, 0);
      uint32_t l4 = 0;
      if (j0 < d) {
        q = *reinterpret_cast<const uint4*>(a.col0 + row + j0);
        l4 = *reinterpret_cast<const uint32_t*>(a.lab0 + row + j0);
      }
      const uint32_t u[4] = {q.x & a.idmask, q.y & a.idmask, q.z & a.idmask, q.w & a.idmask};
      bool ins[4];
#pragma unroll
      for (int k = 0; k < 4; ++k)
        ins[k] = j0 + k < d && ((l4 >> (8 * k)) & 0xffu) == want_lab && a.S[u[k]] == want_s;
      route_tokens<2>(t, st, ins, u, s);
    }
    if (has && gl == 0) fan += d;
  }
  route_finish(t, st);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    fan += __shfl_xor_sync(0xffffffffu, fan, o);
    accepted += __shfl_xor_sync(0xffffffffu, accepted, o);
  }
  if (lane == 0 && fan) atomicAdd(&t.cnt->fanout, fan);
  if (lane == 0 && accepted) atomicAdd(&t.cnt->pool_n, accepted);
}

// final hop: the tokens that arrived at hop C complete a cycle iff their vertex is adjacent to the source
// (tp.hpp:236-262); the acknowledgement is one byte stored at the owner of the source
__global__ void __launch_bounds__(kBlock) k_fz_final_m(FzArgs a, NlcArgs t, int hn) {
  const TokSrc src = tok_src(t, c_peer.tcap);
  const uint2* __restrict__ in = c_peer.tin[t.par ^ 1][c_peer.rank];
  for (unsigned long long ti = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; ti < src.total;
       ti += (unsigned long long)gridDim.x * blockDim.x) {
    const uint2 tk = in[tok_locate(src, ti, c_peer.tcap)];
    const uint32_t v = tk.x, s = tk.y;
    if (t.ok[s]) continue;  // own source already done, or acknowledged from this GPU before
    if (a.S[s] != (1u << c_fz.I[hn]) || a.lab8[s] != c_fz.lab[hn]) continue;
    const uint64_t row = (uint64_t)a.rowblk[v - a.base] * 8;
    uint32_t b = 0, e = a.deg[v - a.base];
    const uint32_t d = e;
    while (b < e) {
      const uint32_t mid = (b + e) >> 1;
      const uint32_t x = a.col0[row + mid] & a.idmask;
      if (x < s) b = mid + 1; else e = mid;
    }
    if (b < d && (a.col0[row + b] & a.idmask) == s) {
      t.ok[s] = 1;
      const uint32_t o = s / c_peer.nlmax;
      if ((int)o != c_peer.rank) c_peer.ok[o][s] = 1;
      t.cnt->found = 1u;
    }
  }
}

}  // namespace pm
