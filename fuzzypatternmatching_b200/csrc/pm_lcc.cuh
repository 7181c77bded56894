// pm_lcc.cuh — local constraint checking (LCC) supersteps on the GPU.
//
// Replaces, for one superstep, the reference's
//   sender    lppm_visitor::visit      (label_propagation_pattern_matching_nonunique_ee.hpp:467-636)
//   receiver  lppm_visitor::pre_visit  (:148-459) + verify_and_update_vertex_state (:646-816)
//   post step verify_and_update_vertex_state free function (:827-1027)
// Paths are relative to /root/reference/include/havoqgt/.
//
// Formulation.  The reference pushes one message per directed active edge and
// updates hash maps at the receiver.  Every value a receiver reads during the
// message phase (T_arr of both ends) is written only in the post step, so the
// phase is a Jacobi sweep and can be evaluated by PULLING: vertex v walks its own
// active adjacency E_v, gathers S[u] (= T_arr(u), 0 when u does not send) and
//   valid(v,u)  = S[u] & NB(S[v]) != 0          NB(T) = OR_{a in T} N(a)   (:673-722)
//   heard(v)    = OR { S[u] : valid(v,u) }                                 (:775)
//   E_v'        = { u in E_v : valid(v,u) or flag(v,u) set outside LCC }   (:791-813, :954-962)
//   T_state(v) &= { p : N(p) != 0 and N(p) subset of heard(v) }            (:901-939)
// E_v is kept as the first adeg[v] slots of v's row in `colw`; a superstep
// compacts the row in place (the GPU analogue of erasing map entries).  The
// template validity test is symmetric, so E stays symmetric between live
// vertices and pulling over E_v sees exactly the messages the reference
// delivers (the one exception, SURVEY A.6 #11 combined with #4, is detected by
// the oracle's hazard counters).  k_lcc_commit then publishes S and builds the
// next frontier, so S is single-buffered yet the sweep stays Jacobi.
#pragma once

#include "pm_common.cuh"

namespace pm {

__constant__ PatConst c_pat;

__device__ __forceinline__ uint32_t nb_of(uint32_t T) {
  uint32_t r = 0;
#pragma unroll
  for (int p = 0; p < 16; ++p)
    if ((T >> p) & 1u) r |= c_pat.N[p];
  return r;
}

// bits p of T whose template neighbourhood is non-empty and entirely heard
__device__ __forceinline__ uint32_t cover_of(uint32_t T, uint32_t heard) {
  uint32_t r = 0;
#pragma unroll
  for (int p = 0; p < 16; ++p) {
    uint32_t need = c_pat.N[p];
    if (((T >> p) & 1u) && need != 0u && (need & ~heard) == 0u) r |= 1u << p;
  }
  return r;
}

__device__ __forceinline__ uint32_t lanemask_lt() {
  uint32_t m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}

struct LccArgs {
  const uint32_t* rowblk;
  const uint32_t* deg;
  const uint32_t* col0;
  uint32_t* colw;
  uint16_t* S;
  uint16_t* Tst;
  uint32_t* adeg;
  const uint8_t* cls;
  DevCounters* cnt;
  RowStat* row;   // accumulator of this superstep
  int bin;        // degree bin this launch serves (row statistics)
};

// ---------------------------------------------------------------------------
// per-pattern initialisation (beta.cpp:484-492 + the label test every vertex
// performs in the first superstep, ee.hpp:371-380 / :523-546):
//   cls[v]  = class of label[v];  S[v] = labelmask(label[v])  (0: v goes inactive)
//   candidates are appended to the frontier bin of their degree
// ---------------------------------------------------------------------------
__global__ void k_init_state(const uint64_t* __restrict__ label, const uint32_t* __restrict__ deg,
                             uint64_t V, uint8_t* __restrict__ cls, uint16_t* __restrict__ S,
                             uint16_t* __restrict__ Tst, uint32_t* __restrict__ adeg,
                             uint32_t* fr_small, uint32_t* fr_mid, uint32_t* fr_big, DevCounters* cnt,
                             int buf) {
  const uint32_t lane = threadIdx.x & 31;
  uint64_t v0 = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) & ~31ull;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (; v0 < V; v0 += stride) {
    uint64_t v = v0 + lane;
    uint32_t c = PM_NOCLASS, d = 0;
    if (v < V) {
      uint64_t lab = label[v];
      d = deg[v];
#pragma unroll
      for (int k = 0; k < 16; ++k)
        if (k < c_pat.ncls && c_pat.clabel[k] == lab) c = k;
      cls[v] = (uint8_t)c;
      uint16_t lm = d ? c_pat.LMc[c] : (uint16_t)0;
      S[v] = lm;
      Tst[v] = lm;
      adeg[v] = 0;
      if (lm == 0) c = PM_NOCLASS;
    }
    const bool cand = (c != PM_NOCLASS);
    const int bin = d <= PM_SMALL_MAX ? 0 : (d <= PM_MID_MAX ? 1 : 2);
#pragma unroll
    for (int b = 0; b < 3; ++b) {
      uint32_t m = __ballot_sync(0xffffffffu, cand && bin == b);
      if (m) {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(&cnt->fr_n[buf][b], (uint32_t)__popc(m));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (cand && bin == b) {
          uint32_t* dst = b == 0 ? fr_small : (b == 1 ? fr_mid : fr_big);
          dst[base + __popc(m & lanemask_lt())] = (uint32_t)v;
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------
// scan: GROUP lanes walk the active adjacency of one vertex with uint4 loads
// (4 slots per lane and pass), gather the neighbour masks, OR the heard masks and
// compact the surviving neighbours to the front of the row.
//   FIRST = first superstep of the first iteration: walk the pristine adjacency
//   col0 (all deg[v] slots), neighbour mask = labelmask via the class array
//   (ee.hpp:519-561 sender, :368-404 receiver); otherwise walk keys(E_v) in colw.
// ---------------------------------------------------------------------------
template <int GROUP, bool FIRST>
__global__ void __launch_bounds__(kBlock) k_lcc_scan(LccArgs a, const uint32_t* __restrict__ list,
                                                      const uint32_t* __restrict__ n_ptr) {
  __shared__ uint16_t s_lm[17];
  if (threadIdx.x < 17) s_lm[threadIdx.x] = c_pat.LMc[threadIdx.x];
  __syncthreads();
  constexpr int GPW = 32 / GROUP;  // groups per warp
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t gl = lane % GROUP;
  const uint32_t gw = lane / GROUP;
  const uint32_t gmask = GROUP == 32 ? 0xffffffffu : (((1u << GROUP) - 1u) << (gw * GROUP));
  const uint32_t lt = lanemask_lt() & gmask;
  const uint32_t n = *n_ptr;
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
  unsigned long long scanned = 0, verts = 0;
  for (uint32_t base = warp * GPW; base < n; base += nwarps * GPW) {
    const uint32_t idx = base + gw;
    const bool has = idx < n;
    uint32_t v = 0, d = 0, Tv = 0;
    if (has) {
      v = list[idx];
      Tv = a.S[v];
      d = FIRST ? a.deg[v] : a.adeg[v];
      if (Tv == 0) d = 0;  // deactivated by NLCC since the last commit (beta.cpp:990-992)
    }
    const uint32_t NBv = nb_of(Tv);
    const uint64_t row = has ? (uint64_t)a.rowblk[v] * 8 : 0;
    const uint32_t* __restrict__ src = FIRST ? a.col0 : a.colw;
    const uint32_t passes = (d + GROUP * 4 - 1) / (GROUP * 4);
    const uint32_t maxp = GROUP == 32 ? passes : __reduce_max_sync(0xffffffffu, passes);
    uint32_t heard = 0, out = 0;
    for (uint32_t p = 0; p < maxp; ++p) {
      const uint32_t j0 = p * GROUP * 4 + gl * 4;
      uint4 q = make_uint4(PM_SENTINEL, PM_SENTINEL, PM_SENTINEL, PM_SENTINEL);
      if (j0 < d) q = *reinterpret_cast<const uint4*>(src + row + j0);
      uint32_t u[4] = {q.x, q.y, q.z, q.w};
      uint32_t m[4];
      bool keep[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const bool act = j0 + k < d;
        const uint32_t uu = u[k] & PM_IDMASK;
        m[k] = 0;
        if (act) m[k] = FIRST ? (uint32_t)s_lm[a.cls[uu]] : (uint32_t)a.S[uu];
      }
      uint32_t cnt_lane = 0;
      uint32_t below = 0;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const bool act = j0 + k < d;
        const bool valid = (m[k] & NBv) != 0u;
        const bool pre = !FIRST && act && (u[k] >> 31);
        keep[k] = valid || pre;
        if (valid) heard |= m[k];
        const uint32_t b = __ballot_sync(0xffffffffu, keep[k]);
        below += __popc(b & lt);
        cnt_lane += __popc(b & gmask);
      }
      // all loads of this pass are complete (ballots synchronise the group) and
      // every write lands at or before a slot read in this or an earlier pass
      uint32_t pos = out + below;
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (keep[k]) a.colw[row + pos++] = u[k] & PM_IDMASK;
      out += cnt_lane;
    }
#pragma unroll
    for (int o = GROUP / 2; o > 0; o >>= 1) heard |= __shfl_xor_sync(0xffffffffu, heard, o);
    if (has && gl == 0) {
      const uint32_t T0 = FIRST ? Tv : (uint32_t)a.Tst[v];
      const uint32_t ts = Tv ? cover_of(T0, heard) : 0u;
      a.Tst[v] = (uint16_t)ts;
      a.adeg[v] = out;
      // a vertex leaves the vertex_state_map (ee.hpp:941-946, :968-970).  In the first
      // superstep only vertices that heard a valid neighbour ever entered the map (:841-852).
      if (ts == 0 && (FIRST ? heard != 0u : Tv != 0u)) a.cnt->nf = 1u;
      scanned += d;
      verts += Tv != 0u;
    }
  }
  // one atomic per warp
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    scanned += __shfl_xor_sync(0xffffffffu, scanned, o);
    verts += __shfl_xor_sync(0xffffffffu, verts, o);
  }
  if (lane == 0 && verts) {
    atomicAdd(&a.row->scanned[a.bin], scanned);
    atomicAdd(&a.row->verts[a.bin], verts);
  }
}

// one CTA per high-degree vertex ("delegates across warps and CTAs")
template <bool FIRST>
__global__ void __launch_bounds__(1024) k_lcc_scan_big(LccArgs a, const uint32_t* __restrict__ list,
                                                        const uint32_t* __restrict__ n_ptr) {
  __shared__ uint16_t s_lm[17];
  __shared__ uint32_t s_wcnt[32];
  __shared__ uint32_t s_heard[32];
  __shared__ uint32_t s_out;
  if (threadIdx.x < 17) s_lm[threadIdx.x] = c_pat.LMc[threadIdx.x];
  const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const uint32_t lt = lanemask_lt();
  const uint32_t n = *n_ptr;
  for (uint32_t idx = blockIdx.x; idx < n; idx += gridDim.x) {
    __syncthreads();
    const uint32_t v = list[idx];
    const uint32_t Tv = a.S[v];
    uint32_t d = FIRST ? a.deg[v] : a.adeg[v];
    if (Tv == 0) d = 0;
    const uint32_t NBv = nb_of(Tv);
    const uint64_t row = (uint64_t)a.rowblk[v] * 8;
    const uint32_t* __restrict__ src = FIRST ? a.col0 : a.colw;
    if (threadIdx.x == 0) s_out = 0;
    uint32_t heard = 0;
    const uint32_t per_pass = blockDim.x * 4;
    for (uint32_t p0 = 0; p0 < d; p0 += per_pass) {
      const uint32_t j0 = p0 + threadIdx.x * 4;
      uint4 q = make_uint4(PM_SENTINEL, PM_SENTINEL, PM_SENTINEL, PM_SENTINEL);
      if (j0 < d) q = *reinterpret_cast<const uint4*>(src + row + j0);
      uint32_t u[4] = {q.x, q.y, q.z, q.w};
      uint32_t m[4];
      bool keep[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const bool act = j0 + k < d;
        const uint32_t uu = u[k] & PM_IDMASK;
        m[k] = 0;
        if (act) m[k] = FIRST ? (uint32_t)s_lm[a.cls[uu]] : (uint32_t)a.S[uu];
      }
      uint32_t below = 0, wtotal = 0;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const bool act = j0 + k < d;
        const bool valid = (m[k] & NBv) != 0u;
        const bool pre = !FIRST && act && (u[k] >> 31);
        keep[k] = valid || pre;
        if (valid) heard |= m[k];
        const uint32_t b = __ballot_sync(0xffffffffu, keep[k]);
        below += __popc(b & lt);
        wtotal += __popc(b);
      }
      if (lane == 0) s_wcnt[wid] = wtotal;
      __syncthreads();  // every warp has read its slots of this pass
      uint32_t wbase = s_out;
      for (uint32_t w = 0; w < wid; ++w) wbase += s_wcnt[w];
      uint32_t pos = wbase + below;
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (keep[k]) a.colw[row + pos++] = u[k] & PM_IDMASK;
      __syncthreads();
      if (threadIdx.x == 0) {
        uint32_t t = s_out;
        for (uint32_t w = 0; w < nw; ++w) t += s_wcnt[w];
        s_out = t;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) heard |= __shfl_xor_sync(0xffffffffu, heard, o);
    if (lane == 0) s_heard[wid] = heard;
    __syncthreads();
    if (threadIdx.x == 0) {
      uint32_t h = 0;
      for (uint32_t w = 0; w < nw; ++w) h |= s_heard[w];
      const uint32_t T0 = FIRST ? Tv : (uint32_t)a.Tst[v];
      const uint32_t ts = Tv ? cover_of(T0, h) : 0u;
      a.Tst[v] = (uint16_t)ts;
      a.adeg[v] = s_out;
      if (ts == 0 && (FIRST ? h != 0u : Tv != 0u)) a.cnt->nf = 1u;
      if (Tv) {
        atomicAdd(&a.row->scanned[2], (unsigned long long)d);
        atomicAdd(&a.row->verts[2], 1ull);
      }
    }
  }
}

// ---------------------------------------------------------------------------
// commit: publish T_arr (ee.hpp:948), drop removed vertices (:941-946), bin the
// survivors by their new |E_v| into the next frontier and accumulate the row
// counts the reference writes after every superstep (:1112-1138).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock) k_lcc_commit(LccArgs a, const uint32_t* __restrict__ l0,
                                                        const uint32_t* __restrict__ l1,
                                                        const uint32_t* __restrict__ l2, uint32_t* n0,
                                                        uint32_t* n1, uint32_t* n2, int cur, int nxt) {
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t c0 = a.cnt->fr_n[cur][0], c1 = a.cnt->fr_n[cur][1], c2 = a.cnt->fr_n[cur][2];
  const uint32_t total = c0 + c1 + c2;
  unsigned long long nv = 0, ne = 0;
  uint32_t i0 = (blockIdx.x * blockDim.x + threadIdx.x) & ~31u;
  const uint32_t stride = gridDim.x * blockDim.x;
  for (; i0 < total; i0 += stride) {
    const uint32_t i = i0 + lane;
    bool alive = false;
    uint32_t v = 0, d = 0;
    if (i < total) {
      v = i < c0 ? l0[i] : (i < c0 + c1 ? l1[i - c0] : l2[i - c0 - c1]);
      const uint16_t ts = a.Tst[v];
      a.S[v] = ts;
      alive = ts != 0;
      d = a.adeg[v];
      if (alive) { nv++; ne += d; }
    }
    const int bin = d <= PM_SMALL_MAX ? 0 : (d <= PM_MID_MAX ? 1 : 2);
#pragma unroll
    for (int b = 0; b < 3; ++b) {
      const uint32_t m = __ballot_sync(0xffffffffu, alive && bin == b);
      if (m) {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(&a.cnt->fr_n[nxt][b], (uint32_t)__popc(m));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (alive && bin == b) {
          uint32_t* dst = b == 0 ? n0 : (b == 1 ? n1 : n2);
          dst[base + __popc(m & lanemask_lt())] = v;
        }
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    nv += __shfl_xor_sync(0xffffffffu, nv, o);
    ne += __shfl_xor_sync(0xffffffffu, ne, o);
  }
  if (lane == 0 && nv) {
    atomicAdd(&a.row->nv, nv);
    atomicAdd(&a.row->ne, ne);
  }
}

// counts after an NLCC constraint (beta.cpp:1094-1120): vertices still in the map
// and the sizes of their edge maps.  Frontier lists are left untouched; entries
// deactivated by NLCC are skipped by the next scan and dropped by its commit.
__global__ void __launch_bounds__(kBlock) k_count_alive(LccArgs a, const uint32_t* __restrict__ l0,
                                                         const uint32_t* __restrict__ l1,
                                                         const uint32_t* __restrict__ l2, int cur) {
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t c0 = a.cnt->fr_n[cur][0], c1 = a.cnt->fr_n[cur][1], c2 = a.cnt->fr_n[cur][2];
  const uint32_t total = c0 + c1 + c2;
  unsigned long long nv = 0, ne = 0;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const uint32_t v = i < c0 ? l0[i] : (i < c0 + c1 ? l1[i - c0] : l2[i - c0 - c1]);
    if (a.S[v]) { nv++; ne += a.adeg[v]; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    nv += __shfl_xor_sync(0xffffffffu, nv, o);
    ne += __shfl_xor_sync(0xffffffffu, ne, o);
  }
  if (lane == 0 && nv) {
    atomicAdd(&a.row->nv, nv);
    atomicAdd(&a.row->ne, ne);
  }
}

}  // namespace pm
