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
#include "pm_comm.cuh"

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

// Appends up to ITEMS values per thread to one of three lists with ONE atomic per list
// and block (a returning atomic on a single address retires ~1 per clock chip-wide, so
// per-warp atomics would serialise the whole grid).  Must be called by every thread of
// the block; order inside a block follows (item, thread).
template <int ITEMS>
__device__ __forceinline__ void block_bin_append(const bool (&flag)[ITEMS], const int (&bin)[ITEMS],
                                                 const uint32_t (&val)[ITEMS], uint32_t* d0, uint32_t* d1,
                                                 uint32_t* d2, uint32_t* counters) {
  __shared__ uint32_t s_w[kBlock / 32][3];
  __shared__ uint32_t s_base[3];
  const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const uint32_t lt = lanemask_lt();
  uint32_t off[ITEMS];
  uint32_t tot[3] = {0, 0, 0};
#pragma unroll
  for (int k = 0; k < ITEMS; ++k) {
    off[k] = 0;
#pragma unroll
    for (int b = 0; b < 3; ++b) {
      const uint32_t m = __ballot_sync(0xffffffffu, flag[k] && bin[k] == b);
      if (flag[k] && bin[k] == b) off[k] = tot[b] + __popc(m & lt);
      tot[b] += __popc(m);
    }
  }
  if (lane == 0) { s_w[w][0] = tot[0]; s_w[w][1] = tot[1]; s_w[w][2] = tot[2]; }
  __syncthreads();
  if (threadIdx.x < 3) {
    uint32_t t = 0;
    for (uint32_t i = 0; i < blockDim.x / 32; ++i) t += s_w[i][threadIdx.x];
    s_base[threadIdx.x] = t ? atomicAdd(&counters[threadIdx.x], t) : 0u;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < ITEMS; ++k)
    if (flag[k]) {
      const int b = bin[k];
      uint32_t base = s_base[b];
      for (uint32_t i = 0; i < w; ++i) base += s_w[i][b];
      uint32_t* dst = b == 0 ? d0 : (b == 1 ? d1 : d2);
      dst[base + off[k]] = val[k];
    }
  __syncthreads();
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
  const uint8_t* lab0;  // [Epad] label of the neighbour in col0 (labels < 64 only)
  uint8_t* labw;        // [Epad] same for colw, moved along by the row compaction
  DevCounters* cnt;
  RowStat* row;   // accumulator of this superstep
  int bin;        // degree bin this launch serves (row statistics)
  uint32_t base;  // first slot of this rank: frontier lists and rank-local arrays are indexed by slot - base
  int par;        // delta inbox the commit of this superstep publishes into
};

// Publishes "the mask of my vertex `slot` is now `mask`" to every peer: the pair is stored straight
// into the sender's region of each peer's delta inbox (NVLink stores, one position per change, reserved
// with one atomic per warp).  This replaces the per-message mailbox traffic of the reference
// (visitor_queue.hpp:395-434): peers only ever need the CHANGES of template_vertices.  All 32 lanes call.
__device__ __forceinline__ void publish_mask(bool changed, uint32_t slot, uint32_t mask, DevCounters* cnt, int par) {
  const uint32_t m = __ballot_sync(0xffffffffu, changed);
  if (m == 0u) return;
  const uint32_t lane = threadIdx.x & 31;
  const int leader = __ffs(m) - 1;
  uint32_t pos = 0;
  if ((int)lane == leader) pos = atomicAdd(&cnt->ndelta, (uint32_t)__popc(m));
  pos = __shfl_sync(0xffffffffu, pos, leader) + __popc(m & ((1u << lane) - 1u));
  if (changed && pos < c_peer.dcap) {
    const uint2 d = make_uint2(slot, mask);
    for (int g = 0; g < c_peer.G; ++g)
      if (g != c_peer.rank) c_peer.din[par][g][(uint64_t)c_peer.rank * c_peer.dcap + pos] = d;
  }
}

// applies the mask changes the peers published in the step that just ended (after comm_step)
__global__ void __launch_bounds__(kBlock) k_apply_deltas(uint16_t* __restrict__ S, const StepMsg* __restrict__ all, int par) {
  const int G = c_peer.G, me = c_peer.rank;
  for (int r = 0; r < G; ++r) {
    if (r == me) continue;
    const uint32_t n = min(all[r].ndelta, c_peer.dcap);
    const uint2* __restrict__ in = c_peer.din[par][me] + (uint64_t)r * c_peer.dcap;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
      const uint2 d = in[i];
      S[d.x] = (uint16_t)d.y;
    }
  }
}

// ---------------------------------------------------------------------------
// per-pattern initialisation (beta.cpp:484-492 + the label test every vertex
// performs in the first superstep, ee.hpp:371-380 / :523-546):
//   cls[v] = class of label[v] for EVERY vertex (the first scan gathers it);
//   candidates (label matches a template vertex, degree > 0) get S[v] =
//   labelmask(label[v]) and are appended to the frontier bin of their degree.
//   S of a non-candidate is never read: every later gather goes through an edge
//   map, and edge maps only ever hold candidates.
// SMALL = labels are bytes < 64 (lab8 + 64-entry class table), else u64 compare.
// ---------------------------------------------------------------------------
template <bool SMALL>
__global__ void __launch_bounds__(kBlock) k_init_state(const uint64_t* __restrict__ label,
                                                        const uint8_t* __restrict__ lab8,
                                                        const uint32_t* __restrict__ deg, uint64_t V,
                                                        uint8_t* __restrict__ cls, uint16_t* __restrict__ S,
                                                        uint32_t* fr_small, uint32_t* fr_mid, uint32_t* fr_big,
                                                        DevCounters* cnt, int buf) {
  __shared__ uint8_t s_cl[64];
  __shared__ uint16_t s_lm[17];
  if (threadIdx.x < 64) s_cl[threadIdx.x] = c_pat.cls_of_label[threadIdx.x];
  if (threadIdx.x < 17) s_lm[threadIdx.x] = c_pat.LMc[threadIdx.x];
  __syncthreads();
  constexpr int IT = 8;
  const uint64_t tile = (uint64_t)blockDim.x * IT;
  for (uint64_t base = (uint64_t)blockIdx.x * tile; base < V; base += (uint64_t)gridDim.x * tile) {
    bool cand[IT];
    int bin[IT];
    uint32_t val[IT];
#pragma unroll
    for (int k = 0; k < IT; ++k) {
      const uint64_t v = base + (uint64_t)k * blockDim.x + threadIdx.x;
      uint32_t c = PM_NOCLASS, d = 0;
      if (v < V) {
        if (SMALL) {
          c = s_cl[lab8[v] & 63];
        } else {
          const uint64_t lab = label[v];
#pragma unroll
          for (int q = 0; q < 16; ++q)
            if (q < c_pat.ncls && c_pat.clabel[q] == lab) c = q;
        }
        cls[v] = (uint8_t)c;
        if (c != PM_NOCLASS) {
          d = deg[v];
          if (d) S[v] = s_lm[c]; else c = PM_NOCLASS;
        }
      }
      cand[k] = c != PM_NOCLASS;
      bin[k] = d <= PM_SMALL_MAX ? 0 : (d <= PM_MID_MAX ? 1 : 2);
      val[k] = (uint32_t)v;
    }
    block_bin_append<IT>(cand, bin, val, fr_small, fr_mid, fr_big, &cnt->fr_n[buf][0]);
  }
}

// ---------------------------------------------------------------------------
// fused per-pattern initialisation + first-superstep signature filter (labels < 64).
// One streaming pass over all vertices: class from the byte label, candidate test, then — for
// candidates — the signature test.  In the first superstep every neighbour u of v sends
// labelmask(label[u]) (ee.hpp:519-561), so heard(v) depends only on WHICH labels occur among v's
// neighbours: heard(v) = OR { LM(l) : l in sig[v], LM(l) & NB(T_v) != 0 } — exactly what walking the
// row would compute.  A candidate whose T_state comes out empty leaves (or never enters) the map; it
// is settled here from 8 bytes of signature instead of its whole row (tables c_pat.req / c_pat.rl).
// Writes cls[v] and S[v] for EVERY vertex with coalesced stores (S = 0 unless the
// vertex survives the first superstep's cover test) and appends the survivors to the
// frontier bin of their degree.  Their rows are walked by the first scan afterwards.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock) k_init_filter(const uint8_t* __restrict__ lab8,
                                                         const uint32_t* __restrict__ deg,
                                                         const unsigned long long* __restrict__ sig, uint64_t V,
                                                         uint8_t* __restrict__ cls, uint16_t* __restrict__ S,
                                                         uint32_t* fr_small, uint32_t* fr_mid, uint32_t* fr_big,
                                                         DevCounters* cnt, int buf) {
  __shared__ uint8_t s_cl[64];
  __shared__ uint16_t s_lm[17];
  __shared__ unsigned long long s_rl[17];
  __shared__ uint32_t s_w[kBlock / 32][3];
  __shared__ uint32_t s_base[3];
  if (threadIdx.x < 64) s_cl[threadIdx.x] = c_pat.cls_of_label[threadIdx.x];
  if (threadIdx.x < 17) { s_lm[threadIdx.x] = c_pat.LMc[threadIdx.x]; s_rl[threadIdx.x] = c_pat.rl[threadIdx.x]; }
  __syncthreads();
  constexpr int VPT = 16;  // vertices per thread: one uint4 of byte labels, four uint4 of degrees
  const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const uint64_t tile = (uint64_t)blockDim.x * VPT;
  unsigned long long ncand = 0;
  bool any_removed = false;
  for (uint64_t base = (uint64_t)blockIdx.x * tile; base < V; base += (uint64_t)gridDim.x * tile) {
    const uint64_t v0 = base + (uint64_t)threadIdx.x * VPT;
    uint32_t k0 = 0, k1 = 0, k2 = 0;  // bit j: vertex v0 + j survives the first superstep, by degree bin
    if (v0 + VPT <= V) {
      const uint4 l16 = *reinterpret_cast<const uint4*>(lab8 + v0);
      const uint32_t lw[4] = {l16.x, l16.y, l16.z, l16.w};
      uint32_t cw[4];
      uint32_t sw[8];
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const uint4 d4 = *reinterpret_cast<const uint4*>(deg + v0 + 4 * g);
        const uint32_t dd[4] = {d4.x, d4.y, d4.z, d4.w};
        cw[g] = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int j = 4 * g + k;
          const uint32_t c = s_cl[(lw[g] >> (8 * k)) & 63u];
          const uint32_t d = dd[k];
          uint32_t lm = d ? (uint32_t)s_lm[c] : 0u;
          if (lm) {
            const unsigned long long sg = sig[v0 + j];
            uint32_t surv = 0;
            for (uint32_t rest = lm; rest; rest &= rest - 1) {
              const int pp = __ffs(rest) - 1;
              const unsigned long long rq = c_pat.req[pp];
              if (rq != 0ull && (sg & rq) == rq) surv = 1;
            }
            any_removed = any_removed || (!surv && (sg & s_rl[c]) != 0ull);  // entered the map and left it (ee.hpp:941-946)
            ncand++;
            if (surv) {
              if (d <= PM_SMALL_MAX) k0 |= 1u << j; else if (d <= PM_MID_MAX) k1 |= 1u << j; else k2 |= 1u << j;
            } else {
              lm = 0;
            }
          }
          cw[g] |= c << (8 * k);
          if (k & 1) sw[j >> 1] |= lm << 16; else sw[j >> 1] = lm;
        }
      }
      *reinterpret_cast<uint4*>(cls + v0) = make_uint4(cw[0], cw[1], cw[2], cw[3]);
      *reinterpret_cast<uint4*>(S + v0) = make_uint4(sw[0], sw[1], sw[2], sw[3]);
      *reinterpret_cast<uint4*>(S + v0 + 8) = make_uint4(sw[4], sw[5], sw[6], sw[7]);
    } else {
      for (int j = 0; j < VPT && v0 + j < V; ++j) {  // ragged tail
        const uint64_t v = v0 + j;
        const uint32_t c = s_cl[lab8[v] & 63];
        const uint32_t d = deg[v];
        uint32_t lm = d ? (uint32_t)s_lm[c] : 0u;
        if (lm) {
          const unsigned long long sg = sig[v];
          uint32_t surv = 0;
          for (uint32_t rest = lm; rest; rest &= rest - 1) {
            const unsigned long long rq = c_pat.req[__ffs(rest) - 1];
            if (rq != 0ull && (sg & rq) == rq) surv = 1;
          }
          any_removed = any_removed || (!surv && (sg & s_rl[c]) != 0ull);
          ncand++;
          if (surv) {
            if (d <= PM_SMALL_MAX) k0 |= 1u << j; else if (d <= PM_MID_MAX) k1 |= 1u << j; else k2 |= 1u << j;
          } else {
            lm = 0;
          }
        }
        cls[v] = (uint8_t)c;
        S[v] = (uint16_t)lm;
      }
    }
    // block-aggregated append of the survivors: one atomic per bin and tile
    const uint32_t n0 = __popc(k0), n1 = __popc(k1), n2 = __popc(k2);
    uint32_t i0 = n0, i1 = n1, i2 = n2;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t0 = __shfl_up_sync(0xffffffffu, i0, o);
      const uint32_t t1 = __shfl_up_sync(0xffffffffu, i1, o);
      const uint32_t t2 = __shfl_up_sync(0xffffffffu, i2, o);
      if (lane >= (uint32_t)o) { i0 += t0; i1 += t1; i2 += t2; }
    }
    if (lane == 31) { s_w[w][0] = i0; s_w[w][1] = i1; s_w[w][2] = i2; }
    __syncthreads();
    if (threadIdx.x < 3) {
      uint32_t t = 0;
      for (uint32_t i = 0; i < blockDim.x / 32; ++i) t += s_w[i][threadIdx.x];
      s_base[threadIdx.x] = t ? atomicAdd(&cnt->fr_n[buf][threadIdx.x], t) : 0u;
    }
    __syncthreads();
    if (k0 | k1 | k2) {
      uint32_t o0 = s_base[0] + i0 - n0, o1 = s_base[1] + i1 - n1, o2 = s_base[2] + i2 - n2;
      for (uint32_t i = 0; i < w; ++i) { o0 += s_w[i][0]; o1 += s_w[i][1]; o2 += s_w[i][2]; }
      for (uint32_t rest = k0; rest; rest &= rest - 1) fr_small[o0++] = (uint32_t)(v0 + __ffs(rest) - 1);
      for (uint32_t rest = k1; rest; rest &= rest - 1) fr_mid[o1++] = (uint32_t)(v0 + __ffs(rest) - 1);
      for (uint32_t rest = k2; rest; rest &= rest - 1) fr_big[o2++] = (uint32_t)(v0 + __ffs(rest) - 1);
    }
    __syncthreads();
  }
  if (any_removed) cnt->nf_init = 1u;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ncand += __shfl_xor_sync(0xffffffffu, ncand, o);
  if (lane == 0 && ncand) atomicAdd(&cnt->filtered_init, ncand);
}

// ---------------------------------------------------------------------------
// scan: GROUP lanes walk the active adjacency of one vertex with uint4 loads
// (4 slots per lane and pass), gather the neighbour masks, OR the heard masks and
// compact the surviving neighbours to the front of the row.
//   FIRST = first superstep of the first iteration: walk the pristine adjacency
//   col0 (all deg[v] slots), neighbour mask = labelmask via the class array
//   (ee.hpp:519-561 sender, :368-404 receiver); otherwise walk keys(E_v) in colw.
// ---------------------------------------------------------------------------
//   STREAM = the label of every neighbour travels next to its id (lab0 / labw,
//   labels < 64): the first superstep then needs no gather at all — ids and labels
//   are both streamed — and later compactions keep labw aligned with colw for NLCC.
template <int GROUP, bool FIRST, bool STREAM>
__global__ void __launch_bounds__(kBlock) k_lcc_scan(LccArgs a, const uint32_t* __restrict__ list,
                                                      const uint32_t* __restrict__ n_ptr) {
  __shared__ uint16_t s_lm[17];
  __shared__ uint16_t s_lml[64];  // label value -> labelmask
  if (threadIdx.x < 17) s_lm[threadIdx.x] = c_pat.LMc[threadIdx.x];
  if (threadIdx.x < 64) s_lml[threadIdx.x] = c_pat.LMc[c_pat.cls_of_label[threadIdx.x]];
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
      Tv = a.S[v + a.base];
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
      uint32_t l4 = 0;
      if (j0 < d) {
        q = *reinterpret_cast<const uint4*>(src + row + j0);
        if (STREAM) l4 = *reinterpret_cast<const uint32_t*>((FIRST ? a.lab0 : a.labw) + row + j0);
      }
      uint32_t u[4] = {q.x, q.y, q.z, q.w};
      uint32_t m[4];
      bool keep[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const bool act = j0 + k < d;
        const uint32_t uu = u[k] & PM_IDMASK;
        m[k] = 0;
        if (act) {
          if (FIRST) m[k] = STREAM ? (uint32_t)s_lml[(l4 >> (8 * k)) & 63u] : (uint32_t)s_lm[a.cls[uu]];
          else m[k] = (uint32_t)a.S[uu];
        }
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
        if (keep[k]) {
          a.colw[row + pos] = u[k] & PM_IDMASK;
          if (STREAM) a.labw[row + pos] = (uint8_t)(l4 >> (8 * k));
          ++pos;
        }
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
template <bool FIRST, bool STREAM>
__global__ void __launch_bounds__(1024) k_lcc_scan_big(LccArgs a, const uint32_t* __restrict__ list,
                                                        const uint32_t* __restrict__ n_ptr) {
  __shared__ uint16_t s_lm[17];
  __shared__ uint16_t s_lml[64];
  if (threadIdx.x < 64) s_lml[threadIdx.x] = c_pat.LMc[c_pat.cls_of_label[threadIdx.x]];
  __shared__ uint32_t s_wcnt[32];
  __shared__ uint32_t s_heard[32];
  if (threadIdx.x < 17) s_lm[threadIdx.x] = c_pat.LMc[threadIdx.x];
  const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const uint32_t lt = lanemask_lt();
  const uint32_t n = *n_ptr;
  for (uint32_t idx = blockIdx.x; idx < n; idx += gridDim.x) {
    __syncthreads();
    const uint32_t v = list[idx];
    const uint32_t Tv = a.S[v + a.base];
    uint32_t d = FIRST ? a.deg[v] : a.adeg[v];
    if (Tv == 0) d = 0;
    const uint32_t NBv = nb_of(Tv);
    const uint64_t row = (uint64_t)a.rowblk[v] * 8;
    const uint32_t* __restrict__ src = FIRST ? a.col0 : a.colw;
    uint32_t outp = 0;  // slots kept so far (every thread tracks the same value)
    uint32_t heard = 0;
    const uint32_t per_pass = blockDim.x * 4;
    for (uint32_t p0 = 0; p0 < d; p0 += per_pass) {
      const uint32_t j0 = p0 + threadIdx.x * 4;
      uint4 q = make_uint4(PM_SENTINEL, PM_SENTINEL, PM_SENTINEL, PM_SENTINEL);
      uint32_t l4 = 0;
      if (j0 < d) {
        q = *reinterpret_cast<const uint4*>(src + row + j0);
        if (STREAM) l4 = *reinterpret_cast<const uint32_t*>((FIRST ? a.lab0 : a.labw) + row + j0);
      }
      uint32_t u[4] = {q.x, q.y, q.z, q.w};
      uint32_t m[4];
      bool keep[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const bool act = j0 + k < d;
        const uint32_t uu = u[k] & PM_IDMASK;
        m[k] = 0;
        if (act) {
          if (FIRST) m[k] = STREAM ? (uint32_t)s_lml[(l4 >> (8 * k)) & 63u] : (uint32_t)s_lm[a.cls[uu]];
          else m[k] = (uint32_t)a.S[uu];
        }
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
      uint32_t wbase = outp, ptotal = 0;
      for (uint32_t w = 0; w < nw; ++w) {
        const uint32_t cw = s_wcnt[w];
        if (w < wid) wbase += cw;
        ptotal += cw;
      }
      uint32_t pos = wbase + below;
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (keep[k]) {
          a.colw[row + pos] = u[k] & PM_IDMASK;
          if (STREAM) a.labw[row + pos] = (uint8_t)(l4 >> (8 * k));
          ++pos;
        }
      outp += ptotal;
      __syncthreads();  // s_wcnt may be overwritten by the next pass only after everyone has read it
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
      a.adeg[v] = outp;
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
  constexpr int IT = 4;
  const uint32_t tile = blockDim.x * IT;
  for (uint32_t base = blockIdx.x * tile; base < total; base += gridDim.x * tile) {
    bool alive[IT];
    int bin[IT];
    uint32_t val[IT];
#pragma unroll
    for (int k = 0; k < IT; ++k) {
      const uint32_t i = base + k * blockDim.x + threadIdx.x;
      alive[k] = false;
      bin[k] = 0;
      val[k] = 0;
      bool changed = false;
      uint32_t cslot = 0, cmask = 0;
      if (i < total) {
        const uint32_t v = i < c0 ? l0[i] : (i < c0 + c1 ? l1[i - c0] : l2[i - c0 - c1]);
        const uint16_t ts = a.Tst[v];
        if (c_peer.G > 1) changed = ts != a.S[v + a.base];  // peers only need the changes
        a.S[v + a.base] = ts;
        cslot = v + a.base;
        cmask = ts;
        alive[k] = ts != 0;
        const uint32_t d = a.adeg[v];
        if (alive[k]) { nv++; ne += d; }
        bin[k] = d <= PM_SMALL_MAX ? 0 : (d <= PM_MID_MAX ? 1 : 2);
        val[k] = v;
      }
      if (c_peer.G > 1) publish_mask(changed, cslot, cmask, a.cnt, a.par);
    }
    block_bin_append<IT>(alive, bin, val, n0, n1, n2, &a.cnt->fr_n[nxt][0]);
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
    if (a.S[v + a.base]) { nv++; ne += a.adeg[v]; }
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
