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
//
// Frontier.  The vertices still in the vertex_state_map are kept as a list of
// 16-byte ENTRIES {local row, row start (sectors), |E_v|, T_state}: a scan streams
// its entries (coalesced), needs no dependent gather before it can fetch the row,
// and writes its result (new |E_v|, new T_state) back into the entry.  A warp takes 32
// consecutive entries and deals the 4-slot chunks of their rows to its lanes, 32 chunks
// per pass with the next pass prefetched; rows above PM_MID_MAX slots get a CTA each.
#pragma once

#include "pm_common.cuh"
#include "pm_comm.cuh"

namespace pm {

__constant__ PatConst c_pat;

// template vertices a valid parent may carry: neighbours over mandatory AND optional edges
// (ee.hpp:673-722; approximate_pattern_matching/local_constraint_checking.hpp:641-651)
__device__ __forceinline__ uint32_t nb_of(uint32_t T) {
  uint32_t r = 0;
#pragma unroll
  for (int p = 0; p < 16; ++p)
    if ((T >> p) & 1u) r |= (uint32_t)c_pat.N[p] | c_pat.No[p];
  return r;
}

// bits p of T that keep their place: exact pattern — the template neighbourhood is non-empty and entirely heard
// (ee.hpp:901-939); approximate pattern — every mandatory neighbour heard (none required: fine), and where a minimum
// optional edge count is set, every optional neighbour heard and at least that many of them
// (approximate_pattern_matching/local_constraint_checking.hpp:1062-1113)
__device__ __forceinline__ uint32_t cover_of(uint32_t T, uint32_t heard) {
  uint32_t r = 0;
#pragma unroll
  for (int p = 0; p < 16; ++p) {
    const uint32_t need = c_pat.N[p];
    bool ok = (need & ~heard) == 0u && (need != 0u || c_pat.approx);
    if (c_pat.approx && c_pat.min_opt[p]) {
      const uint32_t opt = c_pat.No[p];
      ok = ok && (opt & ~heard) == 0u && __popc(opt) >= (int)c_pat.min_opt[p];
    }
    if (((T >> p) & 1u) && ok) r |= 1u << p;
  }
  return r;
}

__device__ __forceinline__ uint32_t lanemask_lt() {
  uint32_t m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}

// Appends up to ITEMS entries per thread to one of two lists with ONE atomic per list
// and block (a returning atomic on a single address retires ~1 per clock chip-wide, so
// per-warp atomics would serialise the whole grid).  Must be called by every thread of
// the block; order inside a block follows (item, thread).
template <int ITEMS>
__device__ __forceinline__ void block_append2(const bool (&flag)[ITEMS], const int (&bin)[ITEMS],
                                              const uint4 (&val)[ITEMS], uint4* d0, uint4* d1, uint32_t* counters) {
  __shared__ uint32_t s_w[kBlock / 32][2];
  __shared__ uint32_t s_base[2];
  const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const uint32_t lt = lanemask_lt();
  uint32_t off[ITEMS];
  uint32_t tot[2] = {0, 0};
#pragma unroll
  for (int k = 0; k < ITEMS; ++k) {
    off[k] = 0;
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      const uint32_t m = __ballot_sync(0xffffffffu, flag[k] && bin[k] == b);
      if (flag[k] && bin[k] == b) off[k] = tot[b] + __popc(m & lt);
      tot[b] += __popc(m);
    }
  }
  if (lane == 0) { s_w[w][0] = tot[0]; s_w[w][1] = tot[1]; }
  __syncthreads();
  if (threadIdx.x < 2) {
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
      (b == 0 ? d0 : d1)[base + off[k]] = val[k];
    }
  __syncthreads();
}

struct LccArgs {
  const uint32_t* col0;
  uint32_t* colw;
  uint16_t* S;
  uint32_t* adeg;
  const uint8_t* cls;
  const uint8_t* lab0;  // [Epad] label of the neighbour in col0 (labels < 64 only)
  DevCounters* cnt;
  RowStat* row;   // accumulator of this superstep
  uint32_t base;  // first compact id of this rank: rank-local arrays (adeg, rowc) are indexed by cid - base
  int par;        // delta inbox the commit of this superstep publishes into
  const uint2* fwx;        // slot -> compact id, per 16 slots: {cid of the word's first survivor, survivor bits}
  const uint32_t* rowc;    // [n_c + 1] row start (sectors) in the DENSE working adjacency, by local compact id
  uint32_t col_shift;      // packed labels: a col0 slot is (label << col_shift) | id
  int typed;               // fwx[w].y holds 2-bit T_state numbers per slot (see PatConst::tsub) instead of survivor bits
  const uint8_t* clsc;     // label class by compact id
  const uint8_t* hubc;     // controller rank + 1 of a hub, by local compact id (null: no delegates)
};

// frontier entry: x = compact id, y = row start in sectors (PM_TOMB: the row is in the big-row list),
// z = |E_v| (deg(v) before the first scan), w = T_state (vertex_state.template_vertices)
__device__ __forceinline__ int bin_of(uint32_t d) { return d <= PM_MID_MAX ? 0 : 1; }

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
// Compact ids.  After the first-superstep filter only a few % of the vertices are left, but their
// masks would stay scattered over a V-sized array: every later gather S[u] would pull a 32-byte
// sector from HBM for 2 useful bytes.  The survivors are therefore renumbered densely, in vertex
// order, and EVERYTHING after the filter — masks, classes, |E_v|, row starts, the working adjacency,
// tokens, hash keys — is indexed by that compact id (cid).  The whole mask array then fits in L2.
//   vid[cid] = slot of the vertex.  slot -> cid: survivors are numbered in slot order, so
//   cid(s) = fwx[s / 16].x + popc(fwx[s / 16].y & bits below s), fwx[w] = {cid of the word's first survivor,
//   survivor bits of its 16 vertices} — 8 bytes per 16 vertices, L2 resident, one load per lookup.  It is
//   assembled from fw[w] = (survivors before the word inside its 4096-slot tile) << 16 | bits (pass 1) and
//   tb[tile] = cid of the tile's first survivor (pass 2).
// With several GPUs rank g owns the cids [off[g], off[g+1]) (c_peer.off), numbered by (rank, slot);
// fw and tb are all-gathered (slot ranges are rank-contiguous and tile aligned).
// The first scan copies the neighbours it keeps as SLOTS (a kept neighbour that did not survive the
// filter is in E_v for the first superstep's edge count, ee.hpp:791-813); the second scan of the call
// (XLATE) renames them to cids while it walks the by then short rows and drops the non-survivors, whose
// masks are zero.  A pattern with a single superstep per LCC call (pattern_stat diameter 1) has no second
// scan: its filter is switched off (every label-matching vertex is numbered) and a rename-only pass follows.
//
// per-pattern initialisation (beta.cpp:484-492) + the label test every vertex performs in the first
// superstep (ee.hpp:371-380 / :523-546) + (labels < 64) the signature filter, in three passes:
//   k_init_flags   one streaming pass over the local vertices: survivor bits + in-tile prefixes (fw), count per tile
//   k_init_scan    exclusive prefix of the tile counts (one block)
//   k_init_assign  per-cid state (mask, class, slot) for EVERY survivor, frontier entries for the local ones
// Signature filter: in the first superstep every neighbour u of v sends labelmask(label[u])
// (ee.hpp:519-561), so heard(v) depends only on WHICH labels occur among v's neighbours:
// heard(v) = OR { LM(l) : l in sig[v], LM(l) & NB(T_v) != 0 } — exactly what walking the row would
// compute.  A candidate whose T_state comes out empty leaves (or never enters) the map; it is settled
// from 8 bytes of signature instead of its whole row (tables c_pat.req / c_pat.rl).
// ---------------------------------------------------------------------------
#define PM_TILE (kBlock * 16)     // vertices per tile: 16 per thread
#define PM_TOMB 0xFFFFFFFFu       // entry.y of a main-list slot whose row lives in the big-row list

// SMALL: labels are bytes < 64 (lab8, signature filter); else u64 labels, every label-matching
// vertex of degree > 0 is a "survivor" and cls[] (by slot) is filled for the first scan's gathers
template <bool SMALL>
__global__ void __launch_bounds__(kBlock) k_init_flags(const uint8_t* __restrict__ lab8, const uint64_t* __restrict__ label,
                                                        const uint32_t* __restrict__ deg,
                                                        const unsigned long long* __restrict__ sig, uint64_t V,
                                                        uint8_t* __restrict__ cls, uint32_t* __restrict__ fw,
                                                        uint32_t* __restrict__ tile_cnt, DevCounters* cnt, int use_sig) {
  __shared__ uint8_t s_cl[64];
  __shared__ uint16_t s_lm[17];
  __shared__ unsigned long long s_rl[17];
  __shared__ uint32_t s_w[kBlock / 32];
  if (threadIdx.x < 64) s_cl[threadIdx.x] = c_pat.cls_of_label[threadIdx.x];
  if (threadIdx.x < 17) { s_lm[threadIdx.x] = c_pat.LMc[threadIdx.x]; s_rl[threadIdx.x] = c_pat.rl[threadIdx.x]; }
  __syncthreads();
  const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const uint64_t n_tiles = (V + PM_TILE - 1) / PM_TILE;
  unsigned long long ncand = 0;
  bool any_removed = false;
  for (uint64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const uint64_t v0 = tile * PM_TILE + (uint64_t)threadIdx.x * 16;
    uint32_t bits = 0;
    if (SMALL && v0 + 16 <= V) {
      const uint4 l16 = *reinterpret_cast<const uint4*>(lab8 + v0);
      const uint32_t lw[4] = {l16.x, l16.y, l16.z, l16.w};
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        // with the signature filter on the degrees need not be read: an isolated vertex has an empty signature
        uint4 d4 = make_uint4(1u, 1u, 1u, 1u);
        if (!use_sig) d4 = *reinterpret_cast<const uint4*>(deg + v0 + 4 * g);
        const uint32_t dd[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint32_t c = s_cl[(lw[g] >> (8 * k)) & 63u];
          const uint32_t lm = dd[k] ? (uint32_t)s_lm[c] : 0u;
          if (lm) {
            const unsigned long long sg = sig[v0 + 4 * g + k];
            uint32_t surv = use_sig ? 0u : 1u;
            for (uint32_t rest = lm; rest; rest &= rest - 1) {
              const int pp = __ffs(rest) - 1;
              const unsigned long long rq = c_pat.req[pp];
              // approximate pattern: nothing may be required (rq == 0), but only a vertex that heard a valid
              // neighbour ever enters the map; never[pp]: the minimum optional edge count cannot be met
              if (c_pat.approx ? ((sg & rq) == rq && (sg & s_rl[c]) != 0ull && !((c_pat.never >> pp) & 1u))
                               : (rq != 0ull && (sg & rq) == rq)) surv = 1;
            }
            any_removed = any_removed || (!surv && (sg & s_rl[c]) != 0ull);  // entered the map and left it (ee.hpp:941-946)
            ncand++;
            bits |= surv << (4 * g + k);
          }
        }
      }
    } else {
      for (int j = 0; j < 16 && v0 + j < V; ++j) {
        const uint64_t v = v0 + j;
        const uint32_t d = deg[v];
        if (SMALL) {
          const uint32_t c = s_cl[lab8[v] & 63];
          const uint32_t lm = d ? (uint32_t)s_lm[c] : 0u;
          if (lm) {
            const unsigned long long sg = sig[v];
            uint32_t surv = use_sig ? 0u : 1u;
            for (uint32_t rest = lm; rest; rest &= rest - 1) {
              const int pp = __ffs(rest) - 1;
              const unsigned long long rq = c_pat.req[pp];
              if (c_pat.approx ? ((sg & rq) == rq && (sg & s_rl[c]) != 0ull && !((c_pat.never >> pp) & 1u))
                               : (rq != 0ull && (sg & rq) == rq)) surv = 1;
            }
            any_removed = any_removed || (!surv && (sg & s_rl[c]) != 0ull);
            ncand++;
            bits |= surv << j;
          }
        } else {
          const uint64_t lab = label[v];
          uint32_t c = PM_NOCLASS;
#pragma unroll
          for (int q = 0; q < 16; ++q)
            if (q < c_pat.ncls && c_pat.clabel[q] == lab) c = q;
          cls[v] = (uint8_t)c;
          if (c != PM_NOCLASS && d) bits |= 1u << j;
        }
      }
    }
    // survivors before this thread's 16 vertices inside the tile
    const uint32_t n = __popc(bits);
    uint32_t incl = n;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= (uint32_t)o) incl += t;
    }
    if (lane == 31) s_w[w] = incl;
    __syncthreads();
    uint32_t before = incl - n;
    for (uint32_t k = 0; k < w; ++k) before += s_w[k];
    if (v0 < V) fw[v0 / 16] = (before << 16) | bits;
    if (threadIdx.x == kBlock - 1) tile_cnt[tile] = before + n;
    __syncthreads();
  }
  if (any_removed) cnt->nf_init = 1u;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ncand += __shfl_xor_sync(0xffffffffu, ncand, o);
  if (lane == 0 && ncand) atomicAdd(&cnt->filtered_init, ncand);
}

// exclusive prefix of the tile counts (in place) with one block; publishes the number of survivors
__global__ void __launch_bounds__(1024) k_init_scan(uint32_t* __restrict__ tile_cnt, uint64_t n_tiles, DevCounters* cnt, int buf) {
  __shared__ uint32_t s_w[32];
  __shared__ uint32_t s_carry;
  const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  for (uint64_t base = 0; base < n_tiles; base += blockDim.x) {
    const uint64_t i = base + threadIdx.x;
    const uint32_t x = i < n_tiles ? tile_cnt[i] : 0u;
    uint32_t incl = x;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= (uint32_t)o) incl += t;
    }
    if (lane == 31) s_w[w] = incl;
    __syncthreads();
    uint32_t wbase = 0;
    for (uint32_t k = 0; k < w; ++k) wbase += s_w[k];
    const uint32_t carry = s_carry;
    if (i < n_tiles) tile_cnt[i] = carry + wbase + incl - x;
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) s_carry = carry + wbase + incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    cnt->n_c = s_carry;
    cnt->fr_n[buf][0] = s_carry;  // the main list holds one entry per survivor, in cid order
  }
}

// several GPUs: tb of every rank's tiles = that rank's first cid + its local tile prefix
__global__ void k_tile_offsets(uint32_t* __restrict__ tb, uint32_t tiles_per_rank) {
  const uint32_t n = tiles_per_rank * (uint32_t)c_peer.G;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) tb[i] += c_peer.off[i / tiles_per_rank];
}

// compact id of slot u if it survived the filter, else PM_SENTINEL: one 8-byte load from an L2 resident table
__device__ __forceinline__ uint32_t cid_of_slot(const uint2* __restrict__ fwx, uint32_t u) {
  const uint2 w = fwx[u >> 4];
  const uint32_t b = u & 15u;
  if (!((w.y >> b) & 1u)) return PM_SENTINEL;
  return w.x + __popc(w.y & ((1u << b) - 1u));
}

// typed table: {cid of the word's first survivor, 2-bit T_state numbers of its 16 slots}; returns the number t of slot
// u (0: no survivor) and its compact id
__device__ __forceinline__ uint32_t typed_lookup(const uint2 w, uint32_t u, uint32_t& cid) {
  const uint32_t j2 = (u & 15u) * 2u;
  const uint32_t nz = (w.y | (w.y >> 1)) & 0x55555555u;  // one bit per survivor
  cid = w.x + __popc(nz & ((1u << j2) - 1u));
  return (w.y >> j2) & 3u;
}

// n_slots: every slot of every rank (replicated per-cid state); [own_lo, own_hi): this rank's slots, which also
// get their row start and frontier entry.  deg / rowblk are indexed by slot - own_lo.
template <bool SMALL>
__global__ void __launch_bounds__(kBlock) k_init_assign(const uint8_t* __restrict__ lab8, const uint8_t* __restrict__ cls,
                                                         const uint32_t* __restrict__ deg, const uint32_t* __restrict__ rowblk,
                                                         const uint32_t* __restrict__ fw, const uint32_t* __restrict__ tb,
                                                         uint64_t n_slots, uint32_t own_lo, uint32_t own_hi,
                                                         uint16_t* __restrict__ S, uint8_t* __restrict__ clsc,
                                                         uint32_t* __restrict__ vid, uint32_t* __restrict__ adeg,
                                                         uint32_t* __restrict__ rowc, uint2* __restrict__ fwx,
                                                         const unsigned long long* __restrict__ sig,
                                                         uint4* fr_main, uint4* fr_big, DevCounters* cnt, int buf,
                                                         int typed, const uint8_t* __restrict__ hub_ctl = nullptr,
                                                         uint8_t* __restrict__ hubc = nullptr) {
  __shared__ uint8_t s_cl[64];
  __shared__ uint16_t s_lm[17];
  if (threadIdx.x < 64) s_cl[threadIdx.x] = c_pat.cls_of_label[threadIdx.x];
  if (threadIdx.x < 17) s_lm[threadIdx.x] = c_pat.LMc[threadIdx.x];
  __syncthreads();
  // A warp takes 32 words = 512 consecutive slots (never across a tile).  Their survivors have CONSECUTIVE compact
  // ids, so after the lanes have listed the survivors' slots in shared memory, lane j handles survivor j: the
  // per-cid stores (masks, classes, slots, row starts, frontier entries) are coalesced and the lanes stay busy
  // however the survivors are spread over the words.
  __shared__ uint32_t s_slot[kBlock / 32][512];
  const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const uint32_t off_me = c_peer.off[c_peer.rank];
  const uint64_t n_words = (n_slots + 15) / 16;
  const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  for (uint64_t w0 = warp * 32; w0 < n_words; w0 += nwarps * 32) {
    const uint64_t wi = w0 + lane;
    uint32_t bits = 0, cid0 = 0;
    if (wi < n_words) {
      const uint32_t w = fw[wi];
      bits = w & 0xFFFFu;
      cid0 = tb[wi >> 8] + (w >> 16);
      // the table the renaming scan reads (cid_of_slot); typed: the survivors OR their T_state numbers in below
      fwx[wi] = make_uint2(cid0, typed ? 0u : bits);
    }
    const uint32_t cnt_w = __popc(bits);
    uint32_t incl = cnt_w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= (uint32_t)o) incl += t;
    }
    const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
    const uint32_t cid_first = __shfl_sync(0xffffffffu, cid0, 0);  // cid of the window's first survivor
    if (total == 0u) continue;
    __syncwarp();
    uint32_t at = incl - cnt_w;
    for (uint32_t rest = bits; rest; rest &= rest - 1, ++at) s_slot[wid][at] = (uint32_t)(wi * 16) + (__ffs(rest) - 1);
    __syncwarp();
    for (uint32_t j = lane; j < total; j += 32) {
      const uint32_t slot = s_slot[wid][j], cid = cid_first + j;
      const uint32_t c = SMALL ? (uint32_t)s_cl[lab8[slot] & 63] : (uint32_t)cls[slot];
      const uint32_t lm = s_lm[c];
      vid[cid] = slot;
      S[cid] = (uint16_t)lm;
      clsc[cid] = (uint8_t)c;
      if (slot >= own_lo && slot < own_hi) {
        const uint32_t lcid = cid - off_me, d = deg[slot - own_lo], rb = rowblk[slot - own_lo];
        adeg[lcid] = d;
        if (hubc) hubc[lcid] = hub_ctl[slot - own_lo];
        rowc[lcid] = (d + 7u) >> 3;  // |E_v| can only shrink: the prefix of these is the row start in the dense working adjacency
        // behind the signature filter (sig != null) the T_state the first superstep ends with is known here:
        // heard(v) = OR of the labelmasks of the valid labels in v's signature (see the header above), so the
        // first scan only has to build the edge map
        uint32_t tw = lm;
        if (sig) {
          const unsigned long long sg = sig[slot - own_lo];
          const uint32_t NBv = nb_of(lm);
          uint32_t heard = 0;
#pragma unroll
          for (int q = 0; q < 16; ++q)
            if (q < c_pat.ncls && c_pat.clabel[q] < 64ull && ((uint32_t)c_pat.LMc[q] & NBv) && ((sg >> c_pat.clabel[q]) & 1ull))
              heard |= c_pat.LMc[q];
          tw = cover_of(lm, heard);
        }
        if (typed) {  // T_state number of this vertex: 2 bits at its place in the word's entry
          uint32_t t = 1;
          if (tw == (uint32_t)c_pat.tsub[c][2]) t = 2;
          if (tw == (uint32_t)c_pat.tsub[c][3]) t = 3;
          atomicOr(&fwx[slot >> 4].y, t << ((slot & 15u) * 2u));
        }
        if (d <= PM_MID_MAX) {
          fr_main[lcid] = make_uint4(cid, rb, d, tw);
        } else {
          fr_main[lcid] = make_uint4(cid, PM_TOMB, 0u, 0u);
          fr_big[atomicAdd(&cnt->fr_n[buf][1], 1u)] = make_uint4(cid, rb, d, tw);
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------
// scan of the main list: every warp takes 32 consecutive entries.
//   FIRST = first superstep of the first iteration: walk the pristine adjacency col0 (all deg[v] slots); a
//   neighbour's mask is the labelmask of its label (ee.hpp:519-561 sender, :368-404 receiver) and the kept
//   neighbours are COPIED into the working adjacency; otherwise walk keys(E_v) in colw and compact in place.
//   STREAM (FIRST only) != 0: labels are bytes < 64 and the label of every neighbour travels with its id — 2: packed
//   into the high bits of the col0 slot itself (one stream), 1: in the parallel byte array lab0 (ids and labels do
//   not fit 32 bits together): validity is one bit test against the row's valid-label set, no gather at all.
//   HEARD = accumulate heard(v) and derive T_state from it.  Off only for the first scan behind the signature
//   filter, whose entries already carry the T_state the first superstep ends with (k_init_assign).
//   XLATE (with !FIRST) = the first scan after the first superstep: the rows still hold SLOTS (the first scan
//   copies them as they are, so that its row walks carry no dependent loads); this scan renames every
//   neighbour to its compact id while it walks the — by now short — rows, dropping neighbours that did not
//   survive the filter (their masks are zero: they would be dropped here anyway).
//   xlate_only: rename and nothing else (a pattern whose LCC call has a single superstep).
// ---------------------------------------------------------------------------

// labels a neighbour may carry to be valid for a vertex with template neighbourhood NBv (labels < 64)
__device__ __forceinline__ unsigned long long valid_labels(uint32_t NBv) {
  unsigned long long r = 0;
#pragma unroll
  for (int q = 0; q < 16; ++q)
    if (q < c_pat.ncls && c_pat.clabel[q] < 64ull && ((uint32_t)c_pat.LMc[q] & NBv)) r |= 1ull << c_pat.clabel[q];
  return r;
}

// stable compaction of a lane's four slots: out[0 .. n) = the kept values in order; returns n
__device__ __forceinline__ uint32_t compact4(const bool (&keep)[4], const uint32_t (&v)[4], uint32_t (&out)[4]) {
  const uint32_t k0 = keep[0], k1 = keep[1], k2 = keep[2], k3 = keep[3];
  const uint32_t p2 = k0 + k1;
  out[0] = k0 ? v[0] : k1 ? v[1] : k2 ? v[2] : v[3];
  out[1] = (k1 && k0) ? v[1] : (k2 && p2 == 1u) ? v[2] : v[3];
  out[2] = (k2 && p2 == 2u) ? v[2] : v[3];
  out[3] = v[3];
  return p2 + k2 + k3;
}

// exclusive prefix and total over the warp of a per-lane count in [0, 4]
__device__ __forceinline__ void warp_prefix4(uint32_t n, uint32_t lt, uint32_t& below, uint32_t& total) {
  const uint32_t b0 = __ballot_sync(0xffffffffu, n & 1u), b1 = __ballot_sync(0xffffffffu, n & 2u),
                 b2 = __ballot_sync(0xffffffffu, n & 4u);
  below = __popc(b0 & lt) + 2u * __popc(b1 & lt) + 4u * __popc(b2 & lt);
  total = __popc(b0) + 2u * __popc(b1) + 4u * __popc(b2);
}

template <bool FIRST, int STREAM, bool XLATE, bool HEARD>
__global__ void __launch_bounds__(kBlock, 4) k_lcc_scan(LccArgs a, uint4* __restrict__ list,
                                                      const uint32_t* __restrict__ n_ptr, int xlate_only) {
  static_assert(!(FIRST && XLATE) && (FIRST || !STREAM) && (HEARD || FIRST), "unsupported combination");
  __shared__ uint16_t s_lm[17];
  __shared__ uint16_t s_lml[64];  // label value -> labelmask
  // per warp, the 32 rows of the batch in flight: {row start (sectors), |E_v|, first chunk, NB(T_v)}, valid-label
  // set, slots kept so far, masks heard so far, and the lanes that own the long rows in order
  __shared__ uint4 s_rowp[kBlock / 32][32];
  __shared__ unsigned long long s_vl[(FIRST && STREAM) ? kBlock / 32 : 1][32];
  __shared__ uint32_t s_rout[kBlock / 32][32];
  __shared__ uint32_t s_drow[FIRST ? kBlock / 32 : 1][32];  // FIRST: row start (sectors) in the dense working adjacency
  __shared__ uint32_t s_hrd[HEARD ? kBlock / 32 : 1][32];
  __shared__ uint8_t s_nz[kBlock / 32][32];
  __shared__ uint16_t s_ts[XLATE ? 256 : 1];  // typed table: (label, T_state number) -> mask
  if (threadIdx.x < 17) s_lm[threadIdx.x] = c_pat.LMc[threadIdx.x];
  if (threadIdx.x < 64) s_lml[threadIdx.x] = c_pat.LMc[c_pat.cls_of_label[threadIdx.x]];
  if (XLATE) s_ts[threadIdx.x] = c_pat.tsub[c_pat.cls_of_label[threadIdx.x >> 2]][threadIdx.x & 3];
  __syncthreads();
  const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const uint32_t lt = lanemask_lt();
  const uint32_t n = *n_ptr;
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
  const uint32_t* __restrict__ src = FIRST ? a.col0 : a.colw;
  // the rows the renaming scan reads still carry the packed labels of the first scan's slots
  const uint32_t xmask = a.col_shift ? (1u << a.col_shift) - 1u : PM_IDMASK;
  unsigned long long scanned = 0, verts = 0;
  for (uint32_t base = warp * 32; base < n; base += nwarps * 32) {
    const uint32_t idx = base + lane;
    const bool has = idx < n;
    uint4 e = make_uint4(0, 0, 0, 0);
    uint32_t Tv = 0;
    bool live = false;
    if (has) {
      e = list[idx];
      live = e.y != PM_TOMB;
      if (live) Tv = a.S[e.x];
    }
    uint32_t d = Tv ? e.z : 0u;  // Tv == 0: deactivated by NLCC since the last commit (beta.cpp:990-992)
    // FIRST: the row is read where the graph store keeps it (e.y) and its kept neighbours are written to the
    // vertex's row of the dense working adjacency, which every later kernel walks (the entry is re-pointed below)
    uint32_t drow = e.y;
    if (FIRST && has && live) drow = a.rowc[e.x - a.base];
    const uint32_t NBv = nb_of(Tv);
    unsigned long long VL = 0;   // FIRST && STREAM: the labels a valid neighbour can carry
    if (FIRST && STREAM) VL = valid_labels(NBv);
    uint32_t heard = 0, out = 0;

    // The 4-slot chunks of the batch's 32 rows are laid end to end and dealt to the lanes, 32 chunks per pass, so
    // a pass is full whatever the row lengths are (a thread per short row wastes sectors, a warp per row of
    // 32..255 slots leaves the pass half empty).  Row parameters travel through shared memory; the store offset
    // inside a row is a segmented prefix over the lanes plus the row's running count.
    const uint32_t nch = (d + 3u) >> 2;
    uint32_t cum = nch;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, cum, o);
      if (lane >= (uint32_t)o) cum += t;
    }
    const uint32_t C = __shfl_sync(0xffffffffu, cum, 31);  // chunks of this batch
    if (C) {
      const uint32_t first = cum - nch;                    // my row's first chunk
      const uint32_t longrows = __ballot_sync(0xffffffffu, nch != 0u);
      __syncwarp();
      s_rowp[wid][lane] = make_uint4(e.y, d, first, NBv);
      if (FIRST && STREAM) s_vl[wid][lane] = VL;
      s_rout[wid][lane] = 0u;
      if (FIRST) s_drow[wid][lane] = drow;
      if (HEARD) s_hrd[wid][lane] = 0u;
      if (nch) s_nz[wid][__popc(longrows & lt)] = (uint8_t)lane;
      __syncwarp();
      const uint32_t le = lt | (1u << lane);
      // one pass ahead: which row a lane's chunk belongs to, and the chunk itself
      uint32_t rN = 0, hN = 0, jN = 0, dN = 0, lN = 0;
      uint64_t rowN = 0;
      uint4 qN = make_uint4(PM_SENTINEL, PM_SENTINEL, PM_SENTINEL, PM_SENTINEL);
      auto fetch = [&](uint32_t g0) {
        // rows whose first chunk lies in this pass (bit = lane that gets it), long rows that started before it
        const uint32_t hb = (nch && first >= g0 && first < g0 + 32u) ? 1u << (first - g0) : 0u;
        hN = __reduce_or_sync(0xffffffffu, hb);
        const uint32_t before = __popc(__ballot_sync(0xffffffffu, nch && first < g0));
        const uint32_t k = before + __popc(hN & le);  // my chunk's row is the k-th long row of the batch (k >= 1)
        rN = s_nz[wid][k - 1u];
        const uint4 rp = s_rowp[wid][rN];
        const uint32_t g = g0 + lane;
        dN = g < C ? rp.y : 0u;
        jN = (g - rp.z) * 4u;
        rowN = (uint64_t)rp.x * 8;
        qN = make_uint4(PM_SENTINEL, PM_SENTINEL, PM_SENTINEL, PM_SENTINEL);
        lN = 0u;
        if (jN < dN) {
          qN = *reinterpret_cast<const uint4*>(src + rowN + jN);
          if (STREAM == 1) lN = *reinterpret_cast<const uint32_t*>(a.lab0 + rowN + jN);
        }
      };
      fetch(0u);
      for (uint32_t g0 = 0; g0 < C; g0 += 32u) {
        const uint32_t r = rN, H = hN, j0 = jN, rd = dN, l4 = lN;
        const uint64_t rrow = rowN;
        const uint4 q = qN;
        if (g0 + 32u < C) fetch(g0 + 32u);  // the next pass is on its way while this one is worked on
        uint32_t rNB = 0;
        unsigned long long rVL = 0;
        if (FIRST && STREAM) rVL = s_vl[wid][r];
        else rNB = s_rowp[wid][r].w;
        const uint32_t u[4] = {q.x, q.y, q.z, q.w};
        uint32_t wr[4];  // what is stored back: the id as it stands, or (XLATE) the neighbour's compact id
        bool keep[4];
        uint32_t hv = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const bool act = j0 + k < rd;
          wr[k] = STREAM == 2 ? u[k] : u[k] & PM_IDMASK;  // packed: the label stays in the slot for the renaming scan
          uint32_t m = 0;
          bool valid = false;
          if (FIRST && STREAM) {
            const uint32_t lab = STREAM == 2 ? (u[k] >> a.col_shift) & 63u : (l4 >> (8 * k)) & 63u;
            valid = act && ((rVL >> lab) & 1ull);
            if (HEARD) m = s_lml[lab];
          } else if (act && XLATE && a.typed) {
            // one gather: compact id and T_state number of the neighbour together; its label rides in the slot
            const uint32_t id = u[k] & xmask;
            const uint32_t t = typed_lookup(a.fwx[id >> 4], id, wr[k]);
            m = s_ts[((u[k] >> a.col_shift) & 63u) * 4u + t];
            if (xlate_only) m = t ? 0xFFFFu : 0u;
            if (!t) wr[k] = PM_SENTINEL;
            valid = xlate_only ? m != 0u : (m & rNB) != 0u;
          } else if (act) {
            if (XLATE) wr[k] = cid_of_slot(a.fwx, u[k] & xmask);
            if (FIRST) m = s_lm[a.cls[wr[k]]];
            else if (!XLATE || wr[k] != PM_SENTINEL) m = xlate_only ? 0xFFFFu : (uint32_t)a.S[wr[k]];
            valid = xlate_only ? m != 0u : (m & rNB) != 0u;
          }
          const bool pre = !FIRST && !XLATE && act && (u[k] >> 31);
          keep[k] = valid || pre;
          if (HEARD && valid) hv |= m;
        }
        uint32_t outv[4], below, total;
        const uint32_t nk = compact4(keep, wr, outv);
        warp_prefix4(nk, lt, below, total);
        // every load of this pass is complete (the counts depend on them); the prefetched pass lies strictly
        // behind every slot written now: a row's writes land at or before slots of it already read
        const uint32_t hl = 31u - __clz((H | 1u) & le);  // first lane of my row's segment in this pass
        const uint32_t off = below - __shfl_sync(0xffffffffu, below, hl);
        const uint32_t rout = s_rout[wid][r];
        const uint64_t wrow = FIRST ? (uint64_t)s_drow[wid][r] * 8 : rrow;
        uint32_t* __restrict__ p = a.colw + wrow + rout + off;
        if (nk > 0u) p[0] = outv[0];
        if (nk > 1u) p[1] = outv[1];
        if (nk > 2u) p[2] = outv[2];
        if (nk > 3u) p[3] = outv[3];
        const bool last = rd != 0u && (lane == 31u || ((H >> (lane + 1u)) & 1u) || g0 + lane + 1u == C);
        if (HEARD) {  // OR over the lanes of my segment
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, hv, o);
            if (lane >= (uint32_t)o + hl) hv |= t;
          }
        }
        __syncwarp();  // every lane has read its row's running count
        if (last) {
          s_rout[wid][r] = rout + off + nk;
          if (HEARD) s_hrd[wid][r] |= hv;
        }
        __syncwarp();
      }
      if (nch) {
        out = s_rout[wid][lane];
        if (HEARD) heard = s_hrd[wid][lane];
      }
    }

    if (has && live) {
      uint32_t ts;
      if (!HEARD) {
        ts = Tv ? e.w : 0u;  // decided by the signature filter (k_init_assign): never empty
      } else {
        const uint32_t T0 = FIRST ? Tv : e.w;
        ts = Tv ? cover_of(T0, heard) : 0u;
        if (FIRST && heard == 0u) ts = 0u;   // never entered the map (ee.hpp:841-866; an approximate pattern may require nothing)
        if (xlate_only) ts = Tv ? e.w : 0u;  // renaming only: T_state stays as it is
        // a vertex leaves the vertex_state_map (ee.hpp:941-946, :968-970).  In the first
        // superstep only vertices that heard a valid neighbour ever entered the map (:841-852).
        if (ts == 0 && (FIRST ? heard != 0u : Tv != 0u)) a.cnt->nf = 1u;
      }
      scanned += d;
      verts += Tv != 0u;
      if (FIRST) e.y = drow;
      e.z = out;
      e.w = ts;
      list[idx] = e;
    }
  }
  // one atomic per warp
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

// ---------------------------------------------------------------------------
// The first-superstep scan of the hot configuration: packed labels (col_shift != 0) behind the signature filter
// (every entry already carries the T_state the superstep ends with, k_init_assign).  Same packing of rows into
// lanes as k_lcc_scan, but a lane takes a whole 32-byte SECTOR (8 slots) per pass — rows of the graph store are
// sector aligned and padded with PM_SENTINEL, whose label field is all ones and never valid, so there is no
// per-slot bounds test — and the per-pass bookkeeping is spread over 256 slots instead of 128: the generic kernel
// is instruction-issue bound at 1.4 warp instructions per slot, this one needs about half.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void warp_prefix8(uint32_t n, uint32_t lt, uint32_t& below, uint32_t& total) {
  const uint32_t b0 = __ballot_sync(0xffffffffu, n & 1u), b1 = __ballot_sync(0xffffffffu, n & 2u),
                 b2 = __ballot_sync(0xffffffffu, n & 4u), b3 = __ballot_sync(0xffffffffu, n & 8u);
  below = __popc(b0 & lt) + 2u * __popc(b1 & lt) + 4u * __popc(b2 & lt) + 8u * __popc(b3 & lt);
  total = __popc(b0) + 2u * __popc(b1) + 4u * __popc(b2) + 8u * __popc(b3);
}

__global__ void __launch_bounds__(kBlock, 3) k_lcc_first_packed(LccArgs a, uint4* __restrict__ list,
                                                                const uint32_t* __restrict__ n_ptr) {
  // per warp, the 32 rows of the batch in flight: {source row (sectors), sectors, first sector of the batch, dense row}
  __shared__ uint4 s_rowp[kBlock / 32][32];
  __shared__ unsigned long long s_vl[kBlock / 32][32];
  __shared__ uint32_t s_rout[kBlock / 32][32];
  __shared__ uint8_t s_nz[kBlock / 32][32];
  // per label class: the labelmask (T_arr of the first superstep, ee.hpp:541-546) and the labels a valid neighbour
  // can carry (PatConst::rl) — one byte gather per entry instead of two 16-step loops over the template
  __shared__ uint16_t s_lm[17];
  __shared__ unsigned long long s_rl[17];
  if (threadIdx.x < 17) { s_lm[threadIdx.x] = c_pat.LMc[threadIdx.x]; s_rl[threadIdx.x] = c_pat.rl[threadIdx.x]; }
  __syncthreads();
  const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const uint32_t lt = lanemask_lt();
  const uint32_t le = lt | (1u << lane);
  const uint32_t n = *n_ptr;
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
  const uint32_t shift = a.col_shift;
  const unsigned long long not_sentinel = ~(1ull << (0xFFFFFFFFu >> shift));
  unsigned long long scanned = 0, verts = 0;
  for (uint32_t base = warp * 32; base < n; base += nwarps * 32) {
    const uint32_t idx = base + lane;
    const bool has = idx < n;
    uint4 e = make_uint4(0, 0, 0, 0);
    uint32_t cl = PM_NOCLASS;
    bool live = false;
    if (has) {
      e = list[idx];
      live = e.y != PM_TOMB;
      if (live) cl = a.clsc[e.x];
    }
    const uint32_t Tv = s_lm[cl];
    const uint32_t d = Tv ? e.z : 0u;
    uint32_t drow = e.y, out = 0;
    if (has && live) drow = a.rowc[e.x - a.base];
    const unsigned long long VL = s_rl[cl] & not_sentinel;
    const uint32_t nch = (d + 7u) >> 3;
    uint32_t cum = nch;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, cum, o);
      if (lane >= (uint32_t)o) cum += t;
    }
    const uint32_t C = __shfl_sync(0xffffffffu, cum, 31);  // sectors of this batch
    if (C) {
      const uint32_t first = cum - nch;
      const uint32_t longrows = __ballot_sync(0xffffffffu, nch != 0u);
      __syncwarp();
      s_rowp[wid][lane] = make_uint4(e.y, nch, first, drow);
      s_vl[wid][lane] = VL;
      s_rout[wid][lane] = 0u;
      if (nch) s_nz[wid][__popc(longrows & lt)] = (uint8_t)lane;
      __syncwarp();
      // one pass ahead: the row a lane's sector belongs to and the sector itself
      uint32_t rN = 0, hN = 0;
      uint4 qa = make_uint4(PM_SENTINEL, PM_SENTINEL, PM_SENTINEL, PM_SENTINEL), qb = qa;
      auto fetch = [&](uint32_t g0) {
        const uint32_t hb = (nch && first >= g0 && first < g0 + 32u) ? 1u << (first - g0) : 0u;
        hN = __reduce_or_sync(0xffffffffu, hb);
        const uint32_t before = __popc(__ballot_sync(0xffffffffu, nch && first < g0));
        rN = s_nz[wid][before + __popc(hN & le) - 1u];
        const uint4 rp = s_rowp[wid][rN];
        const uint32_t g = g0 + lane;
        qa = make_uint4(PM_SENTINEL, PM_SENTINEL, PM_SENTINEL, PM_SENTINEL);
        qb = qa;
        if (g < C) {
          const uint4* __restrict__ p = reinterpret_cast<const uint4*>(a.col0 + ((uint64_t)rp.x + (g - rp.z)) * 8);
          qa = p[0];
          qb = p[1];
        }
      };
      fetch(0u);
      for (uint32_t g0 = 0; g0 < C; g0 += 32u) {
        const uint32_t r = rN, H = hN;
        const uint4 q0 = qa, q1 = qb;
        if (g0 + 32u < C) fetch(g0 + 32u);  // the next pass is on its way while this one is worked on
        const unsigned long long rVL = s_vl[wid][r];
        const uint32_t u[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
        uint32_t keep[8], nk = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          keep[k] = (uint32_t)(rVL >> (u[k] >> shift)) & 1u;  // a sentinel's label field is never valid
          nk += keep[k];
        }
        uint32_t below, total;
        warp_prefix8(nk, lt, below, total);
        const uint32_t hl = 31u - __clz((H | 1u) & le);  // first lane of my row's segment in this pass
        const uint32_t off = below - __shfl_sync(0xffffffffu, below, hl);
        const uint32_t rout = s_rout[wid][r];
        uint32_t* __restrict__ p = a.colw + (uint64_t)s_rowp[wid][r].w * 8 + rout + off;
#pragma unroll
        for (int k = 0; k < 8; ++k)
          if (keep[k]) *p++ = u[k];  // label and id: the renaming scan strips the label
        const bool in_batch = g0 + lane < C;
        const bool last = in_batch && (lane == 31u || ((H >> (lane + 1u)) & 1u) || g0 + lane + 1u == C);
        __syncwarp();  // every lane has read its row's running count
        if (last) s_rout[wid][r] = rout + off + nk;
        __syncwarp();
      }
      if (nch) out = s_rout[wid][lane];
    }
    if (has && live) {
      scanned += d;
      verts += Tv != 0u;
      e.y = drow;
      e.z = out;
      e.w = Tv ? e.w : 0u;  // decided by the signature filter (k_init_assign): never empty
      list[idx] = e;
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

// The same scan carrying the SECOND superstep along (one rank, no big rows; typed slot table, see k_init_assign).
// After the first superstep T_arr(u) of every survivor u equals the T_state the signature filter gave it, so the second
// superstep's test of a kept neighbour — T_arr(u) & NB(T_arr(v)) != 0 (ee.hpp:673-722) — needs one 8-byte gather per
// LABEL-VALID slot: compact id and mask number of u together.  The row written is E_v after the second superstep, in
// compact ids; heard(v) and T_state(v) of the second superstep come out of the same pass.  This replaces the first
// scan, its commit and the renaming scan (two dependent gathers per slot) of the unfused path.  Row counts: the first
// superstep's (vertices = survivors, edges = label-valid slots) are accumulated here, the second's by the commit.
__global__ void __launch_bounds__(kBlock, 3) k_lcc_first_fused(LccArgs a, uint4* __restrict__ list,
                                                               const uint32_t* __restrict__ n_ptr, RowStat* __restrict__ row1) {
  // per warp, the 32 rows of the batch in flight: {source row (sectors), sectors, first sector of the batch, dense row}
  __shared__ uint4 s_rowp[kBlock / 32][32];
  __shared__ unsigned long long s_vl[kBlock / 32][32];
  __shared__ uint32_t s_rout[kBlock / 32][32];
  __shared__ uint8_t s_nz[kBlock / 32][32];
  __shared__ uint32_t s_nb1[kBlock / 32][32];  // NB(T_arr(v)) of the second superstep
  __shared__ uint32_t s_hrd[kBlock / 32][32];  // masks heard in the second superstep
  __shared__ uint16_t s_ts[256], s_lm[17];  // (label, T_state number) -> mask; class -> labelmask
  s_ts[threadIdx.x] = c_pat.tsub[c_pat.cls_of_label[threadIdx.x >> 2]][threadIdx.x & 3];
  if (threadIdx.x < 17) s_lm[threadIdx.x] = c_pat.LMc[threadIdx.x];
  __syncthreads();
  const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const uint32_t lt = lanemask_lt();
  const uint32_t le = lt | (1u << lane);
  const uint32_t n = *n_ptr;
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
  const uint32_t shift = a.col_shift, idmask = (1u << shift) - 1u;
  const unsigned long long not_sentinel = ~(1ull << (0xFFFFFFFFu >> shift));
  unsigned long long scanned = 0, verts = 0, kept0 = 0;
  for (uint32_t base = warp * 32; base < n; base += nwarps * 32) {
    const uint32_t idx = base + lane;
    const bool has = idx < n;
    uint4 e = make_uint4(0, 0, 0, 0);
    uint32_t Tv = 0;
    bool live = false;
    if (has) {
      e = list[idx];
      live = e.y != PM_TOMB;
      if (live) Tv = s_lm[a.clsc[e.x]];  // T_arr of the first superstep: the labelmask (ee.hpp:541-546)
    }
    const uint32_t d = Tv ? e.z : 0u;
    uint32_t drow = e.y, out = 0, heard = 0;
    if (has && live) drow = a.rowc[e.x - a.base];
    const unsigned long long VL = valid_labels(nb_of(Tv)) & not_sentinel;
    const uint32_t nch = (d + 7u) >> 3;
    uint32_t cum = nch;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, cum, o);
      if (lane >= (uint32_t)o) cum += t;
    }
    const uint32_t C = __shfl_sync(0xffffffffu, cum, 31);  // sectors of this batch
    if (C) {
      const uint32_t first = cum - nch;
      const uint32_t longrows = __ballot_sync(0xffffffffu, nch != 0u);
      __syncwarp();
      s_rowp[wid][lane] = make_uint4(e.y, nch, first, drow);
      s_vl[wid][lane] = VL;
      s_rout[wid][lane] = 0u;
      s_nb1[wid][lane] = nb_of(e.w);  // e.w: T_state after the first superstep = T_arr of the second
      s_hrd[wid][lane] = 0u;
      if (nch) s_nz[wid][__popc(longrows & lt)] = (uint8_t)lane;
      __syncwarp();
      // one pass ahead: the row a lane's sector belongs to and the sector itself
      uint32_t rN = 0, hN = 0;
      uint4 qa = make_uint4(PM_SENTINEL, PM_SENTINEL, PM_SENTINEL, PM_SENTINEL), qb = qa;
      auto fetch = [&](uint32_t g0) {
        const uint32_t hb = (nch && first >= g0 && first < g0 + 32u) ? 1u << (first - g0) : 0u;
        hN = __reduce_or_sync(0xffffffffu, hb);
        const uint32_t before = __popc(__ballot_sync(0xffffffffu, nch && first < g0));
        rN = s_nz[wid][before + __popc(hN & le) - 1u];
        const uint4 rp = s_rowp[wid][rN];
        const uint32_t g = g0 + lane;
        qa = make_uint4(PM_SENTINEL, PM_SENTINEL, PM_SENTINEL, PM_SENTINEL);
        qb = qa;
        if (g < C) {
          const uint4* __restrict__ p = reinterpret_cast<const uint4*>(a.col0 + ((uint64_t)rp.x + (g - rp.z)) * 8);
          qa = p[0];
          qb = p[1];
        }
      };
      fetch(0u);
      for (uint32_t g0 = 0; g0 < C; g0 += 32u) {
        const uint32_t r = rN, H = hN;
        const uint4 q0 = qa, q1 = qb;
        if (g0 + 32u < C) fetch(g0 + 32u);  // the next pass is on its way while this one is worked on
        const unsigned long long rVL = s_vl[wid][r];
        const uint32_t rNB = s_nb1[wid][r];
        const uint32_t u[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
        uint32_t cidv[8], nk = 0, hv = 0, n0 = 0;
        // first superstep: label test (a sentinel's label field is never valid); second: the kept neighbour's entry
        // of the typed table — every gather of the pass is issued before the first one is used
        uint2 w[8];
        uint32_t k0 = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const uint32_t lab = u[k] >> shift;
          if ((uint32_t)(rVL >> lab) & 1u) {
            k0 |= 1u << k;
            w[k] = a.fwx[(u[k] & idmask) >> 4];
          }
        }
        n0 = __popc(k0);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          cidv[k] = PM_SENTINEL;
          if ((k0 >> k) & 1u) {
            uint32_t cid;
            const uint32_t t = typed_lookup(w[k], u[k] & idmask, cid);
            const uint32_t m = s_ts[(u[k] >> shift) * 4u + t];  // 0: u did not survive the first superstep
            if (m & rNB) {
              hv |= m;
              cidv[k] = cid;
              ++nk;
            }
          }
        }
        kept0 += n0;
        uint32_t below, total;
        warp_prefix8(nk, lt, below, total);
        const uint32_t hl = 31u - __clz((H | 1u) & le);  // first lane of my row's segment in this pass
        const uint32_t off = below - __shfl_sync(0xffffffffu, below, hl);
        const uint32_t rout = s_rout[wid][r];
        uint32_t* __restrict__ p = a.colw + (uint64_t)s_rowp[wid][r].w * 8 + rout + off;
#pragma unroll
        for (int k = 0; k < 8; ++k)
          if (cidv[k] != PM_SENTINEL) *p++ = cidv[k];
        const bool in_batch = g0 + lane < C;
        const bool last = in_batch && (lane == 31u || ((H >> (lane + 1u)) & 1u) || g0 + lane + 1u == C);
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {  // OR over the lanes of my segment
          const uint32_t t = __shfl_up_sync(0xffffffffu, hv, o);
          if (lane >= (uint32_t)o + hl) hv |= t;
        }
        __syncwarp();  // every lane has read its row's running count
        if (last) {
          s_rout[wid][r] = rout + off + nk;
          s_hrd[wid][r] |= hv;
        }
        __syncwarp();
      }
      if (nch) { out = s_rout[wid][lane]; heard = s_hrd[wid][lane]; }
    }
    if (has && live) {
      scanned += d;
      verts += Tv != 0u;
      // second superstep: T_state &= { p : N(p) heard } (ee.hpp:901-939); an empty one leaves the map (:941-946)
      const uint32_t ts = Tv ? cover_of(e.w, heard) : 0u;
      if (Tv && ts == 0u) a.cnt->nf = 1u;
      e.y = drow;
      e.z = out;
      e.w = ts;
      list[idx] = e;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    scanned += __shfl_xor_sync(0xffffffffu, scanned, o);
    verts += __shfl_xor_sync(0xffffffffu, verts, o);
    kept0 += __shfl_xor_sync(0xffffffffu, kept0, o);
  }
  if (lane == 0 && verts) {
    atomicAdd(&a.row->scanned[0], scanned);
    atomicAdd(&a.row->verts[0], verts);
    // the first superstep's row (ee.hpp:1112-1138): every survivor is in the map, E_v = the label-valid neighbours
    atomicAdd(&a.row->nv, verts);
    atomicAdd(&a.row->ne, kept0);
    // the second superstep walked those label-valid slots
    atomicAdd(&row1->scanned[0], kept0);
    atomicAdd(&row1->verts[0], verts);
  }
}

// ---------------------------------------------------------------------------
// The renaming scan with the typed table (one rank; see PatConst::tsub): second superstep of the first LCC call over
// the rows the first scan wrote (packed slots).  Same sector-per-lane packing as k_lcc_first_packed; one 8-byte gather
// per slot gives the neighbour's compact id and T_state number, its label rides in the slot: mask = tsub[label][number].
// The row is compacted in place (writes land at or before slots already read; the prefetched pass lies behind them).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock, 3) k_lcc_xlate8(LccArgs a, uint4* __restrict__ list,
                                                          const uint32_t* __restrict__ n_ptr) {
  __shared__ uint4 s_rowp[kBlock / 32][32];  // {row start (sectors), |E_v|, first sector of the batch, NB(T_arr(v))}
  __shared__ uint32_t s_rout[kBlock / 32][32];
  __shared__ uint32_t s_hrd[kBlock / 32][32];
  __shared__ uint8_t s_nz[kBlock / 32][32];
  __shared__ uint16_t s_ts[256];  // (label, T_state number) -> mask
  s_ts[threadIdx.x] = c_pat.tsub[c_pat.cls_of_label[threadIdx.x >> 2]][threadIdx.x & 3];
  __syncthreads();
  const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const uint32_t lt = lanemask_lt();
  const uint32_t le = lt | (1u << lane);
  const uint32_t n = *n_ptr;
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
  const uint32_t shift = a.col_shift, idmask = (1u << shift) - 1u;
  unsigned long long scanned = 0, verts = 0;
  for (uint32_t base = warp * 32; base < n; base += nwarps * 32) {
    const uint32_t idx = base + lane;
    const bool has = idx < n;
    uint4 e = make_uint4(0, 0, 0, 0);
    uint32_t Tv = 0;
    bool live = false;
    if (has) {
      e = list[idx];
      live = e.y != PM_TOMB;
      if (live) Tv = a.S[e.x];
    }
    const uint32_t d = Tv ? e.z : 0u;
    uint32_t out = 0, heard = 0;
    const uint32_t nch = (d + 7u) >> 3;
    uint32_t cum = nch;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, cum, o);
      if (lane >= (uint32_t)o) cum += t;
    }
    const uint32_t C = __shfl_sync(0xffffffffu, cum, 31);  // sectors of this batch
    if (C) {
      const uint32_t first = cum - nch;
      const uint32_t longrows = __ballot_sync(0xffffffffu, nch != 0u);
      __syncwarp();
      s_rowp[wid][lane] = make_uint4(e.y, d, first, nb_of(Tv));
      s_rout[wid][lane] = 0u;
      s_hrd[wid][lane] = 0u;
      if (nch) s_nz[wid][__popc(longrows & lt)] = (uint8_t)lane;
      __syncwarp();
      uint32_t rN = 0, hN = 0, jN = 0;
      uint4 qa = make_uint4(0, 0, 0, 0), qb = qa;
      auto fetch = [&](uint32_t g0) {
        const uint32_t hb = (nch && first >= g0 && first < g0 + 32u) ? 1u << (first - g0) : 0u;
        hN = __reduce_or_sync(0xffffffffu, hb);
        const uint32_t before = __popc(__ballot_sync(0xffffffffu, nch && first < g0));
        rN = s_nz[wid][before + __popc(hN & le) - 1u];
        const uint4 rp = s_rowp[wid][rN];
        const uint32_t g = g0 + lane;
        jN = (g - rp.z) * 8u;  // first slot of my sector inside its row
        if (g < C) {
          const uint4* __restrict__ p = reinterpret_cast<const uint4*>(a.colw + ((uint64_t)rp.x + (g - rp.z)) * 8);
          qa = p[0];
          qb = p[1];
        }
      };
      fetch(0u);
      for (uint32_t g0 = 0; g0 < C; g0 += 32u) {
        const uint32_t r = rN, H = hN, j0 = jN;
        const uint4 q0 = qa, q1 = qb;
        const bool in_batch = g0 + lane < C;
        if (g0 + 32u < C) fetch(g0 + 32u);  // the next pass is on its way while this one is worked on
        const uint4 rp = s_rowp[wid][r];
        const uint32_t rNB = rp.w, rd = in_batch ? rp.y : 0u;
        const uint32_t u[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
        uint2 w[8];
#pragma unroll
        for (int k = 0; k < 8; ++k)
          if (j0 + k < rd) w[k] = a.fwx[(u[k] & idmask) >> 4];  // every gather of the pass before the first use
        uint32_t cidv[8], nk = 0, hv = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          cidv[k] = PM_SENTINEL;
          if (j0 + k < rd) {
            uint32_t cid;
            const uint32_t t = typed_lookup(w[k], u[k] & idmask, cid);
            const uint32_t m = s_ts[((u[k] >> shift) & 63u) * 4u + t];  // 0: the neighbour did not survive the first superstep
            if (m & rNB) {
              hv |= m;
              cidv[k] = cid;
              ++nk;
            }
          }
        }
        uint32_t below, total;
        warp_prefix8(nk, lt, below, total);
        const uint32_t hl = 31u - __clz((H | 1u) & le);  // first lane of my row's segment in this pass
        const uint32_t off = below - __shfl_sync(0xffffffffu, below, hl);
        const uint32_t rout = s_rout[wid][r];
        uint32_t* __restrict__ p = a.colw + (uint64_t)rp.x * 8 + rout + off;
#pragma unroll
        for (int k = 0; k < 8; ++k)
          if (cidv[k] != PM_SENTINEL) *p++ = cidv[k];
        const bool last = in_batch && (lane == 31u || ((H >> (lane + 1u)) & 1u) || g0 + lane + 1u == C);
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {  // OR over the lanes of my segment
          const uint32_t t = __shfl_up_sync(0xffffffffu, hv, o);
          if (lane >= (uint32_t)o + hl) hv |= t;
        }
        __syncwarp();  // every lane has read its row's running count
        if (last) {
          s_rout[wid][r] = rout + off + nk;
          s_hrd[wid][r] |= hv;
        }
        __syncwarp();
      }
      if (nch) { out = s_rout[wid][lane]; heard = s_hrd[wid][lane]; }
    }
    if (has && live) {
      const uint32_t ts = Tv ? cover_of(e.w, heard) : 0u;   // ee.hpp:901-939
      if (Tv && ts == 0u) a.cnt->nf = 1u;                    // left the map (:941-946, :968-970)
      scanned += d;
      verts += Tv != 0u;
      e.z = out;
      e.w = ts;
      list[idx] = e;
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

// one CTA per high-degree vertex ("delegates across warps and CTAs")
template <bool FIRST, int STREAM, bool XLATE, bool HEARD>
__global__ void __launch_bounds__(1024) k_lcc_scan_big(LccArgs a, uint4* __restrict__ list,
                                                        const uint32_t* __restrict__ n_ptr, int xlate_only) {
  __shared__ uint16_t s_lm[17];
  __shared__ uint16_t s_lml[64];
  if (threadIdx.x < 64) s_lml[threadIdx.x] = c_pat.LMc[c_pat.cls_of_label[threadIdx.x]];
  __shared__ uint32_t s_wcnt[32];
  __shared__ uint32_t s_heard[32];
  __shared__ uint16_t s_ts[XLATE ? 256 : 1];  // typed table: (label, T_state number) -> mask
  if (XLATE && threadIdx.x < 256) s_ts[threadIdx.x] = c_pat.tsub[c_pat.cls_of_label[threadIdx.x >> 2]][threadIdx.x & 3];
  const uint32_t xmask = a.col_shift ? (1u << a.col_shift) - 1u : PM_IDMASK;
  if (threadIdx.x < 17) s_lm[threadIdx.x] = c_pat.LMc[threadIdx.x];
  const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const uint32_t lt = lanemask_lt();
  const uint32_t n = *n_ptr;
  for (uint32_t idx = blockIdx.x; idx < n; idx += gridDim.x) {
    __syncthreads();
    const uint4 e = list[idx];
    const uint32_t Tv = a.S[e.x];
    const uint32_t d = Tv ? e.z : 0u;
    const uint32_t NBv = nb_of(Tv);
    unsigned long long VL = 0;
    if (FIRST && STREAM) VL = valid_labels(NBv);
    const uint64_t row = (uint64_t)e.y * 8;
    const uint32_t drow = FIRST ? a.rowc[e.x - a.base] : e.y;  // FIRST: kept neighbours go to the dense working adjacency
    const uint64_t wrow = (uint64_t)drow * 8;
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
        if (STREAM == 1) l4 = *reinterpret_cast<const uint32_t*>(a.lab0 + row + j0);
      }
      const uint32_t u[4] = {q.x, q.y, q.z, q.w};
      uint32_t wr[4];  // what is stored back: the id as it stands, or (XLATE) the neighbour's compact id
      bool keep[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const bool act = j0 + k < d;
        wr[k] = STREAM == 2 ? u[k] : u[k] & PM_IDMASK;  // packed: the label stays in the slot for the renaming scan
        uint32_t m = 0;
        bool valid = false;
        if (FIRST && STREAM) {
          const uint32_t lab = STREAM == 2 ? (u[k] >> a.col_shift) & 63u : (l4 >> (8 * k)) & 63u;
          valid = act && ((VL >> lab) & 1ull);
          if (HEARD) m = s_lml[lab];
        } else if (act && XLATE && a.typed) {
          const uint32_t id = u[k] & xmask;
          const uint32_t t = typed_lookup(a.fwx[id >> 4], id, wr[k]);
          m = s_ts[((u[k] >> a.col_shift) & 63u) * 4u + t];
          if (xlate_only) m = t ? 0xFFFFu : 0u;
          if (!t) wr[k] = PM_SENTINEL;
          valid = xlate_only ? m != 0u : (m & NBv) != 0u;
        } else if (act) {
          if (XLATE) wr[k] = cid_of_slot(a.fwx, u[k] & xmask);
          if (FIRST) m = s_lm[a.cls[wr[k]]];
          else if (!XLATE || wr[k] != PM_SENTINEL) m = xlate_only ? 0xFFFFu : (uint32_t)a.S[wr[k]];
          valid = xlate_only ? m != 0u : (m & NBv) != 0u;
        }
        const bool pre = !FIRST && !XLATE && act && (u[k] >> 31);
        keep[k] = valid || pre;
        if (HEARD && valid) heard |= m;
      }
      uint32_t outv[4], below, wtotal;
      const uint32_t nk = compact4(keep, wr, outv);
      warp_prefix4(nk, lt, below, wtotal);
      if (lane == 0) s_wcnt[wid] = wtotal;
      __syncthreads();  // every warp has read its slots of this pass
      uint32_t wbase = outp, ptotal = 0;
      for (uint32_t w = 0; w < nw; ++w) {
        const uint32_t cw = s_wcnt[w];
        if (w < wid) wbase += cw;
        ptotal += cw;
      }
      uint32_t* __restrict__ p = a.colw + wrow + wbase + below;
      if (nk > 0u) p[0] = outv[0];
      if (nk > 1u) p[1] = outv[1];
      if (nk > 2u) p[2] = outv[2];
      if (nk > 3u) p[3] = outv[3];
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
      uint32_t ts;
      if (!HEARD) {
        ts = Tv ? e.w : 0u;
      } else {
        const uint32_t T0 = FIRST ? Tv : e.w;
        ts = Tv ? cover_of(T0, h) : 0u;
        if (FIRST && h == 0u) ts = 0u;       // never entered the map
        if (xlate_only) ts = Tv ? e.w : 0u;  // renaming only: T_state stays as it is
        if (ts == 0 && (FIRST ? h != 0u : Tv != 0u)) a.cnt->nf = 1u;
      }
      uint4 e2 = e;
      e2.y = drow;
      e2.z = outp;
      e2.w = ts;
      list[idx] = e2;
      if (Tv) {
        atomicAdd(&a.row->scanned[2], (unsigned long long)d);
        atomicAdd(&a.row->verts[2], 1ull);
      }
    }
  }
}

// ---------------------------------------------------------------------------
// commit: publish T_arr (ee.hpp:948) and |E_v|, drop removed vertices (:941-946), bin the
// survivors by their new |E_v| into the next frontier and accumulate the row
// counts the reference writes after every superstep (:1112-1138).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kBlock) k_lcc_commit(LccArgs a, const uint4* __restrict__ l0,
                                                        const uint4* __restrict__ l1, uint4* n0, uint4* n1,
                                                        int cur, int nxt) {
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t c0 = a.cnt->fr_n[cur][0], c1 = a.cnt->fr_n[cur][1];
  const uint32_t total = c0 + c1;
  unsigned long long nv = 0, ne = 0;
  constexpr int IT = 4;
  const bool publish = c_peer.G > 1;  // several ranks: mask changes go to the peers' delta inboxes
  const uint32_t tile = blockDim.x * IT;
  for (uint32_t base = blockIdx.x * tile; base < total; base += gridDim.x * tile) {
    bool alive[IT];
    int bin[IT];
    uint4 val[IT];
    uint32_t was[IT];
    // all loads of the tile first: the stores to S below may alias the loads of the next item as far as the compiler
    // knows, which would chain the items' load latencies one behind the other
#pragma unroll
    for (int k = 0; k < IT; ++k) {
      const uint32_t i = base + k * blockDim.x + threadIdx.x;
      val[k] = make_uint4(0, PM_TOMB, 0, 0);
      was[k] = 0;
      if (i < total) {
        val[k] = i < c0 ? l0[i] : l1[i - c0];
        if (publish && val[k].y != PM_TOMB) was[k] = a.S[val[k].x];  // peers only need the changes
      }
    }
#pragma unroll
    for (int k = 0; k < IT; ++k) {
      alive[k] = false;
      bin[k] = 0;
      bool changed = false;
      uint32_t cslot = 0, cmask = 0;
      const uint4 e = val[k];
      if (e.y != PM_TOMB) {
        const uint32_t ts = e.w, d = e.z;
        cslot = e.x;
        cmask = ts;
        changed = publish && ts != was[k];
        a.S[cslot] = (uint16_t)ts;
        a.adeg[e.x - a.base] = d;
        alive[k] = ts != 0;
        const uint32_t hc = a.hubc ? (uint32_t)a.hubc[e.x - a.base] : 0u;
        if (alive[k] && hc) {  // a hub: its row of the count files belongs to its controller
          atomicAdd(&a.row->hub_nv[hc - 1u], 1ull);
          atomicAdd(&a.row->hub_ne[hc - 1u], (unsigned long long)d);
        } else if (alive[k]) { nv++; ne += d; }
        bin[k] = bin_of(d);
      } else {
        val[k] = make_uint4(0, 0, 0, 0);
      }
      if (publish) publish_mask(changed, cslot, cmask, a.cnt, a.par);
    }
    block_append2<IT>(alive, bin, val, n0, n1, &a.cnt->fr_n[nxt][0]);
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
  // the last block to finish re-arms the counters of the list just consumed: they are the NEXT commit's output
  // counters (every block read them on entry; no other kernel of the stream reads them before that commit)
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(&a.cnt->ticket, 1u) == gridDim.x - 1u) {
      a.cnt->fr_n[cur][0] = a.cnt->fr_n[cur][1] = a.cnt->fr_n[cur][2] = a.cnt->fr_n[cur][3] = 0u;
      a.cnt->ticket = 0u;
    }
  }
}

// counts after an NLCC constraint (beta.cpp:1094-1120): vertices still in the map
// and the sizes of their edge maps.  Frontier lists are left untouched; entries
// deactivated by NLCC are skipped by the next scan and dropped by its commit.
__global__ void __launch_bounds__(kBlock) k_count_alive(LccArgs a, const uint4* __restrict__ l0,
                                                         const uint4* __restrict__ l1, int cur) {
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t c0 = a.cnt->fr_n[cur][0], c1 = a.cnt->fr_n[cur][1];
  const uint32_t total = c0 + c1;
  unsigned long long nv = 0, ne = 0;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const uint4 e = i < c0 ? l0[i] : l1[i - c0];
    if (e.y != PM_TOMB && a.S[e.x]) {
      const uint32_t hc = a.hubc ? (uint32_t)a.hubc[e.x - a.base] : 0u;
      if (hc) {
        atomicAdd(&a.row->hub_nv[hc - 1u], 1ull);
        atomicAdd(&a.row->hub_ne[hc - 1u], (unsigned long long)e.z);
      } else { nv++; ne += e.z; }
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

}  // namespace pm
