// pm_graph.cuh — device-resident graph store.
//
// Replaces delegate_partitioned_graph construction
// (/root/reference/include/havoqgt/impl/delegate_partitioned_graph.ipp:112-165:
//  count_edge_degrees -> partition_low/high_degree) with a radix-sort based CSR
// build that runs entirely on the GPU:
//   degm[v]   multigraph out-degree, duplicates and self loops counted (ipp:437-470);
//             this is what the degree labels are derived from
//   col0      distinct neighbours of v, ascending; the reference keeps parallel
//             edges in its CSR but everything downstream of the first superstep is
//             keyed by neighbour id (vertex_active_edges_map, beta.cpp:293-294), so
//             a distinct-neighbour CSR carries the same information
//   rows are padded to multiples of 8 slots (one 32-byte sector) so that every
//   row starts sector- and uint4-aligned; rowblk[v] is the row start in sectors.
#pragma once

#include <cub/cub.cuh>

#include "pm_common.cuh"
#include "pm_comm.cuh"

namespace pm {

// slot(v) = (v mod G) * nlmax + v / G (see PeerTab); G == 1: slot = v
__device__ __forceinline__ uint32_t d_slot_of(uint32_t v, uint32_t G, uint32_t nlmax) {
  return G == 1 ? v : (v % G) * nlmax + v / G;
}

// keys[i] = slot(src) << 32 | slot(dst).  drop_foreign: slots whose source another rank owns become
// ~0 (sorted to the end and cut off) — the caller passed the whole edge list to every rank.
__global__ void k_slot_keys(const uint32_t* __restrict__ src, const uint32_t* __restrict__ dst, uint64_t n,
                            uint32_t G, uint32_t nlmax, uint32_t rank, int drop_foreign,
                            unsigned long long* __restrict__ keys) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    const uint32_t s = src[i], d = dst[i];
    unsigned long long k = ((unsigned long long)d_slot_of(s, G, nlmax) << 32) | d_slot_of(d, G, nlmax);
    if (drop_foreign && s % G != rank) k = ~0ull;
    keys[i] = k;
  }
}

// multigraph out-degree of the local rows (duplicates and self loops counted, ipp:437-470)
__global__ void k_degm_of_keys(const unsigned long long* __restrict__ keys, uint64_t n, uint32_t base,
                               uint32_t* __restrict__ degm) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) atomicAdd(&degm[(uint32_t)(keys[i] >> 32) - base], 1u);
}

// first index whose key is >= bound[g] (keys ascending): the owner ranges of a sorted key array
__global__ void k_lower_bounds(const unsigned long long* __restrict__ keys, uint64_t n,
                               const unsigned long long* __restrict__ bound, int nb,
                               unsigned long long* __restrict__ out) {
  const int g = threadIdx.x;
  if (g >= nb) return;
  uint64_t lo = 0, hi = n;
  const unsigned long long b = bound[g];
  while (lo < hi) {
    const uint64_t mid = (lo + hi) >> 1;
    if (keys[mid] < b) lo = mid + 1; else hi = mid;
  }
  out[g] = lo;
}

__global__ void k_distinct_degree(const unsigned long long* __restrict__ ukeys, uint64_t n, uint32_t base,
                                  uint32_t* __restrict__ deg) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) atomicAdd(&deg[(uint32_t)(ukeys[i] >> 32) - base], 1u);
}

__global__ void k_row_sectors(const uint32_t* __restrict__ deg, uint64_t V,
                              uint32_t* __restrict__ sectors, unsigned long long* __restrict__ deg64) {
  uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (; v <= V; v += stride) {
    uint32_t d = v < V ? deg[v] : 0u;
    sectors[v] = (d + 7u) >> 3;
    deg64[v] = d;
  }
}

__global__ void k_scatter_cols(const unsigned long long* __restrict__ ukeys, uint64_t n, uint32_t base,
                               const uint32_t* __restrict__ rowblk,
                               const unsigned long long* __restrict__ ustart,
                               uint32_t* __restrict__ col0) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    unsigned long long k = ukeys[i];
    uint32_t v = (uint32_t)(k >> 32) - base;
    uint64_t pos = (uint64_t)rowblk[v] * 8 + (i - ustart[v]);
    col0[pos] = (uint32_t)k;
  }
}

__global__ void k_csr_degrees(const unsigned long long* __restrict__ rowptr,
                              const unsigned long long* __restrict__ degm64, uint64_t V,
                              uint32_t* __restrict__ deg, uint32_t* __restrict__ degm,
                              uint32_t* __restrict__ sectors, unsigned long long* __restrict__ degm_total) {
  uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  unsigned long long sum = 0;
  for (; v <= V; v += stride) {
    uint32_t d = 0;
    if (v < V) {
      d = (uint32_t)(rowptr[v + 1] - rowptr[v]);
      deg[v] = d;
      const unsigned long long dm = degm64[v];
      degm[v] = (uint32_t)dm;
      sum += dm;
    }
    sectors[v] = (d + 7u) >> 3;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  if ((threadIdx.x & 31) == 0 && sum) atomicAdd(degm_total, sum);
}

// one warp copies one row of the compact CSR into its padded, sector aligned slot range (rows [v0, v1);
// the padding slots are written too, so the destination needs no memset)
__global__ void k_csr_place(const unsigned long long* __restrict__ rowptr, const uint32_t* __restrict__ col,
                            uint64_t v0, uint64_t v1, const uint32_t* __restrict__ rowblk, uint32_t* __restrict__ col0) {
  const uint32_t lane = threadIdx.x & 31;
  uint64_t v = v0 + (((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  for (; v < v1; v += nwarps) {
    const unsigned long long b = rowptr[v], e = rowptr[v + 1];
    const uint64_t o = (uint64_t)rowblk[v] * 8;
    const unsigned long long d = e - b, padded = (d + 7ull) & ~7ull;
    for (unsigned long long j = lane; j < padded; j += 32) col0[o + j] = j < d ? col[b + j] : PM_SENTINEL;
  }
}

// Several ranks: the neighbours of a row arrive ascending by VERTEX id and must leave ascending by SLOT,
// slot(v) = (v mod G) * nlmax + v / G.  Inside a row the G residue classes keep their order and follow one
// another, so every neighbour's position is (neighbours of smaller residue) + (its rank inside its class):
// two passes of ballots over the row, no sort.  One warp per row; rows [v0, v1); padding written too.
__global__ void k_csr_place_slots(const unsigned long long* __restrict__ rowptr, const uint32_t* __restrict__ col,
                                  uint64_t v0, uint64_t v1, const uint32_t* __restrict__ rowblk,
                                  uint32_t* __restrict__ col0, uint32_t G, uint32_t nlmax) {
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t lt = (1u << lane) - 1u;
  uint64_t v = v0 + (((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  for (; v < v1; v += nwarps) {
    const unsigned long long b = rowptr[v], e = rowptr[v + 1];
    const uint64_t o = (uint64_t)rowblk[v] * 8;
    const unsigned long long d = e - b, padded = (d + 7ull) & ~7ull;
    uint32_t start[PM_MAX_RANKS];
#pragma unroll
    for (int g = 0; g < PM_MAX_RANKS; ++g) start[g] = 0u;
    for (unsigned long long j0 = 0; j0 < d; j0 += 32) {  // neighbours per residue class
      const bool act = j0 + lane < d;
      const uint32_t r = act ? col[b + j0 + lane] % G : 0u;
#pragma unroll
      for (int g = 0; g < PM_MAX_RANKS; ++g)
        if ((uint32_t)g < G) start[g] += __popc(__ballot_sync(0xffffffffu, act && r == (uint32_t)g));
    }
    uint32_t run = 0;
#pragma unroll
    for (int g = 0; g < PM_MAX_RANKS; ++g) {  // counts -> first position of every class
      const uint32_t n = start[g];
      start[g] = run;
      run += n;
    }
    for (unsigned long long j0 = 0; j0 < d; j0 += 32) {
      const bool act = j0 + lane < d;
      const uint32_t u = act ? col[b + j0 + lane] : 0u;
      const uint32_t r = u % G;
#pragma unroll
      for (int g = 0; g < PM_MAX_RANKS; ++g)
        if ((uint32_t)g < G) {
          const uint32_t m = __ballot_sync(0xffffffffu, act && r == (uint32_t)g);
          if (act && r == (uint32_t)g) col0[o + start[g] + __popc(m & lt)] = r * nlmax + u / G;
          start[g] += __popc(m);
        }
    }
    for (unsigned long long j = d + lane; j < padded; j += 32) col0[o + j] = PM_SENTINEL;
  }
}

__global__ void k_fill_u64(unsigned long long* __restrict__ p, uint64_t n, unsigned long long value) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) p[i] = value;
}

// label = ceil(log2(degree + 1)) == bit length of the degree
// (include/havoqgt/vertex_data_db_degree.hpp:109; exact integer form)
__global__ void k_labels_degree_log2(const uint32_t* __restrict__ degm, uint64_t V,
                                     uint64_t* __restrict__ label) {
  uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (; v < V; v += stride) label[v] = (uint64_t)(32 - __clz(degm[v]));
}

__global__ void k_labels_to_bytes(const uint64_t* __restrict__ label, uint64_t V, uint8_t* __restrict__ lab8) {
  uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (; v < V; v += stride) lab8[v] = (uint8_t)label[v];
}

// Neighbour-label signature: sig[v] = OR over the distinct neighbours u of v of (1 << label[u]).
// It is a property of graph + labels (built once, next to the labels, outside any search) and
// lets the first LCC superstep decide most candidates without walking their rows.
// 8-lane groups take rows of up to kSigBig slots; longer rows are queued for k_build_sig_big.
#define PM_SIG_BIG 2048u
// idmask: the id bits of a col0 slot as it stands (a previous labelling may have packed labels into it);
// shift != 0: pack the new label of every neighbour into its slot (and write no label stream), else strip.
__global__ void __launch_bounds__(256) k_build_sig(const uint32_t* __restrict__ rowblk, const uint32_t* __restrict__ deg,
                                                   uint32_t* __restrict__ col0, const uint8_t* __restrict__ lab8,
                                                   uint64_t V, unsigned long long* __restrict__ sig,
                                                   uint8_t* __restrict__ lab0, uint32_t idmask, uint32_t shift,
                                                   uint32_t* __restrict__ big_list, uint32_t* __restrict__ big_n) {
  const uint32_t lane = threadIdx.x & 31, gl = lane & 7, gw = lane >> 3;
  const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  for (uint64_t base = warp * 4; base < V; base += nwarps * 4) {
    const uint64_t v = base + gw;
    uint32_t d = v < V ? deg[v] : 0u;
    if (d > PM_SIG_BIG) {
      if (gl == 0) big_list[atomicAdd(big_n, 1u)] = (uint32_t)v;
      d = 0;
    }
    const uint64_t row = v < V ? (uint64_t)rowblk[v] * 8 : 0;
    const uint32_t passes = (d + 31) / 32;
    const uint32_t maxp = __reduce_max_sync(0xffffffffu, passes);
    unsigned long long m = 0;
    for (uint32_t p = 0; p < maxp; ++p) {
      const uint32_t j0 = p * 32 + gl * 4;
      if (j0 < d) {
        uint4 q = *reinterpret_cast<const uint4*>(col0 + row + j0);
        const uint32_t i0 = q.x & idmask, i1 = q.y & idmask, i2 = q.z & idmask, i3 = q.w & idmask;
        const uint32_t l0 = lab8[i0];
        const uint32_t l1 = j0 + 1 < d ? (uint32_t)lab8[i1] : 0u;
        const uint32_t l2 = j0 + 2 < d ? (uint32_t)lab8[i2] : 0u;
        const uint32_t l3 = j0 + 3 < d ? (uint32_t)lab8[i3] : 0u;
        m |= 1ull << l0;
        if (j0 + 1 < d) m |= 1ull << l1;
        if (j0 + 2 < d) m |= 1ull << l2;
        if (j0 + 3 < d) m |= 1ull << l3;
        if (lab0)  // the label stream of the row (padding slots carry label 0); rows start sector aligned
          *reinterpret_cast<uint32_t*>(lab0 + row + j0) = l0 | (l1 << 8) | (l2 << 16) | (l3 << 24);
        if (shift != 0u || idmask != 0xFFFFFFFFu) {  // (re)pack or strip the labels carried by the slots themselves
          q.x = shift ? i0 | (l0 << shift) : i0;
          if (j0 + 1 < d) q.y = shift ? i1 | (l1 << shift) : i1;
          if (j0 + 2 < d) q.z = shift ? i2 | (l2 << shift) : i2;
          if (j0 + 3 < d) q.w = shift ? i3 | (l3 << shift) : i3;
          *reinterpret_cast<uint4*>(col0 + row + j0) = q;
        }
      } else if (lab0 && j0 < ((d + 7u) & ~7u)) {
        *reinterpret_cast<uint32_t*>(lab0 + row + j0) = 0u;
      }
    }
    m |= __shfl_xor_sync(0xffffffffu, m, 1);
    m |= __shfl_xor_sync(0xffffffffu, m, 2);
    m |= __shfl_xor_sync(0xffffffffu, m, 4);
    if (v < V && gl == 0 && deg[v] <= PM_SIG_BIG) sig[v] = m;
  }
}

__global__ void __launch_bounds__(1024) k_build_sig_big(const uint32_t* __restrict__ rowblk, const uint32_t* __restrict__ deg,
                                                        uint32_t* __restrict__ col0, const uint8_t* __restrict__ lab8,
                                                        unsigned long long* __restrict__ sig, uint8_t* __restrict__ lab0,
                                                        uint32_t idmask, uint32_t shift,
                                                        const uint32_t* __restrict__ big_list, const uint32_t* __restrict__ big_n) {
  __shared__ unsigned long long s_m[32];
  const uint32_t n = *big_n;
  for (uint32_t i = blockIdx.x; i < n; i += gridDim.x) {
    const uint32_t v = big_list[i], d = deg[v];
    const uint64_t row = (uint64_t)rowblk[v] * 8;
    unsigned long long m = 0;
    const uint32_t padded = (d + 7u) & ~7u;
    for (uint32_t j = threadIdx.x; j < padded; j += blockDim.x) {
      const uint32_t id = j < d ? col0[row + j] & idmask : 0u;
      const uint8_t l = j < d ? lab8[id] : (uint8_t)0;
      if (j < d) {
        m |= 1ull << l;
        if (shift != 0u || idmask != 0xFFFFFFFFu) col0[row + j] = shift ? id | ((uint32_t)l << shift) : id;
      }
      if (lab0) lab0[row + j] = l;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m |= __shfl_xor_sync(0xffffffffu, m, o);
    if ((threadIdx.x & 31) == 0) s_m[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned long long t = 0;
      for (uint32_t w = 0; w < (blockDim.x >> 5); ++w) t |= s_m[w];
      sig[v] = t;
    }
    __syncthreads();
  }
}

// col0 slots back to plain neighbour ids (the labels they carried no longer apply)
__global__ void k_col0_strip(uint32_t* __restrict__ col0, uint64_t n, uint32_t idmask) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) {
    const uint32_t q = col0[i];
    if (q != PM_SENTINEL) col0[i] = q & idmask;
  }
}

// the label stream of the run_fuzzy path from packed slots (values in padding slots are never read: rows are
// walked up to their degree)
__global__ void k_unpack_lab0(const uint32_t* __restrict__ col0, uint64_t n, uint32_t shift, uint8_t* __restrict__ lab0) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) lab0[i] = (uint8_t)(col0[i] >> shift);
}

inline uint32_t col_idmask(const pm_ctx* c) { return c->col_shift ? (1u << c->col_shift) - 1u : 0xFFFFFFFFu; }

inline int ensure_lab0(pm_ctx* c) {
  if (c->lab0 || !c->labels_small) return 0;
  int rc;
  if ((rc = dev_alloc(c, &c->lab0, c->Epad + 64, &c->graph_bytes))) return rc;
  k_unpack_lab0<<<grid_for(), kBlock, 0, c->stream>>>(c->col0, c->Epad + 64, c->col_shift, c->lab0);
  c->launches++;
  return 0;
}

// after the labels changed: byte labels + neighbour-label signatures (labels < 64 only).
// c->label holds the labels of the LOCAL rows; lab8 is replicated (indexed by slot) because the
// label of a neighbour owned by another rank is needed to build lab0 and the signatures.
// max_label: largest label value (sizes the packed label field)
inline int labels_derive(pm_ctx* c, bool small, uint64_t max_label) {
  dev_free(c->lab8);
  dev_free(c->sig);
  dev_free(c->lab0);
  c->labels_small = false;
  const uint32_t old_mask = col_idmask(c);
  if (!small) {
    if (c->col_shift) {
      k_col0_strip<<<grid_for(), kBlock, 0, c->stream>>>(c->col0, c->Epad, old_mask);
      c->launches++;
      c->col_shift = 0;
    }
    return 0;
  }
  // packed labels: the label of every neighbour rides in the unused high bits of its slot when
  // id bits + label bits <= 32 (R-MAT scale 26 on one GPU with degree labels: 26 + 6)
  // The label field is at most 6 bits wide (labels index 64-bit sets) and its all-ones value is left to the padding
  // slots (PM_SENTINEL), which therefore never test valid.
  uint32_t idbits = 1;
  while (idbits < 32 && ((c->nlmax * c->n_ranks - 1) >> idbits)) ++idbits;
  const uint32_t cand_shift = std::max<uint32_t>(idbits, 26u);
  const bool fits = cand_shift < 32 && max_label < (1ull << (32 - cand_shift)) - 1;
  const uint32_t new_shift = (fits && !getenv("PM_NO_PACK")) ? cand_shift : 0u;
  int rc;
  uint32_t *big_list = nullptr, *big_n = nullptr;
  const uint64_t Vs = c->nlmax * c->n_ranks, base = c->nlmax * c->rank;
  if ((rc = dev_alloc(c, &c->lab8, Vs, &c->graph_bytes))) return rc;
  if ((rc = dev_alloc(c, &c->sig, c->nloc, &c->graph_bytes))) return rc;
  if (!new_shift && (rc = dev_alloc(c, &c->lab0, c->Epad + 64, &c->graph_bytes))) return rc;
  if ((rc = dev_alloc(c, &big_list, c->Epad / PM_SIG_BIG + 1024))) return rc;
  if ((rc = dev_alloc(c, &big_n, 1))) { dev_free(big_list); return rc; }
  cudaStream_t st = c->stream;
  cudaMemsetAsync(big_n, 0, 4, st);
  k_labels_to_bytes<<<grid_for(), kBlock, 0, st>>>(c->label, c->nloc, c->lab8 + base);
  if ((rc = comm_allgather_slots(c, c->lab8))) { dev_free(big_list); dev_free(big_n); return rc; }
  // signatures and the label stream lab0 come out of the same pass over the adjacency (one gather per slot)
  if (c->lab0) cudaMemsetAsync(c->lab0 + c->Epad, 0, 64, st);
  k_build_sig<<<grid_for(), 256, 0, st>>>(c->rowblk, c->deg, c->col0, c->lab8, c->nloc, c->sig, c->lab0, old_mask, new_shift,
                                          big_list, big_n);
  k_build_sig_big<<<148, 1024, 0, st>>>(c->rowblk, c->deg, c->col0, c->lab8, c->sig, c->lab0, old_mask, new_shift, big_list, big_n);
  c->col_shift = new_shift;
  c->launches += 3;
  cudaError_t e = cudaGetLastError();
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  dev_free(big_list);
  dev_free(big_n);
  if (e != cudaSuccess) return fail(c, PM_ERR_CUDA, std::string("labels_derive: ") + cudaGetErrorString(e));
  c->labels_small = true;
  return 0;
}

inline void graph_free(pm_ctx* c) {
  comm_close_all(c);  // peers' mappings of the previous store
  dev_free(c->lab8);
  dev_free(c->sig);
  dev_free(c->lab0);
  c->labels_small = false;
  dev_free(c->rowblk);
  dev_free(c->deg);
  dev_free(c->degm);
  dev_free(c->col0);
  c->col_shift = 0;
  dev_free(c->hub_ctl);
  c->hubs.clear();
  c->delegate_threshold = 0;
  dev_free(c->label);
  c->has_graph = c->has_labels = false;
  c->graph_bytes = 0;
}

// slots per rank: every rank's range starts 16-aligned so that the vectorised kernels can treat the
// local part of a replicated array like a whole array
inline void graph_set_partition(pm_ctx* c, uint64_t V) {
  c->V = V;
  if (c->n_ranks == 1) {
    c->nlmax = V;
  } else {
    // 4096-aligned: the compact id tables (pm_lcc.cuh) work on tiles of 4096 slots that must not straddle ranks
    const uint64_t per = (V + c->n_ranks - 1) / c->n_ranks;
    c->nlmax = (per + 4095) / 4096 * 4096;
  }
  c->nloc = c->nlmax;
}

// Moves every key to the rank that owns its source slot.  keys: sorted ascending, n entries (device).
// On return *out / *n_out hold this rank's keys (unsorted concatenation of G sorted runs).
inline int graph_route_keys(pm_ctx* c, const unsigned long long* keys, uint64_t n, unsigned long long** out,
                            uint64_t* n_out) {
  const int G = c->n_ranks;
  cudaStream_t st = c->stream;
  unsigned long long *d_bound = nullptr, *d_lb = nullptr, *d_all = nullptr;
  int rc;
  if ((rc = dev_alloc(c, &d_bound, G + 1)) || (rc = dev_alloc(c, &d_lb, G + 1)) || (rc = dev_alloc(c, &d_all, (uint64_t)G * G))) {
    dev_free(d_bound); dev_free(d_lb); dev_free(d_all);
    return rc;
  }
  std::vector<unsigned long long> bound(G + 1), lb(G + 1);
  for (int g = 0; g <= G; ++g) bound[g] = (unsigned long long)(c->nlmax * g) << 32;
  cudaMemcpyAsync(d_bound, bound.data(), 8 * (G + 1), cudaMemcpyHostToDevice, st);
  k_lower_bounds<<<1, 32, 0, st>>>(keys, n, d_bound, G + 1, d_lb);
  c->launches++;
  cudaMemcpyAsync(lb.data(), d_lb, 8 * (G + 1), cudaMemcpyDeviceToHost, st);
  cudaError_t e = cudaStreamSynchronize(st);
  std::vector<unsigned long long> cnt(G), all((size_t)G * G);
  for (int g = 0; g < G; ++g) cnt[g] = lb[g + 1] - lb[g];
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_all + (size_t)c->rank * G, cnt.data(), 8 * G, cudaMemcpyHostToDevice, st);
  ncclResult_t nr = ncclSuccess;
  if (e == cudaSuccess) nr = ncclAllGather(d_all + (size_t)c->rank * G, d_all, G, ncclUint64, comm_of(c), st);
  if (e == cudaSuccess && nr == ncclSuccess) e = cudaMemcpyAsync(all.data(), d_all, 8 * (size_t)G * G, cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess && nr == ncclSuccess) e = cudaStreamSynchronize(st);
  dev_free(d_bound); dev_free(d_lb); dev_free(d_all);
  if (e != cudaSuccess) return fail(c, PM_ERR_CUDA, std::string("graph_route_keys: ") + cudaGetErrorString(e));
  if (nr != ncclSuccess) return fail(c, PM_ERR_COMM, std::string("graph_route_keys: ") + ncclGetErrorString(nr));
  uint64_t total = 0;
  std::vector<uint64_t> roff(G);
  for (int g = 0; g < G; ++g) { roff[g] = total; total += all[(size_t)g * G + c->rank]; }
  unsigned long long* recv = nullptr;
  if ((rc = dev_alloc(c, &recv, total))) return rc;
  // chunked so that no single message exceeds 2^30 keys
  const uint64_t chunk = 1ull << 30;
  nr = ncclGroupStart();
  for (int g = 0; g < G && nr == ncclSuccess; ++g) {
    for (uint64_t o = 0; o < cnt[g] && nr == ncclSuccess; o += chunk)
      nr = ncclSend(keys + lb[g] + o, (size_t)std::min<uint64_t>(chunk, cnt[g] - o), ncclUint64, g, comm_of(c), st);
    const uint64_t rn = all[(size_t)g * G + c->rank];
    for (uint64_t o = 0; o < rn && nr == ncclSuccess; o += chunk)
      nr = ncclRecv(recv + roff[g] + o, (size_t)std::min<uint64_t>(chunk, rn - o), ncclUint64, g, comm_of(c), st);
  }
  if (nr == ncclSuccess) nr = ncclGroupEnd();
  if (nr == ncclSuccess && cudaStreamSynchronize(st) != cudaSuccess) nr = ncclUnhandledCudaError;
  if (nr != ncclSuccess) { dev_free(recv); return fail(c, PM_ERR_COMM, std::string("edge exchange: ") + ncclGetErrorString(nr)); }
  *out = recv;
  *n_out = total;
  return 0;
}

// d_src / d_dst: device arrays of n directed slots with GLOBAL vertex ids (consumed, freed by the caller).
// route = true: the ranks hold disjoint parts of one edge list and every slot is sent to the owner of
// its source (the reference's edge shuffle, ipp:401-520); route = false: slots with a foreign source
// are dropped (every rank was handed the whole list).
inline int graph_build_from_device_slots(pm_ctx* c, uint64_t V, uint64_t n_in, const uint32_t* d_src,
                                         const uint32_t* d_dst, bool route) {
  if (V == 0 || V > (1ull << 31)) return fail(c, PM_ERR_ARG, "n_vertices must be in [1, 2^31]");
  graph_free(c);
  graph_set_partition(c, V);
  cudaStream_t st = c->stream;
  const uint64_t NL = c->nloc;
  const uint32_t base = (uint32_t)(c->nlmax * c->rank);
  const int G = c->n_ranks;
  uint64_t bytes = 0;
  int rc;
  if ((rc = dev_alloc(c, &c->degm, NL, &bytes))) return rc;
  if ((rc = dev_alloc(c, &c->deg, NL, &bytes))) return rc;
  if ((rc = dev_alloc(c, &c->rowblk, NL + 1, &bytes))) return rc;
  PM_CUDA(c, cudaMemsetAsync(c->degm, 0, NL * sizeof(uint32_t), st));
  PM_CUDA(c, cudaMemsetAsync(c->deg, 0, NL * sizeof(uint32_t), st));

  unsigned long long *keys = nullptr, *keys2 = nullptr, *ustart = nullptr, *d_nuniq = nullptr, *deg64 = nullptr;
  uint32_t* sectors = nullptr;
  void* tmp = nullptr;
  size_t tmp_bytes = 0;
  auto cleanup = [&]() {
    dev_free(keys); dev_free(keys2); dev_free(ustart); dev_free(d_nuniq); dev_free(sectors); dev_free(deg64);
    if (tmp) cudaFree(tmp);
    tmp = nullptr;
  };
#define PM_G(call) do { int rc_ = (call); if (rc_) { cleanup(); return rc_; } } while (0)
#define PM_GC(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { cleanup(); \
    return fail(c, PM_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); } } while (0)
  auto need_tmp = [&](size_t want) -> cudaError_t {
    if (want <= tmp_bytes) return cudaSuccess;
    if (tmp) cudaFree(tmp);
    tmp = nullptr;
    tmp_bytes = std::max<size_t>(want, 16);
    return cudaMalloc(&tmp, tmp_bytes);
  };
  const int grid = grid_for();
  // only the bits that can be set take part in the sorts
  int bits = 32;
  while (bits < 64 && ((c->nlmax * G - 1) >> (bits - 32))) ++bits;
  uint64_t n = n_in;
  PM_G(dev_alloc(c, &keys, n));
  PM_G(dev_alloc(c, &keys2, n));
  PM_G(dev_alloc(c, &d_nuniq, 2));
  if (n) {
    k_slot_keys<<<grid, kBlock, 0, st>>>(d_src, d_dst, n, (uint32_t)G, (uint32_t)c->nlmax, (uint32_t)c->rank,
                                         (G > 1 && !route) ? 1 : 0, keys);
    c->launches++;
    PM_GC(cudaGetLastError());
  }
  if (G > 1) {
    // sort once so that every owner's keys are one contiguous range (foreign keys, if dropped, sort last)
    size_t tb = 0;
    cub::DoubleBuffer<unsigned long long> db0(keys, keys2);
    PM_GC(cub::DeviceRadixSort::SortKeys(nullptr, tb, db0, (int64_t)n, 0, 64, st));
    PM_GC(need_tmp(tb));
    PM_GC(cub::DeviceRadixSort::SortKeys(tmp, tb, db0, (int64_t)n, 0, 64, st));
    if (db0.Current() != keys) std::swap(keys, keys2);
    if (route) {
      unsigned long long* mine = nullptr;
      uint64_t n_mine = 0;
      PM_G(graph_route_keys(c, keys, n, &mine, &n_mine));
      dev_free(keys);
      dev_free(keys2);
      keys = mine;
      n = n_mine;
      PM_G(dev_alloc(c, &keys2, n));
    } else {
      // keep [lower_bound(base), lower_bound(base + nlmax))
      unsigned long long h_b[2] = {(unsigned long long)base << 32, (unsigned long long)(base + c->nlmax) << 32}, h_lb[2];
      unsigned long long* d_b = d_nuniq;  // scratch: 2 entries
      unsigned long long* d_lb = nullptr;
      PM_G(dev_alloc(c, &d_lb, 2));
      cudaMemcpyAsync(d_b, h_b, 16, cudaMemcpyHostToDevice, st);
      k_lower_bounds<<<1, 32, 0, st>>>(keys, n, d_b, 2, d_lb);
      cudaMemcpyAsync(h_lb, d_lb, 16, cudaMemcpyDeviceToHost, st);
      cudaError_t e = cudaStreamSynchronize(st);
      dev_free(d_lb);
      PM_GC(e);
      const uint64_t n_mine = h_lb[1] - h_lb[0];
      PM_GC(cudaMemcpyAsync(keys2, keys + h_lb[0], n_mine * 8, cudaMemcpyDeviceToDevice, st));
      std::swap(keys, keys2);
      n = n_mine;
    }
  }
  c->E_multi = n;
  if (n) {
    k_degm_of_keys<<<grid, kBlock, 0, st>>>(keys, n, base, c->degm);
    c->launches++;
    PM_GC(cudaGetLastError());
  }
  size_t tb1 = 0, tb2 = 0;
  cub::DoubleBuffer<unsigned long long> db(keys, keys2);
  PM_GC(cub::DeviceRadixSort::SortKeys(nullptr, tb1, db, (int64_t)n, 0, bits, st));
  PM_GC(cub::DeviceSelect::Unique(nullptr, tb2, keys, keys2, d_nuniq, (int64_t)n, st));
  PM_GC(need_tmp(std::max(tb1, tb2)));
  PM_GC(cub::DeviceRadixSort::SortKeys(tmp, tb1, db, (int64_t)n, 0, bits, st));
  unsigned long long* sorted = db.Current();
  unsigned long long* ukeys = db.Alternate();
  PM_GC(cub::DeviceSelect::Unique(tmp, tb2, sorted, ukeys, d_nuniq, (int64_t)n, st));
  unsigned long long h_nuniq = 0;
  PM_GC(cudaMemcpyAsync(&h_nuniq, d_nuniq, sizeof(h_nuniq), cudaMemcpyDeviceToHost, st));
  PM_GC(cudaStreamSynchronize(st));
  c->E = h_nuniq;
  if (h_nuniq) {
    k_distinct_degree<<<grid, kBlock, 0, st>>>(ukeys, h_nuniq, base, c->deg);
    c->launches++;
    PM_GC(cudaGetLastError());
  }
  // row starts (in sectors) and starts in the unique list
  PM_G(dev_alloc(c, &sectors, NL + 1));
  PM_G(dev_alloc(c, &ustart, NL + 1));
  PM_G(dev_alloc(c, &deg64, NL + 1));
  k_row_sectors<<<grid, kBlock, 0, st>>>(c->deg, NL, sectors, deg64);
  c->launches++;
  size_t tb3 = 0, tb4 = 0, tb5 = 0;
  uint32_t* d_max = (uint32_t*)d_nuniq;
  cub::DeviceScan::ExclusiveSum(nullptr, tb3, sectors, c->rowblk, (int64_t)(NL + 1), st);
  cub::DeviceScan::ExclusiveSum(nullptr, tb4, deg64, ustart, (int64_t)(NL + 1), st);
  cub::DeviceReduce::Max(nullptr, tb5, c->degm, d_max, (int64_t)NL, st);
  PM_GC(need_tmp(std::max(tb3, std::max(tb4, tb5))));
  PM_GC(cub::DeviceScan::ExclusiveSum(tmp, tb3, sectors, c->rowblk, (int64_t)(NL + 1), st));
  PM_GC(cub::DeviceScan::ExclusiveSum(tmp, tb4, deg64, ustart, (int64_t)(NL + 1), st));
  uint32_t h_total_sectors = 0, h_max = 0;
  PM_GC(cudaMemcpyAsync(&h_total_sectors, c->rowblk + NL, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  PM_GC(cub::DeviceReduce::Max(tmp, tb5, c->degm, d_max, (int64_t)NL, st));
  PM_GC(cudaMemcpyAsync(&h_max, d_max, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  PM_GC(cudaStreamSynchronize(st));
  c->max_deg = h_max;
  // +64 slots of slack: the scan kernels read whole uint4 windows
  c->Epad = (uint64_t)h_total_sectors * 8;
  const uint64_t alloc_slots = c->Epad + 64;
  PM_G(dev_alloc(c, &c->col0, alloc_slots, &bytes));
  PM_GC(cudaMemsetAsync(c->col0, 0xFF, alloc_slots * sizeof(uint32_t), st));
  if (h_nuniq) {
    k_scatter_cols<<<grid, kBlock, 0, st>>>(ukeys, h_nuniq, base, c->rowblk, ustart, c->col0);
    c->launches++;
    PM_GC(cudaGetLastError());
  }
  PM_GC(cudaStreamSynchronize(st));
  cleanup();
  dev_cache().flush(c->device);  // the sort buffers of a build are never asked for again
#undef PM_G
#undef PM_GC
  c->graph_bytes = bytes;
  c->has_graph = true;
  c->state_ready = false;
  return 0;
}

// Host CSR -> device store.  h_* are HOST pointers (copied here: this is the host->device traffic an
// end-to-end run pays).  The adjacency travels in chunks on a copy stream straight into the (still unused)
// working adjacency `colw`; each chunk's rows are placed into their padded positions of `col0` while the
// next chunk is still on the bus, so the PCIe copy is the only thing on the critical path.
// Several ranks: V is the global vertex count and the host arrays describe the n_rows vertices THIS rank owns
// (local row i = vertex i * n_ranks + rank, neighbours as global vertex ids ascending); one rank: n_rows = V.
inline int graph_build_from_host_csr(pm_ctx* c, uint64_t V, uint64_t n_rows, const uint64_t* h_rowptr,
                                     const uint32_t* h_col, const uint64_t* h_degm) {
  if (V == 0 || V > (1ull << 31)) return fail(c, PM_ERR_ARG, "n_vertices must be in [1, 2^31]");
  const bool dbg = getenv("PM_DEBUG_BUILD") != nullptr;
  double t_prev = wall_s();
  auto lap = [&](const char* what) {
    if (!dbg) return;
    cudaStreamSynchronize(c->stream);
    const double t = wall_s();
    fprintf(stderr, "[pm build] %-28s %.2f ms\n", what, (t - t_prev) * 1e3);
    t_prev = t;
  };
  graph_free(c);
  lap("graph_free");
  graph_set_partition(c, V);  // single rank: slot = vertex id, the rows can be placed as they are
  cudaStream_t st = c->stream;
  const uint64_t NL = c->nloc;  // local rows of the store (rows past n_rows are alignment padding: empty)
  const uint64_t E = h_rowptr[n_rows];
  c->E = E;
  uint64_t bytes = 0;
  int rc;
  unsigned long long *d_rowptr = nullptr, *d_degm64 = nullptr, *d_total = nullptr;
  uint32_t* sectors = nullptr;
  void* tmp = nullptr;
  cudaStream_t cs = nullptr;
  std::vector<cudaEvent_t> evs;
  uint32_t* stage[2] = {nullptr, nullptr};
  auto cleanup = [&]() {
    dev_free(d_rowptr); dev_free(d_degm64); dev_free(d_total); dev_free(sectors);
    dev_free(stage[0]); dev_free(stage[1]);
    if (tmp) cudaFree(tmp);
    tmp = nullptr;
    for (auto e : evs) cudaEventDestroy(e);
    evs.clear();
    if (cs) cudaStreamDestroy(cs);
    cs = nullptr;
  };
  if ((rc = dev_alloc(c, &c->degm, NL, &bytes)) || (rc = dev_alloc(c, &c->deg, NL, &bytes)) ||
      (rc = dev_alloc(c, &c->rowblk, NL + 1, &bytes)) || (rc = dev_alloc(c, &d_rowptr, NL + 1)) ||
      (rc = dev_alloc(c, &d_degm64, NL)) || (rc = dev_alloc(c, &d_total, 1)) || (rc = dev_alloc(c, &sectors, NL + 1))) {
    cleanup();
    return rc;
  }
#define PM_GC(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { cleanup(); \
    return fail(c, PM_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); } } while (0)
  lap("alloc 1");
  PM_GC(cudaMemcpyAsync(d_rowptr, h_rowptr, (n_rows + 1) * 8, cudaMemcpyHostToDevice, st));
  PM_GC(cudaMemcpyAsync(d_degm64, h_degm, n_rows * 8, cudaMemcpyHostToDevice, st));
  PM_GC(cudaMemsetAsync(d_total, 0, 8, st));
  const int grid = grid_for();
  if (NL > n_rows) {
    k_fill_u64<<<grid, kBlock, 0, st>>>(d_rowptr + n_rows + 1, NL - n_rows, (unsigned long long)E);
    PM_GC(cudaMemsetAsync(d_degm64 + n_rows, 0, (NL - n_rows) * 8, st));
  }
  k_csr_degrees<<<grid, kBlock, 0, st>>>(d_rowptr, d_degm64, NL, c->deg, c->degm, sectors, d_total);
  c->launches++;
  size_t tb = 0, tb2 = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, tb, sectors, c->rowblk, (int64_t)(NL + 1), st);
  uint32_t* d_max = nullptr;
  cub::DeviceReduce::Max(nullptr, tb2, c->degm, d_max, (int64_t)NL, st);
  tb = std::max(tb, tb2);
  PM_GC(cudaMalloc(&tmp, std::max<size_t>(tb, 16) + 16));
  PM_GC(cub::DeviceScan::ExclusiveSum(tmp, tb, sectors, c->rowblk, (int64_t)(NL + 1), st));
  uint32_t h_total = 0, h_max = 0;
  unsigned long long h_em = 0;
  PM_GC(cudaMemcpyAsync(&h_total, c->rowblk + NL, 4, cudaMemcpyDeviceToHost, st));
  d_max = sectors;  // reuse: sectors are consumed
  PM_GC(cub::DeviceReduce::Max(tmp, tb, c->degm, d_max, (int64_t)NL, st));
  PM_GC(cudaMemcpyAsync(&h_max, d_max, 4, cudaMemcpyDeviceToHost, st));
  PM_GC(cudaMemcpyAsync(&h_em, d_total, 8, cudaMemcpyDeviceToHost, st));
  PM_GC(cudaStreamSynchronize(st));
  c->max_deg = h_max;
  c->E_multi = h_em;  // multigraph slot count = sum of multigraph degrees
  c->Epad = (uint64_t)h_total * 8;
  lap("row pointers, degrees, scan");
  const uint64_t alloc_slots = c->Epad + 64;
  if ((rc = dev_alloc(c, &c->col0, alloc_slots, &bytes))) {
    cleanup();
    return rc;
  }
  lap("alloc col0");
  PM_GC(cudaMemsetAsync(c->col0 + c->Epad, 0xFF, 64 * 4, st));
  if (E) {
    PM_GC(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking));
    // chunks of whole rows, >= 64 MiB each; two staging buffers: chunk k + 1 travels while chunk k is placed
    const uint64_t n_chunks = std::max<uint64_t>(1, std::min<uint64_t>(32, E >> 24));
    std::vector<std::pair<uint64_t, uint64_t>> chunks;
    uint64_t v0 = 0, max_slots = 0;
    for (uint64_t k = 1; k <= n_chunks && v0 < n_rows; ++k) {
      // rows [v0, v1): v1 = first row that starts at or after the k-th share of the slots
      uint64_t v1 = n_rows;
      if (k < n_chunks) v1 = (uint64_t)(std::lower_bound(h_rowptr + v0, h_rowptr + n_rows, E / n_chunks * k) - h_rowptr);
      if (v1 <= v0) continue;
      chunks.push_back({v0, v1});
      max_slots = std::max<uint64_t>(max_slots, h_rowptr[v1] - h_rowptr[v0]);
      v0 = v1;
    }
    if ((rc = dev_alloc(c, &stage[0], max_slots + 4)) || (rc = dev_alloc(c, &stage[1], max_slots + 4))) {
      cleanup();
      return rc;
    }
    std::vector<cudaEvent_t> placed(chunks.size(), nullptr);
    for (size_t k = 0; k < chunks.size(); ++k) {
      const uint64_t a0 = chunks[k].first, a1 = chunks[k].second;
      const uint64_t b = h_rowptr[a0], e = h_rowptr[a1];
      uint32_t* buf = stage[k & 1];
      cudaEvent_t ev;
      PM_GC(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
      evs.push_back(ev);
      if (k >= 2) PM_GC(cudaStreamWaitEvent(cs, placed[k - 2], 0));  // the buffer's previous chunk has been placed
      if (e > b) PM_GC(cudaMemcpyAsync(buf, h_col + b, (e - b) * 4, cudaMemcpyHostToDevice, cs));
      PM_GC(cudaEventRecord(ev, cs));
      PM_GC(cudaStreamWaitEvent(st, ev, 0));
      // the kernels index the staged columns with the CSR's own offsets: hand them the buffer shifted by the chunk start
      const uint32_t* col_view = reinterpret_cast<const uint32_t*>(reinterpret_cast<uintptr_t>(buf) - (uintptr_t)b * 4);
      if (c->n_ranks == 1)
        k_csr_place<<<grid, kBlock, 0, st>>>(d_rowptr, col_view, a0, a1, c->rowblk, c->col0);
      else  // vertex ids -> slots, rows re-ordered by slot
        k_csr_place_slots<<<grid, kBlock, 0, st>>>(d_rowptr, col_view, a0, a1, c->rowblk, c->col0, (uint32_t)c->n_ranks,
                                                   (uint32_t)c->nlmax);
      c->launches++;
      PM_GC(cudaEventCreateWithFlags(&placed[k], cudaEventDisableTiming));
      evs.push_back(placed[k]);
      PM_GC(cudaEventRecord(placed[k], st));
    }
    PM_GC(cudaGetLastError());
  }
  PM_GC(cudaStreamSynchronize(st));
#undef PM_GC
  lap("adjacency copy + placement");
  cleanup();
  lap("free temporaries");
  c->graph_bytes = bytes;
  c->has_graph = true;
  c->state_ready = false;
  return 0;
}

}  // namespace pm
