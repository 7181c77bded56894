// pmgpu.cu — C ABI of libpmgpu.so (see include/pmgpu.h) and the host-side driver
// loop that mirrors /root/reference/src/run_pattern_matching_beta.cpp:544-1351.
// Compiled for sm_100a only.  There is no CPU path: every entry point that
// computes anything launches the kernels in pm_graph.cuh / pm_lcc.cuh / pm_nlcc.cuh.

#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>

#include "pm_common.cuh"
#include "pm_comm.cuh"
#include "pm_graph.cuh"
#include "pm_lcc.cuh"
#include "pm_nlcc.cuh"
#include "pm_nlcc_multi.cuh"
#include "pm_rmat.cuh"
#include "pm_fuzzy.cuh"
#include "pm_io.hpp"

using namespace pm;

namespace {

// c_pat / c_peer are __constant__ symbols shared by every context of a device.  The context that uploaded them last
// owns them; another live context on the same device re-uploads its own tables before it launches anything
// (after a device-wide synchronisation: the other context's kernels may still be reading the old ones).
pm_ctx* g_const_owner[64] = {nullptr};

int claim_constants(pm_ctx* c) {
  pm_ctx*& owner = g_const_owner[c->device & 63];
  if (owner == c) return 0;
  const bool contested = owner != nullptr;
  owner = c;
  if (!contested || !c->state_ready) return 0;  // pm_state_reset uploads both tables itself
  PM_CUDA(c, cudaDeviceSynchronize());
  PM_CUDA(c, cudaMemcpyToSymbolAsync(c_pat, &c->pc, sizeof(PatConst), 0, cudaMemcpyHostToDevice, c->stream));
  return comm_upload_peers(c);
}

int sync_counters(pm_ctx* c) {
  PM_CUDA(c, cudaMemcpyAsync(c->h_cnt, c->cnt, sizeof(DevCounters), cudaMemcpyDeviceToHost, c->stream));
  PM_CUDA(c, cudaStreamSynchronize(c->stream));
  return 0;
}

LccArgs lcc_args(pm_ctx* c, int row) {
  LccArgs a;
  a.col0 = c->col0; a.colw = c->colw;
  a.S = c->S; a.adeg = c->adeg; a.cls = c->cls; a.cnt = c->cnt;
  a.lab0 = c->lab0;
  a.row = c->rowstat + row;
  a.base = c->cid_off[c->rank];
  a.par = c->step_parity;
  a.fwx = c->fwx;
  a.rowc = c->rowc;
  a.col_shift = c->col_shift;
  a.typed = c->typed ? 1 : 0;
  a.clsc = c->clsc;
  a.hubc = (c->n_ranks > 1 && !c->hubs.empty() && !c->fuzzy_ids) ? c->hubc : nullptr;
  return a;
}

NlcArgs nlc_args(pm_ctx* c, uint2* matches, uint64_t match_cap) {
  NlcArgs a;
  a.rowblk = c->rowc; a.colw = c->colw; a.S = c->S; a.adeg = c->adeg; a.cls = c->clsc;
  a.ok = c->ok; a.src_list = c->src_list; a.hset = c->hset; a.hset_mask = c->hset_use - 1;
  a.pool = c->pool; a.pool_cap = c->pool_cap; a.matches = matches; a.match_cap = match_cap;
  a.cnt = c->cnt;
  a.base = c->cid_off[c->rank];
  a.vid = c->vid;
  a.par = c->step_parity;
  a.all = c->step_msg ? c->step_msg + 1 : nullptr;
  return a;
}

// number of vertices this rank owns (owner(v) = v mod G)
uint64_t n_owned(const pm_ctx* c) {
  return c->n_ranks == 1 ? c->V : (c->V + c->n_ranks - 1 - c->rank) / c->n_ranks;
}

// (Re)publishes every buffer the peers address directly and uploads the peer table.  Collective.
int comm_publish(pm_ctx* c) {
  comm_close_all(c);
  PeerTab& t = c->peers;
  std::memset(&t, 0, sizeof(t));
  t.G = c->n_ranks; t.rank = c->rank; t.nlmax = (uint32_t)c->nlmax; t.base = (uint32_t)(c->nlmax * c->rank);
  t.dcap = (uint32_t)c->dcap; t.tcap = c->tcap;
  for (int g = 0; g <= PM_MAX_RANKS; ++g) t.off[g] = c->cid_off[g];
  void* out[PM_MAX_RANKS];
  int rc;
#define PM_SHARE(buf, field, T)                                              \
  if (buf) {                                                                 \
    if ((rc = comm_share(c, (void*)(buf), out))) return rc;                  \
    for (int g = 0; g < c->n_ranks; ++g) t.field[g] = (T)out[g];             \
  }
  PM_SHARE(c->rowc, rowblk, const uint32_t*)
  PM_SHARE(c->adeg, adeg, const uint32_t*)
  PM_SHARE(c->colw, colw, uint32_t*)
  PM_SHARE(c->ok, ok, uint8_t*)
  PM_SHARE(c->din[0], din[0], uint2*)
  PM_SHARE(c->din[1], din[1], uint2*)
  PM_SHARE(c->tin[0], tin[0], uint2*)
  PM_SHARE(c->tin[1], tin[1], uint2*)
  PM_SHARE(c->sync_in, sync_in, StepMsg*)
#undef PM_SHARE
  return comm_upload_peers(c);
}

// keep_scratch: a new graph is about to be opened by the same context.  The NLCC token pool and key table
// are sized by what the searches needed, not by the graph, so a single rank keeps them (growing them again
// would cost one multi-GB cudaMalloc per step of the growth); several ranks release everything because the
// peers' mappings of the old buffers must go.
void state_free(pm_ctx* c, bool keep_scratch = false) {
  comm_close_all(c);  // nobody may still map the buffers freed below
  const bool keep = keep_scratch && c->n_ranks == 1;
  dev_free(c->din[0]); dev_free(c->din[1]); dev_free(c->tin[0]); dev_free(c->tin[1]); dev_free(c->step_msg);
  dev_free(c->sync_in);
  if (c->h_step) cudaFreeHost(c->h_step);
  c->h_step = nullptr;
  c->dcap = c->tcap = 0;
  dev_free(c->S); dev_free(c->adeg); dev_free(c->cls);
  if (!keep) { dev_free(c->colw); c->colw_cap = 0; }
  dev_free(c->rowc); dev_free(c->vid); dev_free(c->clsc); dev_free(c->fw); dev_free(c->hubc); dev_free(c->tb); dev_free(c->fwx);
  for (int b = 0; b < 2; ++b) for (int k = 0; k < 2; ++k) dev_free(c->fr[b][k]);
  dev_free(c->cnt); dev_free(c->ok); dev_free(c->src_list);
  if (!keep) {
    dev_free(c->hset); dev_free(c->pool);
    c->hset_cap = c->pool_cap = 0;
  }
  // the small pinned read-back buffers (h_cnt, h_rowstat) live as long as the context: pinned allocation
  // calls synchronise the device
  c->state_ready = false;
}

// NLCC scratch sizes (token pool, key table) remembered from earlier searches of the same pattern directory over
// the same graph through the same path; anything else starts from the defaults again
void load_nlcc_sizes(pm_ctx* c, const char* path) {
  const std::string key = c->pat_dir + "|" + std::to_string(c->V) + "|" + std::to_string(c->E_multi) + "|" +
                          std::to_string(c->labels_version) + "|" + path;
  if (key == c->pat_key) return;
  c->pat_key = key;
  const size_t n = c->pat.constraints.size();
  auto it = c->pool_cache.find(key);
  if (it != c->pool_cache.end() && it->second.size() == n) c->pool_seen = it->second;
  else c->pool_seen.assign(n, 0);
  auto it2 = c->keys_cache.find(key);
  if (it2 != c->keys_cache.end() && it2->second.size() == n) c->keys_seen = it2->second;
  else c->keys_seen.assign(n, 0);
}

// ---- delegates: attribution of hubs to their controller ranks (several ranks) -----------------------------------
// controller rank of vertex v if it is a hub, else -1 (delegate id = position in the ascending hub list,
// impl/delegate_partitioned_graph.ipp:501-512, 681; controller = delegate_id % ranks, delegate_partitioned_graph.hpp:231-233)
int hub_controller(const pm_ctx* c, uint64_t v) {
  if (c->hubs.empty()) return -1;
  auto it = std::lower_bound(c->hubs.begin(), c->hubs.end(), (uint32_t)v);
  if (it == c->hubs.end() || *it != (uint32_t)v) return -1;
  return (int)((it - c->hubs.begin()) % c->n_ranks);
}
bool hubs_on(const pm_ctx* c) { return c->n_ranks > 1 && !c->hubs.empty(); }

// Collective.  Adds to rows[0 .. n) the hubs the PEERS hold for this rank's controller role (RowStat::hub_nv / hub_ne
// of rowstat[first .. first + n) on every rank).  The device rows must be complete on the stream.
int hub_rows_merge(pm_ctx* c, int first, int n, pm_row_t* rows) {
  if (!hubs_on(c) || n <= 0) return 0;
  RowStat* d_all = nullptr;
  int rc = dev_alloc(c, &d_all, (uint64_t)n * c->n_ranks);
  if (rc) return rc;
  std::vector<RowStat> all((size_t)n * c->n_ranks);
  ncclResult_t nr = ncclAllGather(c->rowstat + first, d_all, (size_t)n * sizeof(RowStat), ncclChar, comm_of(c), c->stream);
  cudaError_t e = cudaMemcpyAsync(all.data(), d_all, all.size() * sizeof(RowStat), cudaMemcpyDeviceToHost, c->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
  dev_free(d_all);
  if (nr != ncclSuccess) return fail(c, PM_ERR_COMM, ncclGetErrorString(nr));
  if (e != cudaSuccess) return fail(c, PM_ERR_CUDA, cudaGetErrorString(e));
  for (int k = 0; k < n; ++k)
    for (int g = 0; g < c->n_ranks; ++g) {
      rows[k].n_vertices += all[(size_t)g * n + k].hub_nv[c->rank];
      rows[k].n_edges += all[(size_t)g * n + k].hub_ne[c->rank];
    }
  return 0;
}

// Collective.  Records of `width` words: those whose key vertex (word `key`) is a hub move to its controller rank;
// `data` comes back holding this rank's records (its own non-hub ones + the hubs it controls), unsorted.
int hub_records_exchange(pm_ctx* c, std::vector<uint32_t>& data, int width, int key) {
  if (!hubs_on(c)) return 0;
  const int G = c->n_ranks;
  std::vector<std::vector<uint32_t>> out(G);
  for (size_t i = 0; i + width <= data.size(); i += width) {
    const int ctl = hub_controller(c, data[i + key]);
    auto& dst = out[ctl < 0 ? c->rank : ctl];
    dst.insert(dst.end(), data.begin() + i, data.begin() + i + width);
  }
  // counts of every (sender, receiver) pair
  unsigned long long* d_cnt = nullptr;
  int rc = dev_alloc(c, &d_cnt, (uint64_t)G * G);
  if (rc) return rc;
  std::vector<unsigned long long> mine(G), all((size_t)G * G);
  for (int g = 0; g < G; ++g) mine[g] = out[g].size();
  cudaMemcpyAsync(d_cnt + (size_t)c->rank * G, mine.data(), 8 * G, cudaMemcpyHostToDevice, c->stream);
  ncclResult_t nr = ncclAllGather(d_cnt + (size_t)c->rank * G, d_cnt, G, ncclUint64, comm_of(c), c->stream);
  cudaError_t e = cudaMemcpyAsync(all.data(), d_cnt, 8 * (size_t)G * G, cudaMemcpyDeviceToHost, c->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
  dev_free(d_cnt);
  if (nr != ncclSuccess) return fail(c, PM_ERR_COMM, ncclGetErrorString(nr));
  if (e != cudaSuccess) return fail(c, PM_ERR_CUDA, cudaGetErrorString(e));
  uint64_t n_send = 0, n_recv = 0;
  std::vector<uint64_t> soff(G), roff(G);
  for (int g = 0; g < G; ++g) {
    soff[g] = n_send;
    roff[g] = n_recv;
    if (g != c->rank) { n_send += mine[g]; n_recv += all[(size_t)g * G + c->rank]; }
  }
  uint32_t *d_send = nullptr, *d_recv = nullptr;
  if ((rc = dev_alloc(c, &d_send, n_send + 1)) || (rc = dev_alloc(c, &d_recv, n_recv + 1))) { dev_free(d_send); return rc; }
  for (int g = 0; g < G; ++g)
    if (g != c->rank && mine[g]) cudaMemcpyAsync(d_send + soff[g], out[g].data(), mine[g] * 4, cudaMemcpyHostToDevice, c->stream);
  nr = ncclGroupStart();
  for (int g = 0; g < G && nr == ncclSuccess; ++g) {
    if (g == c->rank) continue;
    if (mine[g]) nr = ncclSend(d_send + soff[g], (size_t)mine[g], ncclUint32, g, comm_of(c), c->stream);
    const uint64_t rn = all[(size_t)g * G + c->rank];
    if (rn && nr == ncclSuccess) nr = ncclRecv(d_recv + roff[g], (size_t)rn, ncclUint32, g, comm_of(c), c->stream);
  }
  if (nr == ncclSuccess) nr = ncclGroupEnd();
  std::vector<uint32_t> got(n_recv);
  if (n_recv) e = cudaMemcpyAsync(got.data(), d_recv, n_recv * 4, cudaMemcpyDeviceToHost, c->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
  dev_free(d_send);
  dev_free(d_recv);
  if (nr != ncclSuccess) return fail(c, PM_ERR_COMM, ncclGetErrorString(nr));
  if (e != cudaSuccess) return fail(c, PM_ERR_CUDA, cudaGetErrorString(e));
  data = std::move(out[c->rank]);
  data.insert(data.end(), got.begin(), got.end());
  return 0;
}

// nem_1's result is independent of message arrival order iff interior hop labels
// are pairwise distinct and P[h-1] != P[h+1].
bool nem1_order_independent(const Constraint& k) {
  const size_t n = k.P.size();
  for (size_t a = 1; a + 1 < n; ++a)
    for (size_t b = a + 1; b + 1 < n; ++b)
      if (k.P[a] == k.P[b]) return false;
  for (size_t h = 1; h + 1 < n; ++h)
    if (k.P[h - 1] == k.P[h + 1]) return false;
  return true;
}

// Token storage for one constraint.  One rank: the token pool.  Several ranks: the two token inboxes
// (G sender regions of pool_cap tokens each), which every peer must map — growing them is collective,
// so the callers agree on pool_cap first.
int nlcc_reserve(pm_ctx* c, uint64_t pool_cap, uint64_t key_cap) {
  uint64_t hc_want = 1;
  while (hc_want < 2 * key_cap) hc_want <<= 1;
  if (pool_cap <= c->pool_cap && hc_want <= c->hset_cap) return 0;
  pool_cap = std::max(pool_cap, c->pool_cap);
  const bool multi = c->n_ranks > 1;
  if (multi) comm_close_all(c);
  dev_free(c->pool);
  dev_free(c->hset);
  dev_free(c->tin[0]);
  dev_free(c->tin[1]);
  c->pool_cap = c->hset_cap = c->tcap = 0;
  const uint64_t hc = hc_want;
  int rc;
  if (!multi) {
    if ((rc = dev_alloc(c, &c->pool, pool_cap))) return rc;
  } else {
    if ((rc = dev_alloc(c, &c->tin[0], pool_cap * c->n_ranks))) return rc;
    if ((rc = dev_alloc(c, &c->tin[1], pool_cap * c->n_ranks))) return rc;
  }
  if ((rc = dev_alloc(c, &c->hset, hc))) return rc;
  c->pool_cap = pool_cap;
  c->hset_cap = hc;
  if (multi) {
    c->tcap = pool_cap;
    if ((rc = comm_publish(c))) return rc;
  }
  return 0;
}

}  // namespace

extern "C" {

int pm_create(pm_ctx** out, int device) {
  if (!out) return PM_ERR_ARG;
  *out = nullptr;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) return PM_ERR_CUDA;  // no CPU fallback
  if (device < 0 || device >= n) return PM_ERR_ARG;
  if (cudaSetDevice(device) != cudaSuccess) return PM_ERR_CUDA;
  pm_ctx* c = new pm_ctx();
  c->device = device;
  if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) {
    delete c;
    return PM_ERR_CUDA;
  }
  for (int b = 0; b < 4; ++b)
    for (int k = 0; k < 2; ++k)
      if (cudaEventCreate(&c->kev[b][k]) != cudaSuccess) { delete c; return PM_ERR_CUDA; }
  *out = c;
  return 0;
}

void pm_destroy(pm_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  // ranks may be destroyed at different times: unmap without the collective barrier
  for (void* p : c->ipc_open) cudaIpcCloseMemHandle(p);
  c->ipc_open.clear();
  if (g_const_owner[c->device & 63] == c) g_const_owner[c->device & 63] = nullptr;
  ncclComm_t comm = comm_of(c);
  c->comm = nullptr;
  state_free(c);
  graph_free(c);
  dev_free(c->rowstat);
  if (c->h_cnt) cudaFreeHost(c->h_cnt);
  if (c->h_rowstat) cudaFreeHost(c->h_rowstat);
  if (c->h_misc) cudaFreeHost(c->h_misc);
  if (c->scan_tmp) cudaFree(c->scan_tmp);
  dev_cache().flush(c->device);
  if (comm) ncclCommDestroy(comm);
  if (c->d_scratch) cudaFree(c->d_scratch);
  for (auto e : c->events) cudaEventDestroy(e);
  for (auto e : c->kev2) cudaEventDestroy(e);
  for (int b = 0; b < 4; ++b) for (int k = 0; k < 2; ++k) if (c->kev[b][k]) cudaEventDestroy(c->kev[b][k]);
  cudaStreamDestroy(c->stream);
  delete c;
}

const char* pm_last_error(const pm_ctx* c) { return c ? c->err.c_str() : "null context"; }
uint64_t pm_kernel_launches(const pm_ctx* c) { return c ? c->launches : 0; }

int pm_comm_unique_id(char id_out[PM_COMM_ID_BYTES]) {
  static_assert(sizeof(ncclUniqueId) <= PM_COMM_ID_BYTES, "PM_COMM_ID_BYTES too small");
  std::memset(id_out, 0, PM_COMM_ID_BYTES);
  if (!pm::dyn::api().ok) return PM_ERR_COMM;  // libnccl.so.2 not found
  ncclUniqueId id;
  if (ncclGetUniqueId(&id) != ncclSuccess) return PM_ERR_COMM;
  std::memcpy(id_out, &id, sizeof(id));
  return 0;
}

int pm_comm_init(pm_ctx* c, int rank, int n_ranks, const char* id_bytes) {
  if (!c || rank < 0 || n_ranks < 1 || rank >= n_ranks) return fail(c, PM_ERR_ARG, "pm_comm_init: bad rank");
  if (c->has_graph) return fail(c, PM_ERR_ARG, "pm_comm_init must precede the graph");
  if (n_ranks > PM_MAX_RANKS) return fail(c, PM_ERR_UNSUPPORTED, "at most 8 ranks (the GPUs of one NVSwitch box)");
  if (n_ranks == 1) { c->rank = 0; c->n_ranks = 1; return 0; }
  if (!id_bytes) return fail(c, PM_ERR_ARG, "pm_comm_init: null id");
  if (!pm::dyn::api().ok) return fail(c, PM_ERR_COMM, "libnccl.so.2 could not be loaded");
  PM_CUDA(c, cudaSetDevice(c->device));
  ncclUniqueId id;
  std::memcpy(&id, id_bytes, sizeof(id));
  ncclComm_t comm = nullptr;
  PM_NCCL(c, ncclCommInitRank(&comm, n_ranks, id, rank));
  c->comm = comm;
  c->rank = rank;
  c->n_ranks = n_ranks;
  PM_CUDA(c, cudaMalloc((void**)&c->d_scratch, 64));
  return 0;
}

// ------------------------------------------------------------------ graph
int pm_graph_from_slots(pm_ctx* c, uint64_t n_vertices, uint64_t n_slots, const uint32_t* src,
                        const uint32_t* dst) {
  if (!c || (n_slots && (!src || !dst))) return fail(c, PM_ERR_ARG, "pm_graph_from_slots: null argument");
  PM_CUDA(c, cudaSetDevice(c->device));
  state_free(c, /*keep_scratch=*/true);
  uint32_t *d_src = nullptr, *d_dst = nullptr;
  int rc;
  if ((rc = dev_alloc(c, &d_src, n_slots))) return rc;
  if ((rc = dev_alloc(c, &d_dst, n_slots))) { dev_free(d_src); return rc; }
  cudaError_t e1 = cudaMemcpyAsync(d_src, src, n_slots * 4, cudaMemcpyHostToDevice, c->stream);
  cudaError_t e2 = cudaMemcpyAsync(d_dst, dst, n_slots * 4, cudaMemcpyHostToDevice, c->stream);
  if (e1 != cudaSuccess || e2 != cudaSuccess) {
    dev_free(d_src); dev_free(d_dst);
    return fail(c, PM_ERR_CUDA, "pm_graph_from_slots: host to device copy failed");
  }
  rc = graph_build_from_device_slots(c, n_vertices, n_slots, d_src, d_dst, /*route=*/false);
  dev_free(d_src);
  dev_free(d_dst);
  return rc;
}

int pm_graph_from_csr(pm_ctx* c, uint64_t n_vertices, const uint64_t* rowptr, const uint32_t* col,
                      const uint64_t* degree_multi) {
  if (!c || !rowptr || !degree_multi) return fail(c, PM_ERR_ARG, "pm_graph_from_csr: null argument");
  PM_CUDA(c, cudaSetDevice(c->device));
  state_free(c, /*keep_scratch=*/true);
  // several ranks: rowptr / col / degree_multi describe the rows of the vertices THIS rank owns
  // (local row i = vertex i * n_ranks + rank, neighbours as global vertex ids)
  c->V = n_vertices;
  const uint64_t own = n_owned(c);
  if (rowptr[own] && !col) return fail(c, PM_ERR_ARG, "pm_graph_from_csr: null argument");
  return graph_build_from_host_csr(c, n_vertices, own, rowptr, col, degree_multi);
}

int pm_get_kernel_stats(const pm_ctx* c, int bin, pm_kernel_stats_t* out) {
  if (!c || !out || bin < 0 || bin > 4) return PM_ERR_ARG;
  *out = c->kstat[bin];
  return 0;
}

int pm_graph_rmat(pm_ctx* c, uint64_t scale, uint64_t gen_ranks) {
  if (!c) return PM_ERR_ARG;
  PM_CUDA(c, cudaSetDevice(c->device));
  state_free(c, /*keep_scratch=*/true);
  return rmat_build(c, scale, gen_ranks);
}

// Delegates: vertices whose multigraph out-degree reaches the threshold (generate_rmat / ingest_edge_list -d).  The hub
// rows stay with their home rank v mod G on the device; what follows the reference is the ATTRIBUTION: the count
// files and the vertex / edge / subgraph rows of a hub belong to its controller rank delegate_id % G.  Collective.
__global__ void k_find_hubs(const uint32_t* __restrict__ degm, uint64_t n_own, unsigned long long threshold, uint32_t G,
                            uint32_t rank, uint32_t* __restrict__ out, uint32_t cap, uint32_t* __restrict__ n_out) {
  uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (; i < n_own; i += stride)
    if ((unsigned long long)degm[i] >= threshold) {
      const uint32_t p = atomicAdd(n_out, 1u);
      if (p < cap) out[p] = (uint32_t)(i * G + rank);
    }
}
__global__ void k_mark_hubs(const uint32_t* __restrict__ local_idx, const uint8_t* __restrict__ ctl, uint32_t n, uint8_t* __restrict__ hub_ctl) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) hub_ctl[local_idx[i]] = ctl[i];
}

int pm_graph_set_delegate_threshold(pm_ctx* c, uint64_t threshold) {
  if (!c || !c->has_graph) return fail(c, PM_ERR_ARG, "pm_graph_set_delegate_threshold: no graph");
  PM_CUDA(c, cudaSetDevice(c->device));
  c->delegate_threshold = threshold;
  c->hubs.clear();
  c->state_ready = false;
  if (threshold == 0) return 0;
  const uint32_t cap = 1u << 20;  // hubs per rank this call can carry
  const int G = c->n_ranks;
  uint32_t *d_list = nullptr, *d_n = nullptr, *d_all = nullptr;
  int rc;
  if ((rc = dev_alloc(c, &d_list, (uint64_t)cap + 1)) || (rc = dev_alloc(c, &d_n, 1)) || (rc = dev_alloc(c, &d_all, ((uint64_t)cap + 1) * G))) {
    dev_free(d_list); dev_free(d_n); dev_free(d_all);
    return rc;
  }
  cudaMemsetAsync(d_n, 0, 4, c->stream);
  k_find_hubs<<<grid_for(), kBlock, 0, c->stream>>>(c->degm, n_owned(c), threshold, (uint32_t)G, (uint32_t)c->rank, d_list + 1, cap, d_n);
  c->launches++;
  cudaMemcpyAsync(d_list, d_n, 4, cudaMemcpyDeviceToDevice, c->stream);  // [0] = count, then the ids
  std::vector<uint32_t> all(((size_t)cap + 1) * G);
  cudaError_t e = cudaSuccess;
  ncclResult_t nr = ncclSuccess;
  if (G > 1) {
    nr = ncclAllGather(d_list, d_all, (size_t)cap + 1, ncclUint32, comm_of(c), c->stream);
    e = cudaMemcpyAsync(all.data(), d_all, all.size() * 4, cudaMemcpyDeviceToHost, c->stream);
  } else {
    e = cudaMemcpyAsync(all.data(), d_list, ((size_t)cap + 1) * 4, cudaMemcpyDeviceToHost, c->stream);
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
  dev_free(d_list); dev_free(d_n); dev_free(d_all);
  if (nr != ncclSuccess) return fail(c, PM_ERR_COMM, ncclGetErrorString(nr));
  if (e != cudaSuccess) return fail(c, PM_ERR_CUDA, cudaGetErrorString(e));
  for (int g = 0; g < G; ++g) {
    const uint32_t n = all[(size_t)g * (cap + 1)];
    if (n > cap) return fail(c, PM_ERR_CAPACITY, "delegate threshold leaves more than 2^20 hubs on one rank");
    c->hubs.insert(c->hubs.end(), all.begin() + (size_t)g * (cap + 1) + 1, all.begin() + (size_t)g * (cap + 1) + 1 + n);
  }
  std::sort(c->hubs.begin(), c->hubs.end());
  // controller + 1 of every hub this rank holds, by local vertex
  dev_free(c->hub_ctl);
  if ((rc = dev_alloc(c, &c->hub_ctl, c->nloc + 1))) return rc;
  PM_CUDA(c, cudaMemsetAsync(c->hub_ctl, 0, c->nloc + 1, c->stream));
  std::vector<uint32_t> idx;
  std::vector<uint8_t> ctl;
  for (size_t i = 0; i < c->hubs.size(); ++i)
    if ((int)(c->hubs[i] % (uint32_t)G) == c->rank) {
      idx.push_back(c->hubs[i] / (uint32_t)G);
      ctl.push_back((uint8_t)(i % (size_t)G + 1));
    }
  if (!idx.empty()) {
    uint32_t* d_idx = nullptr;
    uint8_t* d_ctl = nullptr;
    if ((rc = dev_alloc(c, &d_idx, idx.size())) || (rc = dev_alloc(c, &d_ctl, ctl.size()))) { dev_free(d_idx); return rc; }
    cudaMemcpyAsync(d_idx, idx.data(), idx.size() * 4, cudaMemcpyHostToDevice, c->stream);
    cudaMemcpyAsync(d_ctl, ctl.data(), ctl.size(), cudaMemcpyHostToDevice, c->stream);
    k_mark_hubs<<<grid_for(), kBlock, 0, c->stream>>>(d_idx, d_ctl, (uint32_t)idx.size(), c->hub_ctl);
    c->launches++;
    e = cudaStreamSynchronize(c->stream);
    dev_free(d_idx);
    dev_free(d_ctl);
    if (e != cudaSuccess) return fail(c, PM_ERR_CUDA, cudaGetErrorString(e));
  }
  return 0;
}

int pm_graph_num_delegates(const pm_ctx* c, uint64_t* n_out) {
  if (!c || !n_out) return PM_ERR_ARG;
  *n_out = c->hubs.size();
  return 0;
}

int pm_graph_info(const pm_ctx* c, pm_graph_info_t* o) {
  if (!c || !o || !c->has_graph) return PM_ERR_ARG;
  o->n_vertices = c->V; o->n_local = n_owned(c); o->n_slots_multi = c->E_multi; o->n_slots = c->E;
  o->n_slots_padded = c->Epad; o->max_degree = c->max_deg; o->device_bytes = c->graph_bytes;
  return 0;
}

int pm_graph_get_degree(const pm_ctx* cc, uint64_t* out) {
  pm_ctx* c = const_cast<pm_ctx*>(cc);
  if (!c || !c->has_graph || !out) return PM_ERR_ARG;
  const uint64_t own = n_owned(c);
  std::vector<uint32_t> h(own);
  PM_CUDA(c, cudaMemcpy(h.data(), c->degm, own * 4, cudaMemcpyDeviceToHost));
  for (uint64_t v = 0; v < own; ++v) out[v] = h[v];
  return 0;
}

int pm_graph_get_csr(const pm_ctx* cc, uint64_t* rowptr_out, uint32_t* col_out) {
  pm_ctx* c = const_cast<pm_ctx*>(cc);
  if (!c || !c->has_graph || !rowptr_out || !col_out) return PM_ERR_ARG;
  const uint64_t own = n_owned(c);
  std::vector<uint32_t> deg(own), blk(own + 1), col(c->Epad);
  PM_CUDA(c, cudaMemcpy(deg.data(), c->deg, own * 4, cudaMemcpyDeviceToHost));
  PM_CUDA(c, cudaMemcpy(blk.data(), c->rowblk, (own + 1) * 4, cudaMemcpyDeviceToHost));
  PM_CUDA(c, cudaMemcpy(col.data(), c->col0, c->Epad * 4, cudaMemcpyDeviceToHost));
  uint64_t o = 0;
  const uint32_t idmask = col_idmask(c);
  for (uint64_t v = 0; v < own; ++v) {
    rowptr_out[v] = o;
    std::memcpy(col_out + o, col.data() + (uint64_t)blk[v] * 8, (size_t)deg[v] * 4);
    if (idmask != 0xFFFFFFFFu)
      for (uint32_t j = 0; j < deg[v]; ++j) col_out[o + j] &= idmask;  // packed labels ride in the high bits
    if (c->n_ranks > 1) {  // slots -> vertex ids, ascending
      for (uint32_t j = 0; j < deg[v]; ++j) col_out[o + j] = (uint32_t)vertex_of(c, col_out[o + j]);
      std::sort(col_out + o, col_out + o + deg[v]);
    }
    o += deg[v];
  }
  rowptr_out[own] = o;
  return 0;
}

// ------------------------------------------------------------------ labels
int pm_labels_degree_log2(pm_ctx* c) {
  if (!c || !c->has_graph) return fail(c, PM_ERR_ARG, "pm_labels_degree_log2: no graph");
  PM_CUDA(c, cudaSetDevice(c->device));
  if (!c->label) { int rc = dev_alloc(c, &c->label, c->nloc, &c->graph_bytes); if (rc) return rc; }
  // the owner of v holds all of v's slots, so its degree (and label) is local (no delegate reduce needed)
  k_labels_degree_log2<<<grid_for(), kBlock, 0, c->stream>>>(c->degm, c->nloc, c->label);
  PM_LAUNCH_CHECK(c);
  c->has_labels = true;
  c->state_ready = false;
  c->labels_version = 0;  // a function of the graph alone
  // bit lengths of 32-bit degrees are <= 32; the largest label of ANY rank sizes the packed label field
  uint64_t maxd = c->max_deg;
  { int rc = comm_allreduce_max_u64(c, &maxd); if (rc) return rc; }
  uint64_t max_label = 0;
  while (maxd >> max_label) ++max_label;
  return labels_derive(c, true, max_label);
}

int pm_labels_set(pm_ctx* c, const uint64_t* labels) {
  if (!c || !c->has_graph || !labels) return fail(c, PM_ERR_ARG, "pm_labels_set: no graph or null labels");
  PM_CUDA(c, cudaSetDevice(c->device));
  if (!c->label) { int rc = dev_alloc(c, &c->label, c->nloc, &c->graph_bytes); if (rc) return rc; }
  const uint64_t* src = labels;
  std::vector<uint64_t> mine;
  if (c->n_ranks > 1) {  // the labels of the rows this rank owns, in local order
    mine.assign(c->nloc, 0);
    const uint64_t own = n_owned(c);
    for (uint64_t i = 0; i < own; ++i) mine[i] = labels[i * c->n_ranks + c->rank];
    src = mine.data();
  }
  PM_CUDA(c, cudaMemcpyAsync(c->label, src, c->nloc * 8, cudaMemcpyHostToDevice, c->stream));
  PM_CUDA(c, cudaStreamSynchronize(c->stream));
  c->has_labels = true;
  c->state_ready = false;
  c->labels_version = ++c->labels_counter;
  bool small = true;
  uint64_t max_label = 0;
  for (uint64_t v = 0; v < c->V && small; ++v) {
    small = labels[v] < 64;
    max_label = std::max<uint64_t>(max_label, labels[v]);
  }
  return labels_derive(c, small, max_label);
}

int pm_labels_from_files(pm_ctx* c, const char* base) {
  if (!c || !c->has_graph || !base) return fail(c, PM_ERR_ARG, "pm_labels_from_files: no graph or null base");
  std::vector<uint64_t> labels(c->V, 0);  // VertexData is value-initialised (beta.cpp:344-349)
  std::string err;
  if (!io::read_vertex_data(base, c->V, labels.data(), nullptr, err)) return fail(c, PM_ERR_IO, err);
  return pm_labels_set(c, labels.data());
}

static void copy_err(const std::string& why, char* err_out, size_t err_cap) {
  if (!err_out || !err_cap) return;
  std::strncpy(err_out, why.c_str(), err_cap - 1);
  err_out[err_cap - 1] = 0;
}

int pm_io_read_vertex_data(const char* base, uint64_t n_vertices, uint64_t* labels_inout, uint64_t* n_pairs_out,
                           char* err_out, size_t err_cap) {
  if (err_out && err_cap) err_out[0] = 0;
  if (!base || !labels_inout) return PM_ERR_ARG;
  std::string err;
  if (!io::read_vertex_data(base, n_vertices, labels_inout, n_pairs_out, err)) { copy_err(err, err_out, err_cap); return PM_ERR_IO; }
  return 0;
}

int pm_io_check_edge_data(const char* base, uint64_t n_vertices, uint64_t* n_records_out, char* err_out, size_t err_cap) {
  if (err_out && err_cap) err_out[0] = 0;
  if (!base) return PM_ERR_ARG;
  std::string err;
  if (!io::check_edge_data(base, n_vertices, n_records_out, err)) { copy_err(err, err_out, err_cap); return PM_ERR_IO; }
  return 0;
}

int pm_io_read_edge_lists(const char* const* files, int n_files, int undirected, uint64_t* n_vertices_out,
                          uint64_t* n_slots_out, uint32_t* src_out, uint32_t* dst_out, char* err_out, size_t err_cap) {
  if (err_out && err_cap) err_out[0] = 0;
  if (!files || n_files < 0) return PM_ERR_ARG;
  std::vector<std::string> fs;
  for (int i = 0; i < n_files; ++i) fs.push_back(files[i] ? files[i] : "");
  std::vector<uint32_t> src, dst;
  uint64_t nv = 0;
  std::string err;
  if (!io::read_edge_lists(fs, undirected != 0, src, dst, nv, err)) { copy_err(err, err_out, err_cap); return PM_ERR_IO; }
  if (n_vertices_out) *n_vertices_out = nv;
  if (n_slots_out) *n_slots_out = src.size();
  if (src_out && dst_out) {
    std::copy(src.begin(), src.end(), src_out);
    std::copy(dst.begin(), dst.end(), dst_out);
  }
  return 0;
}

int pm_labels_get(const pm_ctx* cc, uint64_t* out) {
  pm_ctx* c = const_cast<pm_ctx*>(cc);
  if (!c || !c->has_labels || !out) return PM_ERR_ARG;
  PM_CUDA(c, cudaStreamSynchronize(c->stream));
  if (c->n_ranks == 1) {
    PM_CUDA(c, cudaMemcpy(out, c->label, c->V * 8, cudaMemcpyDeviceToHost));
    return 0;
  }
  // collective: gather every rank's local labels, then slot order -> vertex order
  const uint64_t Vs = c->nlmax * c->n_ranks;
  unsigned long long* d = nullptr;
  int rc = dev_alloc(c, &d, Vs);
  if (rc) return rc;
  ncclResult_t nr = ncclAllGather(c->label, d, c->nlmax, ncclUint64, comm_of(c), c->stream);
  std::vector<uint64_t> h(Vs);
  cudaError_t e = cudaMemcpyAsync(h.data(), d, Vs * 8, cudaMemcpyDeviceToHost, c->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
  dev_free(d);
  if (nr != ncclSuccess) return fail(c, PM_ERR_COMM, ncclGetErrorString(nr));
  if (e != cudaSuccess) return fail(c, PM_ERR_CUDA, cudaGetErrorString(e));
  for (uint64_t v = 0; v < c->V; ++v) out[v] = h[slot_of(c, v)];
  return 0;
}

// ----------------------------------------------------------------- pattern
int pm_pattern_load_dir(pm_ctx* c, const char* dir) {
  if (!c || !dir) return PM_ERR_ARG;
  Pattern p;
  std::string why = load_pattern_dir(dir, p);
  if (!why.empty()) return fail(c, PM_ERR_PATTERN, why);
  PatConst pc;
  std::memset(&pc, 0, sizeof(pc));
  for (int i = 0; i < 16; ++i) {
    pc.N[i] = p.N[i];
    pc.No[i] = p.No[i];
    pc.min_opt[i] = (uint8_t)std::min(p.min_opt[i], 255);
    if (p.min_opt[i] > __builtin_popcount(p.No[i])) pc.never |= 1u << i;
  }
  pc.approx = p.approximate ? 1 : 0;
  // distinct template labels -> classes (ee.hpp:371-380 compares every template label)
  for (size_t i = 0; i < p.vertex_label.size(); ++i) {
    int k = 0;
    while (k < pc.ncls && pc.clabel[k] != p.vertex_label[i]) ++k;
    if (k == pc.ncls) pc.clabel[pc.ncls++] = p.vertex_label[i];
    pc.LMc[k] |= (uint16_t)(1u << i);
  }
  for (int l = 0; l < 64; ++l) pc.cls_of_label[l] = PM_NOCLASS;
  for (int k = 0; k < pc.ncls; ++k)
    if (pc.clabel[k] < 64) pc.cls_of_label[pc.clabel[k]] = (uint8_t)k;
  // first-superstep tables over the neighbour-label signature (labels < 64): every neighbour u sends
  // labelmask(label[u]) (ee.hpp:519-561), a sender is valid iff its mask meets NB(T_v) (:673-722), and
  // template vertex p survives iff N(p) is covered by what was heard (:901-939)
  // (approximate pattern: the optional neighbours count as required where a minimum optional edge count is set,
  //  local_constraint_checking.hpp:1090-1099; valid senders come over mandatory and optional edges, :641-651)
  for (int i = 0; i < p.n_vertices && i < 16; ++i)
    for (int b = 0; b < 16; ++b) {
      const bool needed = ((pc.N[i] >> b) & 1u) || (pc.min_opt[i] && ((pc.No[i] >> b) & 1u));
      if (needed && b < (int)p.vertex_label.size() && p.vertex_label[b] < 64) pc.req[i] |= 1ull << p.vertex_label[b];
    }
  for (int k = 0; k < pc.ncls; ++k) {
    uint32_t nb = 0;
    for (int a = 0; a < 16; ++a) if ((pc.LMc[k] >> a) & 1u) nb |= (uint32_t)pc.N[a] | pc.No[a];
    for (int q = 0; q < pc.ncls; ++q)
      if ((pc.LMc[q] & nb) && pc.clabel[q] < 64) pc.rl[k] |= 1ull << pc.clabel[q];
  }
  // typed slot table: the (at most three) non-empty subsets of every class with one or two template vertices
  pc.typed = 1;
  for (int k = 0; k < pc.ncls; ++k) {
    const uint32_t lm = pc.LMc[k];
    if (__builtin_popcount(lm) > 2 || pc.clabel[k] >= 64) { pc.typed = 0; continue; }
    int n = 1;
    for (uint32_t sub = lm; sub; sub = (sub - 1) & lm) pc.tsub[k][n++] = (uint16_t)sub;
  }
  c->pat = p;
  c->pc = pc;
  c->has_pattern = true;
  c->state_ready = false;
  c->pat_dir = dir;
  c->pat_key.clear();  // the remembered NLCC sizes are looked up per (pattern, graph, path) when a search starts
  c->pool_seen.assign(p.constraints.size(), 0);
  c->keys_seen.assign(p.constraints.size(), 0);
  c->subgraphs.assign(p.constraints.size(), {});
  c->subgraph_width.assign(p.constraints.size(), 0);
  c->subgraph_count.assign(p.constraints.size(), 0);
  return 0;
}

int pm_pattern_info(const pm_ctx* c, pm_pattern_info_t* o) {
  if (!c || !o || !c->has_pattern) return PM_ERR_ARG;
  o->n_vertices = c->pat.n_vertices; o->n_edges = c->pat.n_edges; o->diameter = c->pat.diameter;
  o->n_constraints = (int)c->pat.constraints.size();
  return 0;
}

// ------------------------------------------------------------------- state
}  // extern "C"

namespace {
// need_colw: the caller walks the working adjacency (the run_fuzzy path does not)
int state_reset(pm_ctx* c, bool need_colw) {
  if (!c || !c->has_graph || !c->has_labels || !c->has_pattern)
    return fail(c, PM_ERR_ARG, "pm_state_reset needs a graph, labels and a pattern");
  PM_CUDA(c, cudaSetDevice(c->device));
  const uint64_t NL = c->nloc;                      // rank-local arrays
  const uint64_t Vs = c->nlmax * c->n_ranks;        // replicated arrays (indexed by slot)
  const uint64_t base = c->nlmax * c->rank;
  const bool multi = c->n_ranks > 1;
  int rc;
  if (!c->S) {
    if ((rc = dev_alloc(c, &c->S, Vs))) return rc;
    if ((rc = dev_alloc(c, &c->adeg, NL + 1))) return rc;
    if ((rc = dev_alloc(c, &c->rowc, NL + 1))) return rc;
    if ((rc = dev_alloc(c, &c->vid, Vs))) return rc;
    if ((rc = dev_alloc(c, &c->clsc, Vs))) return rc;
    if ((rc = dev_alloc(c, &c->fw, Vs / 16 + 2))) return rc;
    if ((rc = dev_alloc(c, &c->hubc, NL + 1))) return rc;
    if ((rc = dev_alloc(c, &c->tb, Vs / PM_TILE + 2))) return rc;
    if ((rc = dev_alloc(c, &c->fwx, Vs / 16 + 2))) return rc;
    if ((rc = dev_alloc(c, &c->cls, Vs))) return rc;
    if ((rc = dev_alloc(c, &c->ok, Vs))) return rc;
    if ((rc = dev_alloc(c, &c->src_list, NL))) return rc;
    // frontier entry lists: main rows, and rows above PM_MID_MAX slots (at most E / PM_MID_MAX of those)
    for (int b = 0; b < 2; ++b) {
      if ((rc = dev_alloc(c, &c->fr[b][0], NL))) return rc;
      if ((rc = dev_alloc(c, &c->fr[b][1], std::min<uint64_t>(NL, c->E / PM_MID_MAX + 1024)))) return rc;
    }
    if ((rc = dev_alloc(c, &c->cnt, 1))) return rc;
    if (!c->h_cnt) PM_CUDA(c, cudaMallocHost((void**)&c->h_cnt, sizeof(DevCounters)));
    PM_CUDA(c, cudaMemsetAsync(c->adeg, 0, (NL + 1) * sizeof(uint32_t), c->stream));
    if (!c->h_misc) PM_CUDA(c, cudaMallocHost((void**)&c->h_misc, 16 * sizeof(uint32_t)));
    if (multi) {
      c->dcap = c->nlmax;  // a rank publishes at most one change per owned vertex and step
      if ((rc = dev_alloc(c, &c->din[0], c->dcap * c->n_ranks))) return rc;
      if ((rc = dev_alloc(c, &c->din[1], c->dcap * c->n_ranks))) return rc;
      if ((rc = dev_alloc(c, &c->step_msg, 1 + c->n_ranks))) return rc;
      PM_CUDA(c, cudaMemsetAsync(c->step_msg, 0, (1 + c->n_ranks) * sizeof(StepMsg), c->stream));
      // step mailbox; sequence numbers start at 1, and every rank restarts from 0 together here
      if ((rc = dev_alloc(c, &c->sync_in, 2 * c->n_ranks))) return rc;
      PM_CUDA(c, cudaMemsetAsync(c->sync_in, 0, 2 * c->n_ranks * sizeof(StepMsg), c->stream));
      PM_CUDA(c, cudaStreamSynchronize(c->stream));
      c->step_seq = 0;
      c->step_nccl = getenv("PM_COMM_NCCL") != nullptr;
      PM_CUDA(c, cudaMallocHost((void**)&c->h_step, sizeof(StepMsg) * c->n_ranks));
    }
    // the peer table (G = 1: everything points at this GPU)
    if ((rc = comm_publish(c))) return rc;
  }
  const int nrow = c->pat.diameter + 1;
  if (!c->rowstat || nrow > c->rowstat_cap) {
    dev_free(c->rowstat);
    if (c->h_rowstat) cudaFreeHost(c->h_rowstat);
    c->h_rowstat = nullptr;
    c->rowstat_cap = std::max(nrow, 64);
    if ((rc = dev_alloc(c, &c->rowstat, c->rowstat_cap))) return rc;
    PM_CUDA(c, cudaMallocHost((void**)&c->h_rowstat, c->rowstat_cap * sizeof(RowStat)));
  }
  while ((int)c->events.size() < nrow + 1) {
    cudaEvent_t e;
    PM_CUDA(c, cudaEventCreate(&e));
    c->events.push_back(e);
  }
  while (c->kev2.size() < (size_t)nrow * 4) {
    cudaEvent_t e;
    PM_CUDA(c, cudaEventCreate(&e));
    c->kev2.push_back(e);
  }
  c->kev2_cls.assign(nrow, 1);
  c->kev2_big.assign(nrow, 0);
  if (g_const_owner[c->device & 63] && g_const_owner[c->device & 63] != c) {
    PM_CUDA(c, cudaDeviceSynchronize());  // another context of this device may still be running on the old tables
    if ((rc = comm_upload_peers(c))) return rc;
  }
  g_const_owner[c->device & 63] = c;
  PM_CUDA(c, cudaMemcpyToSymbolAsync(c_pat, &c->pc, sizeof(PatConst), 0, cudaMemcpyHostToDevice, c->stream));
  PM_CUDA(c, cudaMemsetAsync(c->cnt, 0, sizeof(DevCounters), c->stream));
  c->cur = 0;
  c->filter_done = false;
  PM_CUDA(c, cudaEventRecord(c->kev[3][0], c->stream));
  {
    const bool small = c->labels_small;
    const uint64_t tpr = (NL + PM_TILE - 1) / PM_TILE;  // tiles per rank (several ranks: NL is a multiple of the tile)
    const int grid = grid_for();
    uint32_t* fw_me = c->fw + base / 16;
    uint32_t* tb_me = c->tb + (uint64_t)c->rank * tpr;
    // pass 1: survivor bits + in-tile prefixes of the local vertices (labels < 64: candidates that pass the
    // signature filter of the first superstep; else every label-matching vertex), survivors per tile
    // a single superstep per LCC call: no second scan could drop the neighbours the filter removed, so
    // the filter stays off and every label-matching vertex gets a compact id
    const bool use_sig = small && c->pat.diameter >= 2;
    // one rank, packed labels: the slot -> compact id table also carries every survivor's T_state number (see
    // PatConst::tsub), and supersteps 0 and 1 of the first LCC call can run as ONE pass (k_lcc_first_fused)
    c->typed = use_sig && !multi && c->col_shift != 0 && c->pc.typed && !getenv("PM_NO_TYPED");
    c->fused01 = c->typed && need_colw && getenv("PM_FUSE") != nullptr;
    if (small)
      k_init_flags<true><<<grid, kBlock, 0, c->stream>>>(c->lab8 + base, nullptr, c->deg, c->sig, NL, nullptr, fw_me, tb_me, c->cnt, use_sig ? 1 : 0);
    else
      k_init_flags<false><<<grid, kBlock, 0, c->stream>>>(nullptr, c->label, c->deg, nullptr, NL, c->cls + base, fw_me, tb_me, c->cnt, 0);
    PM_LAUNCH_CHECK(c);
    c->filter_done = use_sig;
    // pass 2: tile prefix, number of survivors
    k_init_scan<<<1, 1024, 0, c->stream>>>(tb_me, tpr, c->cnt, 0);
    PM_LAUNCH_CHECK(c);
    // the compact id ranges of the ranks; every rank needs every rank's slot -> cid tables
    for (int g = 0; g <= PM_MAX_RANKS; ++g) c->cid_off[g] = 0;
    if (multi) {
      if (!small && (rc = comm_allgather_slots(c, c->cls))) return rc;  // classes of foreign neighbours (first scan)
      if ((rc = comm_allgather_seg(c, c->fw, c->nlmax / 16))) return rc;
      if ((rc = comm_allgather_seg(c, c->tb, tpr))) return rc;
      if ((rc = comm_step(c))) return rc;
      if ((rc = comm_step_fetch(c))) return rc;
      c->step_parity ^= 1;
      for (int g = 0; g < c->n_ranks; ++g) c->cid_off[g + 1] = c->cid_off[g] + c->h_step[g].n_c;
    } else {
      if ((rc = sync_counters(c))) return rc;
      c->cid_off[1] = c->h_cnt->n_c;
    }
    for (int g = c->n_ranks + 1; g <= PM_MAX_RANKS; ++g) c->cid_off[g] = c->cid_off[c->n_ranks];
    for (int g = 0; g <= PM_MAX_RANKS; ++g) c->peers.off[g] = c->cid_off[g];
    if ((rc = comm_upload_peers(c))) return rc;
    if (multi) {
      k_tile_offsets<<<grid, kBlock, 0, c->stream>>>(c->tb, (uint32_t)tpr);
      PM_LAUNCH_CHECK(c);
    }
    // pass 3: per-cid state of every survivor (all ranks), row starts and frontier entries of the local ones
    if (small)
      k_init_assign<true><<<grid, kBlock, 0, c->stream>>>(c->lab8, nullptr, c->deg, c->rowblk, c->fw, c->tb, multi ? Vs : NL,
                                                          (uint32_t)base, (uint32_t)(base + NL), c->S, c->clsc, c->vid, c->adeg, c->rowc, c->fwx,
                                                          use_sig ? c->sig : nullptr, c->fr[0][0], c->fr[0][1], c->cnt, 0,
                                                          c->typed ? 1 : 0, hubs_on(c) ? c->hub_ctl : nullptr,
                                                          hubs_on(c) ? c->hubc : nullptr);
    else
      k_init_assign<false><<<grid, kBlock, 0, c->stream>>>(nullptr, c->cls, c->deg, c->rowblk, c->fw, c->tb, multi ? Vs : NL,
                                                           (uint32_t)base, (uint32_t)(base + NL), c->S, c->clsc, c->vid, c->adeg, c->rowc, c->fwx,
                                                           nullptr, c->fr[0][0], c->fr[0][1], c->cnt, 0, 0, hubs_on(c) ? c->hub_ctl : nullptr,
                                                           hubs_on(c) ? c->hubc : nullptr);
    PM_LAUNCH_CHECK(c);
    // the DENSE working adjacency: the row of local compact id i starts at the exclusive prefix of the survivors'
    // degrees (in 32-byte sectors, so rows stay sector aligned) — a few hundred million slots instead of the
    // graph's billions, and contiguous for every scan after the first
    const uint64_t n_c_local = c->cid_off[c->rank + 1] - c->cid_off[c->rank];
    PM_CUDA(c, cudaMemsetAsync(c->rowc + n_c_local, 0, sizeof(uint32_t), c->stream));
    {
      const uint32_t* in = c->rowc;  // k_init_assign left every survivor's row length in sectors here: scanned in place
      size_t tb = 0;
      PM_CUDA(c, cub::DeviceScan::ExclusiveSum(nullptr, tb, in, c->rowc, (int64_t)(n_c_local + 1), c->stream));
      if (tb > c->scan_tmp_bytes) {
        if (c->scan_tmp) cudaFree(c->scan_tmp);
        c->scan_tmp = nullptr;
        c->scan_tmp_bytes = std::max<size_t>(tb * 2, 1 << 16);
        PM_CUDA(c, cudaMalloc(&c->scan_tmp, c->scan_tmp_bytes));
      }
      tb = c->scan_tmp_bytes;
      PM_CUDA(c, cub::DeviceScan::ExclusiveSum(c->scan_tmp, tb, in, c->rowc, (int64_t)(n_c_local + 1), c->stream));
      c->launches += 2;
      PM_CUDA(c, cudaMemcpyAsync(c->h_misc, c->rowc + n_c_local, sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    }
  }
  PM_CUDA(c, cudaEventRecord(c->kev[3][1], c->stream));
  {
    int rc2 = sync_counters(c);
    if (rc2) return rc2;
    if (need_colw) {
      uint64_t want = (uint64_t)c->h_misc[0] * 8 + 64;  // + slack: the scan kernels read whole uint4 windows
      if (multi) {  // peers map the working adjacency (remote edge flags): it grows on every rank together
        if ((rc2 = comm_allreduce_max_u64(c, &want))) return rc2;
      }
      if (want > c->colw_cap) {
        if (multi) comm_close_all(c);
        dev_free(c->colw);
        c->colw_cap = 0;
        const uint64_t cap = want + want / 8;
        if ((rc2 = dev_alloc(c, &c->colw, cap))) return rc2;
        c->colw_cap = cap;
        if (multi && (rc2 = comm_publish(c))) return rc2;
      }
    }
    for (int b = 0; b < 2; ++b) c->bin_live[b] = c->h_cnt->fr_n[0][b] != 0;
    if (c->bin_live[1]) c->fused01 = false;  // rows above PM_MID_MAX slots take the CTA-per-row kernels: unfused path
    c->init_ms = 0;
    c->init_candidates = c->h_cnt->filtered_init;
    if (c->filter_done) PM_CUDA(c, cudaEventElapsedTime(&c->init_ms, c->kev[3][0], c->kev[3][1]));
  }
  c->fuzzy_ids = false;
  load_nlcc_sizes(c, "beta");
  c->rows.clear();
  c->step_rows.clear();
  c->iter_seconds.clear();
  c->itr = 0;
  c->summary = pm_run_summary_t{};
  for (auto& s : c->subgraphs) s.clear();
  std::fill(c->subgraph_count.begin(), c->subgraph_count.end(), 0);
  c->state_ready = true;
  return 0;
}
}  // namespace

extern "C" {

int pm_state_reset(pm_ctx* c) { return state_reset(c, true); }

// --------------------------------------------------------------------- LCC
int pm_lcc(pm_ctx* c, int init_step, int* not_finished, pm_counts_t* counts_out) {
  if (!c || !c->state_ready) return fail(c, PM_ERR_ARG, "pm_lcc: call pm_state_reset first");
  PM_CUDA(c, cudaSetDevice(c->device));
  { int rc0 = claim_constants(c); if (rc0) return rc0; }
  cudaStream_t st = c->stream;
  const int D = c->pat.diameter;
  const int grid = grid_for();
  const bool sm0 = c->labels_small;      // first scan: neighbour labels are streamed next to the ids (lab0)
  const bool ts_known = c->filter_done;  // ... and the signature filter already settled every entry's T_state
  const bool packed = c->col_shift != 0; // ... inside the col0 slots themselves (else in the parallel byte array lab0)
  const double t0 = wall_s();
  const bool dbg_steps = getenv("PM_DEBUG_LCC") != nullptr && !c->fused01;  // per-superstep split of the row times on stderr
  PM_CUDA(c, cudaMemsetAsync(c->rowstat, 0, D * sizeof(RowStat), st));
  PM_CUDA(c, cudaMemsetAsync(&c->cnt->nf, 0, sizeof(uint32_t), st));
  for (int k = 0; k < D; ++k) {  // fixed superstep count (ee.hpp:1069)
    const bool first = init_step && k == 0;
    const bool xlate = init_step && k == 1;  // rows still hold slots: this scan renames them to compact ids
    PM_CUDA(c, cudaEventRecord(c->events[k], st));  // also the start of the main scan for the kernel-class timing
    LccArgs a = lcc_args(c, k);
    const int cur = c->cur, nxt = cur ^ 1;
    // kernel classes timed with CUDA events on this stream: 0 = first-superstep scan of the main list,
    // 1 = later scans of the main list, 2 = CTA-per-row scans, 4 = the renaming (XLATE) scan
    const int cls_main = first ? 0 : xlate ? 4 : 1;
    cudaEvent_t* ev = &c->kev2[(size_t)k * 4];
    const bool fused = init_step && c->fused01;  // supersteps 0 and 1 in one pass: k == 0 scans, k == 1 only commits
    if (fused && k == 0) {
      k_lcc_first_fused<<<148 * 6, kBlock, 0, st>>>(a, c->fr[cur][0], &c->cnt->fr_n[cur][0], c->rowstat + 1);
      PM_LAUNCH_CHECK(c);
      PM_CUDA(c, cudaEventRecord(ev[1], st));
      c->kev2_cls[k] = 0;
      c->kev2_big[k] = 0;
      continue;
    }
    if (fused && k == 1) {
      PM_CUDA(c, cudaEventRecord(ev[1], st));
    } else if (c->bin_live[0]) {
      uint4* l = c->fr[cur][0];
      const uint32_t* np = &c->cnt->fr_n[cur][0];
      if (first && ts_known && packed && !getenv("PM_GENERIC_FIRST")) k_lcc_first_packed<<<148 * 6, kBlock, 0, st>>>(a, l, np);
      else if (first && ts_known && packed) k_lcc_scan<true, 2, false, false><<<grid, kBlock, 0, st>>>(a, l, np, 0);
      else if (first && ts_known) k_lcc_scan<true, 1, false, false><<<grid, kBlock, 0, st>>>(a, l, np, 0);
      else if (first && sm0 && packed) k_lcc_scan<true, 2, false, true><<<grid, kBlock, 0, st>>>(a, l, np, 0);
      else if (first && sm0) k_lcc_scan<true, 1, false, true><<<grid, kBlock, 0, st>>>(a, l, np, 0);
      else if (first) k_lcc_scan<true, 0, false, true><<<grid, kBlock, 0, st>>>(a, l, np, 0);
      else if (xlate && c->typed && getenv("PM_XLATE8")) k_lcc_xlate8<<<148 * 6, kBlock, 0, st>>>(a, l, np);  // measured: no faster
      else if (xlate) k_lcc_scan<false, 0, true, true><<<grid, kBlock, 0, st>>>(a, l, np, 0);
      else k_lcc_scan<false, 0, false, true><<<grid, kBlock, 0, st>>>(a, l, np, 0);
      PM_LAUNCH_CHECK(c);
    }
    if (!(fused && k == 1)) PM_CUDA(c, cudaEventRecord(ev[1], st));
    if (c->bin_live[1] && !(fused && k == 1)) {
      uint4* l = c->fr[cur][1];
      const uint32_t* np = &c->cnt->fr_n[cur][1];
      if (first && ts_known && packed) k_lcc_scan_big<true, 2, false, false><<<148, 1024, 0, st>>>(a, l, np, 0);
      else if (first && ts_known) k_lcc_scan_big<true, 1, false, false><<<148, 1024, 0, st>>>(a, l, np, 0);
      else if (first && sm0 && packed) k_lcc_scan_big<true, 2, false, true><<<148, 1024, 0, st>>>(a, l, np, 0);
      else if (first && sm0) k_lcc_scan_big<true, 1, false, true><<<148, 1024, 0, st>>>(a, l, np, 0);
      else if (first) k_lcc_scan_big<true, 0, false, true><<<148, 1024, 0, st>>>(a, l, np, 0);
      else if (xlate) k_lcc_scan_big<false, 0, true, true><<<148, 1024, 0, st>>>(a, l, np, 0);
      else k_lcc_scan_big<false, 0, false, true><<<148, 1024, 0, st>>>(a, l, np, 0);
      PM_LAUNCH_CHECK(c);
    }
    c->kev2_big[k] = (c->bin_live[1] && !(fused && k == 1)) ? 1 : 0;
    if (c->kev2_big[k]) PM_CUDA(c, cudaEventRecord(ev[2], st));
    c->kev2_cls[k] = cls_main;
    k_lcc_commit<<<grid, kBlock, 0, st>>>(a, c->fr[cur][0], c->fr[cur][1], c->fr[nxt][0], c->fr[nxt][1], cur, nxt);
    PM_LAUNCH_CHECK(c);
    if (dbg_steps) PM_CUDA(c, cudaEventRecord(ev[3], st));
    c->cur = nxt;
    if (c->n_ranks > 1) {
      // the commit stored this rank's mask changes into every peer's delta inbox; the StepMsg
      // all-gather is the barrier after which the peers' changes can be applied to the local replica
      int rc2 = comm_step(c);
      if (rc2) return rc2;
      k_apply_deltas<<<grid, kBlock, 0, st>>>(c->S, c->step_msg + 1, c->step_parity);
      PM_LAUNCH_CHECK(c);
      c->step_parity ^= 1;
    }
  }
  if (init_step && D == 1) {
    // no second scan in this call: rename the rows of the vertices still in the map to compact ids now
    LccArgs a = lcc_args(c, 0);
    const int cur = c->cur;
    k_lcc_scan<false, 0, true, true><<<grid, kBlock, 0, st>>>(a, c->fr[cur][0], &c->cnt->fr_n[cur][0], 1);
    PM_LAUNCH_CHECK(c);
    k_lcc_scan_big<false, 0, true, true><<<148, 1024, 0, st>>>(a, c->fr[cur][1], &c->cnt->fr_n[cur][1], 1);
    PM_LAUNCH_CHECK(c);
  }
  PM_CUDA(c, cudaEventRecord(c->events[D], st));
  PM_CUDA(c, cudaMemcpyAsync(c->h_rowstat, c->rowstat, D * sizeof(RowStat), cudaMemcpyDeviceToHost, st));
  int rc = sync_counters(c);
  if (rc) return rc;
  bool removed = c->h_cnt->nf || (init_step && c->filter_done && c->h_cnt->nf_init);
  if (c->n_ranks > 1) {  // global_not_finished is reduced over the ranks (beta.cpp:609)
    if ((rc = comm_step_fetch(c))) return rc;
    for (int g = 0; g < c->n_ranks; ++g)
      removed = removed || c->h_step[g].nf || (init_step && c->filter_done && c->h_step[g].pad);
  }
  if (removed && not_finished) *not_finished = 1;
  {
    for (int b = 0; b < 2; ++b) c->bin_live[b] = c->h_cnt->fr_n[c->cur][b] != 0;
  }
  if (dbg_steps) {
    fprintf(stderr, "[pm] rank %d lcc itr %d init %.3f scan/commit/barrier+deltas ms:", c->rank, (int)c->itr, init_step ? c->init_ms : 0.f);
    for (int k = 0; k < D; ++k) {
      const cudaEvent_t* ev = &c->kev2[(size_t)k * 4];
      float a_ms = 0, b_ms = 0, c_ms = 0;
      cudaEventElapsedTime(&a_ms, c->events[k], c->kev2_big[k] ? ev[2] : ev[1]);
      cudaEventElapsedTime(&b_ms, c->kev2_big[k] ? ev[2] : ev[1], ev[3]);
      cudaEventElapsedTime(&c_ms, ev[3], c->events[k + 1]);
      fprintf(stderr, " %.3f/%.3f/%.3f", a_ms, b_ms, c_ms);
    }
    fprintf(stderr, "\n");
  }
  for (int k = 0; k < D; ++k) {
    float ms = 0;
    PM_CUDA(c, cudaEventElapsedTime(&ms, c->events[k], c->events[k + 1]));
    pm_row_t r;
    r.itr = c->itr; r.kind = 0; r.index = k;
    r.n_vertices = c->h_rowstat[k].nv; r.n_edges = c->h_rowstat[k].ne; r.seconds = ms * 1e-3;
    c->rows.push_back(r);
    if (counts_out) { counts_out[k].n_vertices = r.n_vertices; counts_out[k].n_edges = r.n_edges; counts_out[k].seconds = r.seconds; }
    c->summary.device_seconds += r.seconds;
    const RowStat& rs = c->h_rowstat[k];
    const uint64_t sc = rs.scanned[0] + rs.scanned[1] + rs.scanned[2];
    const uint64_t vs = rs.verts[0] + rs.verts[1] + rs.verts[2];
    c->summary.edges_processed += sc;
    {
      float kms = 0;
      const cudaEvent_t* ev = &c->kev2[(size_t)k * 4];
      PM_CUDA(c, cudaEventElapsedTime(&kms, c->events[k], ev[1]));
      pm_kernel_stats_t& ks = c->kstat[c->kev2_cls[k]];
      if (init_step && c->fused01 && k == 1) {  // walked inside the fused first scan: its slots, no launch of its own
        c->kstat[0].slots += rs.scanned[0];
        c->kstat[0].vertices += rs.verts[0];
      } else if (rs.verts[0]) { ks.launches++; ks.ms += kms; ks.slots += rs.scanned[0]; ks.vertices += rs.verts[0]; }
      kms = 0;
      if (c->kev2_big[k]) PM_CUDA(c, cudaEventElapsedTime(&kms, ev[1], ev[2]));
      if (rs.verts[2]) { c->kstat[2].launches++; c->kstat[2].ms += kms; c->kstat[2].slots += rs.scanned[2]; c->kstat[2].vertices += rs.verts[2]; }
    }
    if (k == 0 && init_step && c->labels_small) {
      // the fused init + signature filter is part of superstep 0
      const float kms = c->init_ms;
      c->rows[c->rows.size() - 1].seconds += kms * 1e-3;
      c->summary.device_seconds += kms * 1e-3;
      c->kstat[3].launches++;
      c->kstat[3].ms += kms;
      c->kstat[3].vertices += c->init_candidates;
      // signature filter: 8 B signature + 2 B mask + 4 B list entry per candidate
      c->summary.algorithmic_bytes += c->init_candidates * 14;
    }
    // SURVEY §8(d) byte model: 4 B column + 2 B neighbour mask per scanned slot;
    // 12 B per scanned vertex (row start, own masks read + written, |E_v|)
    // 6.25 B per scanned slot + 12.25 B per scanned vertex
    c->summary.algorithmic_bytes += sc * 25 / 4 + vs * 49 / 4;
  }
  if (hubs_on(c)) {  // the hubs peers hold for this rank's controller role (count files, ee.hpp:1131-1138)
    if ((rc = hub_rows_merge(c, 0, D, &c->rows[c->rows.size() - D]))) return rc;
    for (int k = 0; counts_out && k < D; ++k) {
      counts_out[k].n_vertices = c->rows[c->rows.size() - D + k].n_vertices;
      counts_out[k].n_edges = c->rows[c->rows.size() - D + k].n_edges;
    }
  }
  c->step_rows.push_back({c->itr, wall_s() - t0});
  return 0;
}

// -------------------------------------------------------------------- NLCC
int pm_nlcc(pm_ctx* c, int pl, int mode, int* pattern_found, int* token_source_deleted,
            pm_counts_t* counts_out) {
  if (!c || !c->state_ready) return fail(c, PM_ERR_ARG, "pm_nlcc: call pm_state_reset first");
  if (pl < 0 || pl >= (int)c->pat.constraints.size()) return fail(c, PM_ERR_ARG, "pm_nlcc: bad constraint index");
  PM_CUDA(c, cudaSetDevice(c->device));
  { int rc0 = claim_constants(c); if (rc0) return rc0; }
  cudaStream_t st = c->stream;
  const Constraint& k = c->pat.constraints[pl];
  const bool tds = mode == PM_NLCC_TDS;
  if (!tds && !nem1_order_independent(k))
    return fail(c, PM_ERR_UNSUPPORTED,
                "constraint " + std::to_string(pl) +
                    ": repeated interior labels make the reference's nem_1 result depend on message "
                    "arrival order (SURVEY A.6 #7); use template driven search for this constraint");
  const int n = (int)k.P.size();
  if (tds && (int)k.enum_idx.size() < n)
    return fail(c, PM_ERR_PATTERN, "pattern_non_local_constraint has no enumeration indices for this walk");
  NlcConst nc;
  std::memset(&nc, 0, sizeof(nc));
  nc.n = n; nc.C = (int)k.C; nc.valid_cycle = k.valid_cycle ? 1 : 0;
  for (int h = 0; h < n; ++h) {
    int cl = PM_NOCLASS;
    for (int q = 0; q < c->pc.ncls; ++q) if (c->pc.clabel[q] == k.P[h]) cl = q;
    nc.cls[h] = (uint8_t)cl;
    nc.I[h] = (uint8_t)k.I[h];
    nc.e[h] = (uint8_t)(tds ? std::min<uint32_t>(k.enum_idx[h], 255u) : h);
    nc.lab[h] = (uint8_t)(k.P[h] < 64 ? k.P[h] : 255);
  }
  const int grid = grid_for();
  const int D = c->pat.diameter;  // rowstat[D] is the TP row accumulator
  PM_CUDA(c, cudaEventRecord(c->events[0], st));
  PM_CUDA(c, cudaMemcpyToSymbolAsync(c_nlc, &nc, sizeof(NlcConst), 0, cudaMemcpyHostToDevice, st));
  int rc;
  const bool multi = c->n_ranks > 1;
  // Size the (vertex, source) set for what this constraint stored last time (or for the current
  // edge maps on its first run); an undersized table is detected and the constraint retried.
  if (c->pool_seen.size() != c->pat.constraints.size()) c->pool_seen.assign(c->pat.constraints.size(), 0);
  if (c->keys_seen.size() != c->pat.constraints.size()) c->keys_seen.assign(c->pat.constraints.size(), 0);
  const uint64_t ne_now = c->rows.empty() ? c->E : c->rows.back().n_edges;
  uint64_t want_pool = c->pool_seen[pl] ? c->pool_seen[pl] + c->pool_seen[pl] / 2 + 4096 : 2 * ne_now + 65536;
  want_pool = std::max<uint64_t>(want_pool, 1ull << 18);
  // level 0 holds one token per source: never smaller than the vertices still in the map (the sizes remembered
  // per pattern directory may come from a smaller graph)
  want_pool = std::max<uint64_t>(want_pool, (c->rows.empty() ? c->nloc : c->rows.back().n_vertices) + 4096);
  uint64_t want_keys = c->keys_seen[pl] ? c->keys_seen[pl] + c->keys_seen[pl] / 2 + 4096 : want_pool;
  // several ranks: growing the inboxes is collective, so every rank must ask for the same sizes.  They do without
  // asking one another: pool_seen / keys_seen hold maxima over ALL ranks (taken from the step messages below) and
  // the row counts are made global here.
  if (multi && !c->pool_seen[pl]) {
    uint64_t w[2] = {ne_now, c->rows.empty() ? c->nloc : c->rows.back().n_vertices};
    if ((rc = comm_allreduce_u64(c, w, 2, ncclMax))) return rc;
    want_pool = std::max<uint64_t>(std::max<uint64_t>(2 * w[0] + 65536, 1ull << 18), w[1] + 4096);
    want_keys = want_pool;
    want_pool = std::max<uint64_t>(want_pool, c->nlmax / 8 + 4096);  // the floor of the later calls: grow once, now
  } else if (multi) {
    want_pool = std::max<uint64_t>(c->pool_seen[pl] + c->pool_seen[pl] / 2 + 4096, 1ull << 18);
    want_pool = std::max<uint64_t>(want_pool, c->nlmax / 8 + 4096);  // level 0: sources of this rank, bounded without a reduction
    want_keys = c->keys_seen[pl] + c->keys_seen[pl] / 2 + 4096;
  }
  if ((rc = nlcc_reserve(c, want_pool, want_keys))) return rc;
  auto pick_table = [&]() {
    uint64_t use = 1;
    while (use < 2 * want_keys) use <<= 1;
    c->hset_use = std::min(use, c->hset_cap);
  };
  pick_table();
  uint2* d_matches = nullptr;
  uint64_t match_cap = 0;
  const int cur = c->cur;
  uint64_t n_matches = 0, hi = 0, fanout = 0;
  int found = 0;
  // cycle constraints close their last two hops by intersecting E_v with E_s (k_nem1_close_cycle);
  // that needs symmetric edge maps, which flags set outside LCC can break only while diameter < 2
  const bool close2 = !tds && k.valid_cycle && (int)k.C >= 2 && c->pat.diameter >= 2;
  // one rank: the (vertex, source) aggregation is skipped where it cannot change anything — hop 1 (the neighbours
  // of a source are distinct) and, up to hop 2, the level that feeds the closing kernel (idempotent effects)
  auto nem1_dedupe = [&](int hn) { return hn >= 2 && !(close2 && hn == (int)k.C - 1 && hn <= 2); };
  const bool routed_close = !getenv("PM_CLOSE_KEYS");  // PM_CLOSE_KEYS: the earlier key-broadcast formulation (several ranks)
  bool any_dedupe = multi && close2 && !routed_close;  // the broadcast closing keys live in the hash set
  for (int hn = 1; !tds && hn <= (int)k.C; ++hn) any_dedupe = any_dedupe || (nem1_dedupe(hn) && !(close2 && hn == (int)k.C));
  for (int attempt = 0;; ++attempt) {
    if (tds && c->keep_subgraphs && !multi) {
      match_cap = c->pool_cap;
      dev_free(d_matches);
      if ((rc = dev_alloc(c, &d_matches, match_cap))) return rc;
    }
    // zero found .. the end, keep the frontier counters and nf
    PM_CUDA(c, cudaMemsetAsync(&c->cnt->found, 0, sizeof(DevCounters) - offsetof(DevCounters, found), st));
    if (!tds && any_dedupe) PM_CUDA(c, cudaMemsetAsync(c->hset, 0xFF, c->hset_use * sizeof(unsigned long long), st));
    // several ranks: `ok` doubles as this GPU's "already acknowledged" cache for foreign sources
    if (multi) PM_CUDA(c, cudaMemsetAsync(c->ok, 0, std::max<uint64_t>(1, c->cid_off[c->n_ranks]), st));
    const bool dbg = getenv("PM_DEBUG_HOPS") != nullptr;
    std::vector<cudaEvent_t> dev;
    auto mark = [&]() { if (dbg) { cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, st); dev.push_back(e); } };
    mark();
    if (multi && close2 && !routed_close) {
      // every rank learns the qualifying (source, neighbour) pairs of the closing hop (k_close_keys_m)
      NlcArgs ka = nlc_args(c, nullptr, 0);
      k_close_keys_m<<<grid, kBlock, 0, st>>>(ka, c->fr[cur][0], c->fr[cur][1], cur, (int)k.C);
      PM_LAUNCH_CHECK(c);
      k_close_keys_count_m<<<1, 1, 0, st>>>(c->cnt);
      PM_LAUNCH_CHECK(c);
      if ((rc = comm_step(c))) return rc;
      c->step_parity ^= 1;
      mark();
      k_close_ingest_m<<<grid, kBlock, 0, st>>>(nlc_args(c, nullptr, 0));
      PM_LAUNCH_CHECK(c);
      mark();
    }
    NlcArgs a = nlc_args(c, d_matches, match_cap);
    k_nlcc_sources<<<grid, kBlock, 0, st>>>(a, c->fr[cur][0], c->fr[cur][1], cur, tds ? 1 : 0);
    PM_LAUNCH_CHECK(c);
    k_nlcc_begin<<<1, 1, 0, st>>>(c->cnt);
    PM_LAUNCH_CHECK(c);
    if (!multi) {
      for (int hn = 1; hn <= (int)k.C + 1; ++hn) {
        const bool fin = hn == (int)k.C + 1;
        const int lvl = hn - 1;
        if (close2 && hn == (int)k.C) {
          k_nem1_close_cycle<<<grid, kBlock, 0, st>>>(a, lvl, hn);
          PM_LAUNCH_CHECK(c);
          break;
        }
        if (tds) {
          if (fin) k_tds_expand<true><<<grid, kBlock, 0, st>>>(a, lvl, hn);
          else k_tds_expand<false><<<grid, kBlock, 0, st>>>(a, lvl, hn);
        } else if (fin) {
          if (k.valid_cycle) k_nem1_final_cycle<<<grid, kBlock, 0, st>>>(a, lvl, hn);
          else k_nem1_expand<true><<<grid, kBlock, 0, st>>>(a, lvl, hn, 0);
        } else {
          k_nem1_expand<false><<<grid, kBlock, 0, st>>>(a, lvl, hn, nem1_dedupe(hn) ? 1 : 0);
        }
        PM_LAUNCH_CHECK(c);
        if (!fin) {
          k_nlcc_close_level<<<1, 1, 0, st>>>(c->cnt, hn, c->pool_cap);
          PM_LAUNCH_CHECK(c);
        }
      }
      // post-processing of token_source_map (beta.cpp:956-1062) and the TP row (beta.cpp:1094-1120) follow
      // without a host round trip: k_nlcc_apply does nothing if the walk ran out of room (the host then
      // retries the constraint with larger buffers from unchanged state)
      k_nlcc_apply<<<grid, kBlock, 0, st>>>(c->S, c->ok, c->src_list, c->cnt, c->step_parity, c->pool_cap, match_cap);
      PM_LAUNCH_CHECK(c);
      PM_CUDA(c, cudaMemsetAsync(c->rowstat + D, 0, sizeof(RowStat), st));
      k_count_alive<<<grid, kBlock, 0, st>>>(lcc_args(c, D), c->fr[cur][0], c->fr[cur][1], cur);
      PM_LAUNCH_CHECK(c);
      PM_CUDA(c, cudaEventRecord(c->events[1], st));
      PM_CUDA(c, cudaMemcpyAsync(c->h_rowstat + D, c->rowstat + D, sizeof(RowStat), cudaMemcpyDeviceToHost, st));
      if ((rc = sync_counters(c))) { dev_free(d_matches); return rc; }
    } else {
      // one kernel + one StepMsg all-gather per hop: tokens are stored into the owners' inboxes
      // (parity a.par), the all-gather publishes the region fill counts and is the barrier
      if ((rc = comm_step(c))) return rc;  // level 0 (the sources) is in inbox `par`
      c->step_parity ^= 1;
      // no rank has a source: nothing to walk (every rank sees the same counts and skips together)
      if ((rc = comm_step_fetch(c))) return rc;
      uint64_t n_src_all = 0;
      for (int g = 0; g < c->n_ranks; ++g) n_src_all += c->h_step[g].out_n[g];
      const bool walk = n_src_all != 0;
      mark();
      for (int hn = 1; walk && hn <= (int)k.C + 1; ++hn) {
        const bool fin = hn == (int)k.C + 1;
        // the tokens picked up now were accepted at hop hn - 1: aggregation only where duplicates can occur
        const int first = (hn == 1 || tds || !nem1_dedupe(hn - 1)) ? 1 : 0;
        a = nlc_args(c, nullptr, 0);
        bool last = fin;
        if (tds) {
          if (fin) k_tds_hop_m<true><<<grid, kBlock, 0, st>>>(a, hn);
          else k_tds_hop_m<false><<<grid, kBlock, 0, st>>>(a, hn);
        } else if (close2 && hn == (int)k.C && routed_close) {
          // closing requests travel to the owners of the sources, who check them against their own rows
          k_nem1_hop_m<3><<<grid, kBlock, 0, st>>>(a, hn, first);
          PM_LAUNCH_CHECK(c);
          mark();
          if ((rc = comm_step(c))) return rc;
          mark();
          c->step_parity ^= 1;
          k_close_check_m<<<grid, kBlock, 0, st>>>(nlc_args(c, nullptr, 0));
          last = true;
        } else if (close2 && hn == (int)k.C) {
          k_nem1_hop_m<2><<<grid, kBlock, 0, st>>>(a, hn, first);
          last = true;
        } else if (fin) {
          k_nem1_hop_m<1><<<grid, kBlock, 0, st>>>(a, hn, first);
        } else {
          k_nem1_hop_m<0><<<grid, kBlock, 0, st>>>(a, hn, first);
        }
        PM_LAUNCH_CHECK(c);
        mark();
        if ((rc = comm_step(c))) return rc;  // after the last hop: the acknowledgements have landed
        mark();
        c->step_parity ^= 1;
        if (last) break;
      }
      if (walk) {
        if ((rc = sync_counters(c))) return rc;
      } else {
        // no rank has a source (every rank saw the same level-0 message): nothing was walked, nothing to read back
        c->h_cnt->overflow = c->h_cnt->found = 0;
        c->h_cnt->pool_n = c->h_cnt->matches = c->h_cnt->fanout = c->h_cnt->peak_out = c->h_cnt->ce_n = 0;
      }
      if (dbg) {
        fprintf(stderr, "[pm] rank %d pl=%d attempt %d tcap %llu hop kernel/barrier ms:", c->rank, pl, attempt, (unsigned long long)c->tcap);
        for (size_t i = 0; i + 1 < dev.size(); ++i) { float ms = 0; cudaEventElapsedTime(&ms, dev[i], dev[i + 1]); fprintf(stderr, " %.3f", ms); }
        fprintf(stderr, "\n");
        for (auto e : dev) cudaEventDestroy(e);
      }
      if (walk && (rc = comm_step_fetch(c))) return rc;
    }
    hi = c->h_cnt->pool_n;
    // the pool / hash set / an inbox region ran out, or more walks completed than the match list holds
    bool overflow = c->h_cnt->overflow || (!multi && c->h_cnt->pool_n > c->pool_cap) ||
                    (!multi && tds && c->keep_subgraphs && (c->h_cnt->matches > match_cap || c->h_cnt->match_drop));
    uint64_t matches_here = c->h_cnt->matches;
    if (multi) {
      matches_here = 0;
      for (int g = 0; g < c->n_ranks; ++g) {
        overflow = overflow || c->h_step[g].overflow;
        if (tds) matches_here += c->h_step[g].out_n[c->rank];  // completed walks routed to this rank
      }
    }
    if (!overflow) {
      n_matches = matches_here;
      // several ranks: what must fit is the fullest inbox region of any hop (tokens arrive undeduplicated)
      uint64_t peak = multi ? (uint64_t)c->h_cnt->peak_out : (uint64_t)c->h_cnt->pool_n;
      uint64_t keys = (uint64_t)c->h_cnt->pool_n + c->h_cnt->ce_n;
      if (multi)  // the same maxima on every rank (see the sizing above)
        for (int g = 0; g < c->n_ranks; ++g) {
          peak = std::max<uint64_t>(peak, c->h_step[g].peak);
          keys = std::max<uint64_t>(keys, c->h_step[g].accepted + c->h_step[g].ce_n);
        }
      c->pool_seen[pl] = std::max<uint64_t>(c->pool_seen[pl], peak);
      c->keys_seen[pl] = std::max<uint64_t>(c->keys_seen[pl], keys);
      c->pool_cache[c->pat_key] = c->pool_seen;
      c->keys_cache[c->pat_key] = c->keys_seen;
      break;
    }
    if (attempt >= 6) { dev_free(d_matches); return fail(c, PM_ERR_CAPACITY, "NLCC token pool exhausted"); }
    want_pool = std::max<uint64_t>(want_pool * 4, (uint64_t)c->h_cnt->matches + 1);
    want_keys *= 4;
    if ((rc = nlcc_reserve(c, want_pool, want_keys))) { dev_free(d_matches); return rc; }
    pick_table();
  }
  fanout = c->h_cnt->fanout;
  found = c->h_cnt->found;
  if (multi)
    for (int g = 0; g < c->n_ranks; ++g) found = found || c->h_step[g].found;  // beta.cpp:1136
  if (getenv("PM_DEBUG")) {
    fprintf(stderr, "[pm] rank %d nlcc pl=%d %s n_src=%u fanout=%llu pool_n=%llu table=%llu levels:", c->rank, pl,
            tds ? "tds" : "nem1", c->h_cnt->n_src, (unsigned long long)fanout, (unsigned long long)c->h_cnt->pool_n,
            (unsigned long long)c->hset_use);
    for (int h = 0; h <= (int)k.C + 1; ++h) fprintf(stderr, " %llu", (unsigned long long)c->h_cnt->lvl[h]);
    fprintf(stderr, "\n");
  }
  // enumerated subgraphs (the file is truncated per outer iteration, beta.cpp:713-717); several ranks:
  // the completed walks are in the inbox the last hop wrote (now `step_parity ^ 1`)
  if (tds) {
    c->subgraph_width[pl] = n;
    c->subgraph_count[pl] = n_matches;
    c->subgraphs[pl].clear();
    c->summary.path_count += n_matches;  // path_count is never reset (tds_batch_1.hpp:14,1243)
    if (n_matches && (c->keep_subgraphs || multi)) {
      uint32_t* d_rows = nullptr;
      if ((rc = dev_alloc(c, &d_rows, n_matches * n))) { dev_free(d_matches); return rc; }
      if (multi) k_tds_collect_m<<<grid, kBlock, 0, st>>>(nlc_args(c, nullptr, 0), n, d_rows);
      else k_tds_materialize<<<grid, kBlock, 0, st>>>(c->pool, d_matches, n_matches, n, c->vid, d_rows);
      c->launches++;
      c->subgraphs[pl].resize(n_matches * n);
      cudaError_t e = cudaMemcpyAsync(c->subgraphs[pl].data(), d_rows, n_matches * n * 4, cudaMemcpyDeviceToHost, st);
      if (e == cudaSuccess) e = cudaStreamSynchronize(st);
      dev_free(d_rows);
      if (e != cudaSuccess) { dev_free(d_matches); return fail(c, PM_ERR_CUDA, cudaGetErrorString(e)); }
      if (multi)
        for (auto& x : c->subgraphs[pl]) x = (uint32_t)vertex_of(c, x);  // slots -> vertex ids
    }
    if (hubs_on(c) && c->keep_subgraphs) {
      // the subgraph line is written by the rank that owns the walk's final vertex (tds_batch_1.hpp:684-693): a hub's controller
      if ((rc = hub_records_exchange(c, c->subgraphs[pl], n, n - 1))) { dev_free(d_matches); return rc; }
      c->subgraph_count[pl] = c->subgraphs[pl].size() / (size_t)n;
    }
  }
  dev_free(d_matches);
  if (multi) {
    // post-processing of token_source_map (beta.cpp:956-1062) and the TP row (beta.cpp:1094-1120)
    k_nlcc_apply<<<grid, kBlock, 0, st>>>(c->S, c->ok, c->src_list, c->cnt, c->step_parity, ~0ull, 0ull);
    PM_LAUNCH_CHECK(c);
    // the deactivations reach the peers' replicas (vertex_data all_min/max_reduce, beta.cpp:1020-1039)
    if ((rc = comm_step(c))) return rc;
    k_apply_deltas<<<grid, kBlock, 0, st>>>(c->S, c->step_msg + 1, c->step_parity);
    PM_LAUNCH_CHECK(c);
    c->step_parity ^= 1;
    PM_CUDA(c, cudaMemsetAsync(c->rowstat + D, 0, sizeof(RowStat), st));
    LccArgs la = lcc_args(c, D);
    k_count_alive<<<grid, kBlock, 0, st>>>(la, c->fr[cur][0], c->fr[cur][1], cur);
    PM_LAUNCH_CHECK(c);
    PM_CUDA(c, cudaEventRecord(c->events[1], st));
    PM_CUDA(c, cudaMemcpyAsync(c->h_rowstat + D, c->rowstat + D, sizeof(RowStat), cudaMemcpyDeviceToHost, st));
    if ((rc = sync_counters(c))) return rc;
  }
  int deleted = c->h_cnt->deleted ? 1 : 0;
  if (multi) {  // token_source_deleted is reduced over the ranks (beta.cpp:1149)
    if ((rc = comm_step_fetch(c))) return rc;
    for (int g = 0; g < c->n_ranks; ++g) deleted = deleted || c->h_step[g].deleted;
  }
  float ms = 0;
  PM_CUDA(c, cudaEventElapsedTime(&ms, c->events[0], c->events[1]));
  pm_row_t r;
  r.itr = c->itr; r.kind = 1; r.index = pl;
  r.n_vertices = c->h_rowstat[D].nv; r.n_edges = c->h_rowstat[D].ne; r.seconds = ms * 1e-3;
  if (hubs_on(c) && (rc = hub_rows_merge(c, D, 1, &r))) return rc;
  c->rows.push_back(r);
  if (counts_out) { counts_out->n_vertices = r.n_vertices; counts_out->n_edges = r.n_edges; counts_out->seconds = r.seconds; }
  if (pattern_found) *pattern_found = found;
  if (token_source_deleted) *token_source_deleted = deleted;
  c->summary.device_seconds += r.seconds;
  c->summary.edges_processed += fanout;
  c->summary.algorithmic_bytes += fanout * 6 + (hi) * 16;  // §8(d): 4+2 B per walked slot, 8 B in + 8 B out per token
  return 0;
}

}  // extern "C"

// ------------------------------------------------------------------ results
namespace pm {

// (vertex, T_arr) of every vertex still in the map (beta.cpp:1386-1394); vid: compact id -> slot
// (null on the run_fuzzy path, whose entries name vertices directly)
__global__ void k_emit_vertices(LccArgs a, const uint32_t* __restrict__ vid, const uint4* __restrict__ l0,
                                const uint4* __restrict__ l1, int cur, uint2* __restrict__ out,
                                unsigned long long* __restrict__ n_out) {
  const uint32_t c0 = a.cnt->fr_n[cur][0], c1 = a.cnt->fr_n[cur][1];
  const uint32_t total = c0 + c1;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const uint4 e = i < c0 ? l0[i] : l1[i - c0];
    if (e.y == PM_TOMB) continue;
    const uint32_t T = a.S[e.x];
    if (T) out[atomicAdd(n_out, 1ull)] = make_uint2(vid ? vid[e.x] : e.x, T);
  }
}

// (vertex, neighbour) for every key of vertex_active_edges_map[v], v in the map (beta.cpp:1397-1403)
__global__ void k_emit_edges(LccArgs a, const uint32_t* __restrict__ vid, const uint4* __restrict__ l0,
                             const uint4* __restrict__ l1, int cur, uint2* __restrict__ out,
                             unsigned long long* __restrict__ n_out) {
  const uint32_t c0 = a.cnt->fr_n[cur][0], c1 = a.cnt->fr_n[cur][1];
  const uint32_t total = c0 + c1;
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t i = warp; i < total; i += nwarps) {
    const uint4 e = i < c0 ? l0[i] : l1[i - c0];
    if (e.y == PM_TOMB || !a.S[e.x]) continue;
    const uint32_t d = e.z;
    const uint64_t row = (uint64_t)e.y * 8;
    unsigned long long base = 0;
    if (lane == 0) base = atomicAdd(n_out, (unsigned long long)d);
    base = __shfl_sync(0xffffffffu, base, 0);
    for (uint32_t j = lane; j < d; j += 32) {
      const uint32_t u = a.colw[row + j] & PM_IDMASK;
      out[base + j] = make_uint2(vid ? vid[e.x] : e.x, vid ? vid[u] : u);
    }
  }
}

}  // namespace pm

namespace {

int fetch_pairs(pm_ctx* c, bool edges, std::vector<uint2>& host) {
  PM_CUDA(c, cudaSetDevice(c->device));
  { int rc0 = claim_constants(c); if (rc0) return rc0; }
  cudaStream_t st = c->stream;
  // size from a fresh count
  const int D = c->pat.diameter;
  PM_CUDA(c, cudaMemsetAsync(c->rowstat + D, 0, sizeof(RowStat), st));
  LccArgs la = lcc_args(c, D);
  const int cur = c->cur;
  k_count_alive<<<grid_for(), kBlock, 0, st>>>(la, c->fr[cur][0], c->fr[cur][1], cur);
  PM_LAUNCH_CHECK(c);
  PM_CUDA(c, cudaMemcpyAsync(c->h_rowstat + D, c->rowstat + D, sizeof(RowStat), cudaMemcpyDeviceToHost, st));
  PM_CUDA(c, cudaStreamSynchronize(st));
  uint64_t n = edges ? c->h_rowstat[D].ne : c->h_rowstat[D].nv;
  for (int g = 0; g < PM_MAX_RANKS; ++g)  // the hubs this rank holds are counted apart (for their controllers)
    n += edges ? c->h_rowstat[D].hub_ne[g] : c->h_rowstat[D].hub_nv[g];
  host.resize(n);
  if (!n && !(hubs_on(c) && !c->fuzzy_ids)) return 0;
  uint2* d_out = nullptr;
  unsigned long long* d_n = nullptr;
  int rc;
  if ((rc = dev_alloc(c, &d_out, n))) return rc;
  if ((rc = dev_alloc(c, &d_n, 1))) { dev_free(d_out); return rc; }
  cudaMemsetAsync(d_n, 0, sizeof(unsigned long long), st);
  const uint32_t* vid = c->fuzzy_ids ? nullptr : c->vid;
  if (edges) k_emit_edges<<<grid_for(), kBlock, 0, st>>>(la, vid, c->fr[cur][0], c->fr[cur][1], cur, d_out, d_n);
  else k_emit_vertices<<<grid_for(), kBlock, 0, st>>>(la, vid, c->fr[cur][0], c->fr[cur][1], cur, d_out, d_n);
  c->launches++;
  cudaError_t e = cudaMemcpyAsync(host.data(), d_out, n * sizeof(uint2), cudaMemcpyDeviceToHost, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  dev_free(d_out);
  dev_free(d_n);
  if (e != cudaSuccess) return fail(c, PM_ERR_CUDA, cudaGetErrorString(e));
  if (c->n_ranks > 1)  // slots -> vertex ids
    for (auto& p : host) {
      p.x = (uint32_t)vertex_of(c, p.x);
      if (edges) p.y = (uint32_t)vertex_of(c, p.y);
    }
  if (hubs_on(c) && !c->fuzzy_ids) {
    // a hub's rows of the result files belong to its controller rank (beta.cpp:1386-1403 iterate the controller's maps)
    std::vector<uint32_t> flat(host.size() * 2);
    for (size_t i = 0; i < host.size(); ++i) { flat[2 * i] = host[i].x; flat[2 * i + 1] = host[i].y; }
    int rc2 = hub_records_exchange(c, flat, 2, 0);
    if (rc2) return rc2;
    host.resize(flat.size() / 2);
    for (size_t i = 0; i < host.size(); ++i) host[i] = make_uint2(flat[2 * i], flat[2 * i + 1]);
  }
  std::sort(host.begin(), host.end(), [](const uint2& x, const uint2& y) { return x.x != y.x ? x.x < y.x : x.y < y.y; });
  return 0;
}

std::string bitset16(uint32_t x) {  // std::bitset<16> operator<<, most significant bit first
  std::string s(16, '0');
  for (int i = 0; i < 16; ++i) if ((x >> i) & 1u) s[15 - i] = '1';
  return s;
}

}  // namespace

extern "C" {

// ---------------------------------------------------------------- the loop
int pm_run(pm_ctx* c, const pm_run_options_t* opt_in, pm_run_summary_t* out) {
  if (!c) return PM_ERR_ARG;
  pm_run_options_t opt{4, 0, 0, 1};
  if (opt_in) opt = *opt_in;
  c->keep_subgraphs = opt.keep_subgraphs != 0;
  int rc = pm_state_reset(c);  // beta.cpp:484-492
  if (rc) return rc;
  const int max_it = opt.max_iterations > 0 ? opt.max_iterations : 1000;
  const size_t ncons = c->pat.constraints.size();
  std::vector<pm_counts_t> counts(c->pat.diameter);
  bool init_step = true;
  int nf = 0;
  const double t_begin = wall_s();
  do {  // beta.cpp:544
    const double it0 = wall_s();
    nf = 0;                                                   // :546
    if ((rc = pm_lcc(c, init_step, &nf, counts.data()))) return rc;  // :577-583
    init_step = false;                                        // :602-604
    if (!opt.lcc_only) {
      if (c->itr == 0) nf = 1;                                // forced token passing, :686-688
      if (nf) {                                               // :695
        nf = 0;                                               // :697
        for (size_t pl = 0; pl < ncons; ++pl) {               // :710
          const int mode = (opt.tds_from_pl >= 0 && (int)pl >= opt.tds_from_pl) ? PM_NLCC_TDS : PM_NLCC_NEM1;  // :762-767
          int found = 0, deleted = 0;
          pm_counts_t tp;
          if ((rc = pm_nlcc(c, (int)pl, mode, &found, &deleted, &tp))) return rc;
          if (deleted) nf = 1;                                // :994-996
          if (deleted && c->pat.constraints[pl].interleave)   // :1163-1184
            if ((rc = pm_lcc(c, 0, &nf, counts.data()))) return rc;
        }
      }
    }
    c->iter_seconds.push_back(wall_s() - it0);                // :1337-1338
    c->itr++;                                                 // :1341
    if ((int)c->itr >= max_it && nf) { c->err = "iteration cap reached (SURVEY A.6 #4)"; break; }
  } while (nf);                                               // :1351
  c->summary.iterations = c->itr;
  c->summary.search_seconds = wall_s() - t_begin;
  c->summary.n_rows = c->rows.size();
  if (!c->rows.empty()) {
    c->summary.n_active_vertices = c->rows.back().n_vertices;
    c->summary.n_active_edges = c->rows.back().n_edges;
  }
  if (out) *out = c->summary;
  return 0;
}

// ------------------------------------------------------ the run_fuzzy loop
int pm_run_fuzzy(pm_ctx* c, const pm_run_options_t* opt_in, pm_run_summary_t* out) {
  if (!c) return PM_ERR_ARG;
  if (!c->has_graph || !c->has_labels || !c->has_pattern) return fail(c, PM_ERR_ARG, "pm_run_fuzzy needs a graph, labels and a pattern");
  if (!c->labels_small) return fail(c, PM_ERR_UNSUPPORTED, "pm_run_fuzzy: labels must be < 64");
  // a walk that repeats a template vertex at interior hops makes the source-keyed aggregation set of
  // token_passing_pattern_matching.hpp:104-109 depend on message order
  for (const Constraint& k : c->pat.constraints)
    for (size_t x = 1; x <= k.C && x < k.I.size(); ++x)
      for (size_t y = x + 1; y <= k.C && y < k.I.size(); ++y)
        if (k.I[x] == k.I[y]) return fail(c, PM_ERR_UNSUPPORTED, "token walk repeats a template vertex at interior hops (order dependent)");
  pm_run_options_t opt{-1, 0, 0, 0};
  if (opt_in) opt = *opt_in;
  int rc = state_reset(c, false);  // allocations, pattern constants, bookkeeping (beta.cpp:484-492 analogue)
  if (rc) return rc;
  c->fuzzy_ids = true;         // this path names vertices by slot (no compact ids)
  load_nlcc_sizes(c, "fuzzy");
  cudaStream_t st = c->stream;
  const int D = c->pat.diameter, grid = grid_for();
  const int max_it = opt.max_iterations > 0 ? opt.max_iterations : 1000;
  const bool multi = c->n_ranks > 1;
  // Several ranks (the reference runs this path under MPI like the other, bsp.hpp:591, 632): the mask array S is
  // replicated by slot; a vertex that leaves the map is published to the peers as a (slot, 0) delta with the same
  // step barrier as pm_lcc; tokens travel through the owners' inboxes (pm_nlcc_multi.cuh) keyed by slot.
  if (multi) {  // the compact id ranges are not used on this path: every owner test goes by slot
    for (int g = 0; g <= PM_MAX_RANKS; ++g) c->peers.off[g] = c->cid_off[g] = 0;
    if ((rc = comm_upload_peers(c))) return rc;
  }
  FzArgs a;
  if ((rc = ensure_lab0(c))) return rc;  // packed labels: the byte label stream of this path is built on demand
  a.rowblk = c->rowblk; a.deg = c->deg; a.col0 = c->col0; a.lab0 = c->lab0; a.lab8 = c->lab8; a.S = c->S; a.cnt = c->cnt;
  a.idmask = col_idmask(c);
  a.base = (uint32_t)(c->nlmax * c->rank);
  a.par = c->step_parity;
  a.row = c->rowstat;
  PM_CUDA(c, cudaMemsetAsync(c->cnt, 0, sizeof(DevCounters), st));
  c->cur = 0;
  bool init = true;
  int nf = 0;
  const double t_begin = wall_s();
  // barrier + exchange of what the last kernel published (mask deltas)
  auto step_and_apply = [&]() -> int {
    if (!multi) return 0;
    int r2 = comm_step(c);
    if (r2) return r2;
    k_apply_deltas<<<grid, kBlock, 0, st>>>(c->S, c->step_msg + 1, c->step_parity);
    PM_LAUNCH_CHECK(c);
    c->step_parity ^= 1;
    a.par = c->step_parity;
    return 0;
  };
  do {
    const double it0 = wall_s();
    // ---- label_propagation_pattern_matching_bsp (bsp.hpp:598-699): `diameter` supersteps
    PM_CUDA(c, cudaMemsetAsync(c->rowstat, 0, D * sizeof(RowStat), st));
    PM_CUDA(c, cudaMemsetAsync(&c->cnt->nf, 0, sizeof(uint32_t), st));
    for (int k = 0; k < D; ++k) {
      PM_CUDA(c, cudaEventRecord(c->events[k], st));
      a.row = c->rowstat + k;
      if (init && k == 0) {
        PM_CUDA(c, cudaMemsetAsync(&c->cnt->fr_n[0][0], 0, 8 * sizeof(uint32_t), st));
        k_fz_init<<<grid, kBlock, 0, st>>>(a, c->sig, c->nloc, c->fr[0][0], 0);
        PM_LAUNCH_CHECK(c);
        c->cur = 0;
        // every rank decided its own vertices: the replicas get the whole array once
        if (multi && (rc = comm_allgather_slots(c, c->S))) return rc;
      } else {
        const int cur = c->cur, nxt = cur ^ 1;
        PM_CUDA(c, cudaMemsetAsync(&c->cnt->fr_n[nxt][0], 0, 4 * sizeof(uint32_t), st));
        k_fz_scan<<<grid, kBlock, 0, st>>>(a, c->fr[cur][0], &c->cnt->fr_n[cur][0]);
        PM_LAUNCH_CHECK(c);
        k_fz_commit<<<grid, kBlock, 0, st>>>(a, c->fr[cur][0], c->fr[nxt][0], cur, nxt);
        PM_LAUNCH_CHECK(c);
        c->cur = nxt;
        if ((rc = step_and_apply())) return rc;
      }
    }
    PM_CUDA(c, cudaEventRecord(c->events[D], st));
    if (multi && (rc = comm_step(c))) return rc;  // global_not_finished is reduced over the ranks (run_pattern_matching.cpp:497-505)
    if (multi) { c->step_parity ^= 1; a.par = c->step_parity; }
    PM_CUDA(c, cudaMemcpyAsync(c->h_rowstat, c->rowstat, D * sizeof(RowStat), cudaMemcpyDeviceToHost, st));
    if ((rc = sync_counters(c))) return rc;
    nf = c->h_cnt->nf ? 1 : 0;
    if (multi) {
      if ((rc = comm_step_fetch(c))) return rc;
      for (int g = 0; g < c->n_ranks; ++g) nf = nf || c->h_step[g].nf;
    }
    for (int k = 0; k < D; ++k) {
      float ms = 0;
      PM_CUDA(c, cudaEventElapsedTime(&ms, c->events[k], c->events[k + 1]));
      pm_row_t r;
      r.itr = c->itr; r.kind = 0; r.index = k; r.n_vertices = c->h_rowstat[k].nv; r.n_edges = 0; r.seconds = ms * 1e-3;
      c->rows.push_back(r);
      c->summary.device_seconds += r.seconds;
      c->summary.edges_processed += c->h_rowstat[k].scanned[0];
      // 4 B id + 1 B label per walked slot, 2 B mask gathers for the label-matching ones are not counted
      c->summary.algorithmic_bytes += c->h_rowstat[k].scanned[0] * 5 + c->h_rowstat[k].verts[0] * 16;
    }
    c->step_rows.push_back({c->itr, wall_s() - it0});
    init = false;
    // ---- token passing only after an LCC call that removed something (run_pattern_matching.cpp:511)
    if (nf) {
      nf = 0;
      PM_CUDA(c, cudaEventRecord(c->events[0], st));
      const int cur = c->cur;
      for (size_t pl = 0; pl < c->pat.constraints.size(); ++pl) {
        const Constraint& k = c->pat.constraints[pl];
        FzTok ft;
        std::memset(&ft, 0, sizeof(ft));
        ft.C = (int)k.C;
        ft.valid_cycle = k.valid_cycle ? 1 : 0;
        if (k.P.size() > 18) return fail(c, PM_ERR_UNSUPPORTED, "token walk longer than 18 vertices");
        for (size_t h = 0; h < k.P.size(); ++h) { ft.lab[h] = (uint8_t)(k.P[h] < 64 ? k.P[h] : 255); ft.I[h] = (uint8_t)k.I[h]; }
        PM_CUDA(c, cudaMemcpyToSymbolAsync(c_fz, &ft, sizeof(ft), 0, cudaMemcpyHostToDevice, st));
        if (c->pool_seen.size() != c->pat.constraints.size()) c->pool_seen.assign(c->pat.constraints.size(), 0);
        // the vertices still in the map bound the sources; several ranks agree on one size (growing the inboxes is collective)
        uint64_t alive_now = c->rows.back().n_vertices;
        if (multi && (rc = comm_allreduce_max_u64(c, &alive_now))) return rc;
        uint64_t want = c->pool_seen[pl] ? c->pool_seen[pl] + c->pool_seen[pl] / 2 + 4096 : std::max<uint64_t>(1ull << 20, 16 * alive_now);
        want = std::max<uint64_t>(want, alive_now + 4096);  // level 0: one token per source
        for (int attempt = 0;; ++attempt) {
          if ((rc = nlcc_reserve(c, want, want))) return rc;
          c->hset_use = c->hset_cap;
          uint64_t use = 1;
          while (use < 2 * want) use <<= 1;
          c->hset_use = std::min(use, c->hset_cap);
          PM_CUDA(c, cudaMemsetAsync(&c->cnt->found, 0, sizeof(DevCounters) - offsetof(DevCounters, found), st));
          PM_CUDA(c, cudaMemsetAsync(c->hset, 0xFF, c->hset_use * sizeof(unsigned long long), st));
          // several ranks: `ok` doubles as this GPU's "already acknowledged" cache for foreign sources
          if (multi) PM_CUDA(c, cudaMemsetAsync(c->ok, 0, c->nlmax * c->n_ranks, st));
          a.par = c->step_parity;
          NlcArgs t = nlc_args(c, nullptr, 0);
          k_fz_sources<<<grid, kBlock, 0, st>>>(a, c->fr[cur][0], cur, c->ok, c->src_list, c->pool, c->pool_cap);
          PM_LAUNCH_CHECK(c);
          k_nlcc_begin<<<1, 1, 0, st>>>(c->cnt);
          PM_LAUNCH_CHECK(c);
          if (!multi) {
            for (int hn = 1; hn <= (int)k.C; ++hn) {  // interior hops
              k_fz_expand<<<grid, kBlock, 0, st>>>(a, t, hn - 1, hn);
              PM_LAUNCH_CHECK(c);
              k_nlcc_close_level<<<1, 1, 0, st>>>(c->cnt, hn, c->pool_cap);
              PM_LAUNCH_CHECK(c);
            }
            if (k.valid_cycle) {  // a path never marks its source in this path (tp.hpp:263-268)
              k_fz_final<<<grid, kBlock, 0, st>>>(a, t, (int)k.C, (int)k.C + 1);
              PM_LAUNCH_CHECK(c);
            }
            if ((rc = sync_counters(c))) return rc;
          } else {
            // one kernel + one step barrier per hop (see pm_nlcc): level 0, the sources, sits in inbox `par`
            if ((rc = comm_step(c))) return rc;
            c->step_parity ^= 1;
            for (int hn = 1; hn <= (int)k.C; ++hn) {
              t = nlc_args(c, nullptr, 0);
              k_fz_expand_m<<<grid, kBlock, 0, st>>>(a, t, hn, hn == 1 ? 1 : 0);
              PM_LAUNCH_CHECK(c);
              if ((rc = comm_step(c))) return rc;
              c->step_parity ^= 1;
            }
            if (k.valid_cycle) {
              t = nlc_args(c, nullptr, 0);
              k_fz_final_m<<<grid, kBlock, 0, st>>>(a, t, (int)k.C + 1);
              PM_LAUNCH_CHECK(c);
              if ((rc = comm_step(c))) return rc;  // the acknowledgements have landed at the owners
              c->step_parity ^= 1;
            }
            a.par = c->step_parity;
            if ((rc = sync_counters(c))) return rc;
            if ((rc = comm_step_fetch(c))) return rc;
          }
          bool overflow = c->h_cnt->overflow || (!multi && c->h_cnt->pool_n > c->pool_cap);
          uint64_t peak = multi ? (uint64_t)c->h_cnt->peak_out : (uint64_t)c->h_cnt->pool_n;
          if (multi)
            for (int g = 0; g < c->n_ranks; ++g) {
              overflow = overflow || c->h_step[g].overflow;
              peak = std::max<uint64_t>(peak, std::max<uint64_t>(c->h_step[g].peak, c->h_step[g].accepted));
            }
          if (!overflow) {
            c->pool_seen[pl] = std::max<uint64_t>(c->pool_seen[pl], peak);
            c->pool_cache[c->pat_key] = c->pool_seen;
            break;
          }
          if (attempt >= 6) return fail(c, PM_ERR_CAPACITY, "token pool exhausted");
          want *= 4;
        }
        c->summary.edges_processed += c->h_cnt->fanout;
        c->summary.algorithmic_bytes += c->h_cnt->fanout * 5 + c->h_cnt->pool_n * 16;
        k_fz_apply<<<grid, kBlock, 0, st>>>(c->S, c->ok, c->src_list, c->cnt, c->step_parity);
        PM_LAUNCH_CHECK(c);
        if ((rc = step_and_apply())) return rc;
        if ((rc = sync_counters(c))) return rc;
        int deleted = c->h_cnt->deleted ? 1 : 0;
        if (multi) {
          if ((rc = comm_step_fetch(c))) return rc;
          for (int g = 0; g < c->n_ranks; ++g) deleted = deleted || c->h_step[g].deleted;
        }
        if (deleted) nf = 1;
      }
      // "itr, TP, 0, |map|" (run_pattern_matching.cpp:664-666)
      PM_CUDA(c, cudaMemsetAsync(c->rowstat + D, 0, sizeof(RowStat), st));
      a.row = c->rowstat + D;
      k_fz_count<<<grid, kBlock, 0, st>>>(a, c->fr[cur][0], cur);
      PM_LAUNCH_CHECK(c);
      PM_CUDA(c, cudaEventRecord(c->events[1], st));
      PM_CUDA(c, cudaMemcpyAsync(c->h_rowstat + D, c->rowstat + D, sizeof(RowStat), cudaMemcpyDeviceToHost, st));
      if ((rc = sync_counters(c))) return rc;
      float ms = 0;
      PM_CUDA(c, cudaEventElapsedTime(&ms, c->events[0], c->events[1]));
      pm_row_t r;
      r.itr = c->itr; r.kind = 1; r.index = 0; r.n_vertices = c->h_rowstat[D].nv; r.n_edges = 0; r.seconds = ms * 1e-3;
      c->rows.push_back(r);
      c->summary.device_seconds += r.seconds;
    }
    c->iter_seconds.push_back(wall_s() - it0);
    c->itr++;
    if ((int)c->itr >= max_it && nf) { c->err = "iteration cap reached"; break; }
  } while (nf);
  c->summary.iterations = c->itr;
  c->summary.search_seconds = wall_s() - t_begin;
  c->summary.n_rows = c->rows.size();
  c->summary.n_active_vertices = c->rows.empty() ? 0 : c->rows.back().n_vertices;
  c->summary.n_active_edges = 0;
  if (out) *out = c->summary;
  return 0;
}

int pm_end_iteration(pm_ctx* c, double seconds) {
  if (!c || !c->state_ready) return PM_ERR_ARG;
  c->iter_seconds.push_back(seconds);
  c->itr++;
  c->summary.iterations = c->itr;
  c->summary.search_seconds += seconds;
  c->summary.n_rows = c->rows.size();
  if (!c->rows.empty()) {
    c->summary.n_active_vertices = c->rows.back().n_vertices;
    c->summary.n_active_edges = c->rows.back().n_edges;
  }
  return 0;
}

int pm_pattern_constraint_info(const pm_ctx* c, int pl, pm_constraint_info_t* o) {
  if (!c || !o || !c->has_pattern || pl < 0 || pl >= (int)c->pat.constraints.size()) return PM_ERR_ARG;
  const Constraint& k = c->pat.constraints[pl];
  o->walk_length = (int)k.P.size();
  o->valid_cycle = k.valid_cycle;
  o->interleave_lcc = k.interleave;
  o->order_independent = nem1_order_independent(k);
  return 0;
}

int pm_pattern_check_dir(const char* dir, pm_pattern_info_t* info_out, pm_constraint_info_t* constraints_out,
                         int constraints_cap, char* err_out, size_t err_cap) {
  if (err_out && err_cap) err_out[0] = 0;
  if (!dir || !info_out) return PM_ERR_ARG;
  Pattern p;
  const std::string why = load_pattern_dir(dir, p);
  if (!why.empty()) {
    if (err_out && err_cap) {
      std::strncpy(err_out, why.c_str(), err_cap - 1);
      err_out[err_cap - 1] = 0;
    }
    return PM_ERR_PATTERN;
  }
  info_out->n_vertices = p.n_vertices; info_out->n_edges = p.n_edges; info_out->diameter = p.diameter;
  info_out->n_constraints = (int)p.constraints.size();
  for (int pl = 0; constraints_out && pl < constraints_cap && pl < (int)p.constraints.size(); ++pl) {
    const Constraint& k = p.constraints[pl];
    constraints_out[pl].walk_length = (int)k.P.size();
    constraints_out[pl].valid_cycle = k.valid_cycle;
    constraints_out[pl].interleave_lcc = k.interleave;
    constraints_out[pl].order_independent = nem1_order_independent(k);
  }
  return 0;
}

int pm_get_rows(const pm_ctx* c, pm_row_t* rows_out) {
  if (!c || !rows_out) return PM_ERR_ARG;
  std::copy(c->rows.begin(), c->rows.end(), rows_out);
  return 0;
}

int pm_get_active_vertices(const pm_ctx* cc, uint64_t* vertices_out, uint16_t* bits_out) {
  pm_ctx* c = const_cast<pm_ctx*>(cc);
  if (!c || !c->state_ready) return PM_ERR_ARG;
  std::vector<uint2> h;
  int rc = fetch_pairs(c, false, h);
  if (rc) return rc;
  for (size_t i = 0; i < h.size(); ++i) {
    if (vertices_out) vertices_out[i] = h[i].x;
    if (bits_out) bits_out[i] = (uint16_t)h[i].y;
  }
  c->summary.n_active_vertices = h.size();
  return 0;
}

int pm_get_active_edges(const pm_ctx* cc, uint64_t* pairs_out) {
  pm_ctx* c = const_cast<pm_ctx*>(cc);
  if (!c || !c->state_ready || !pairs_out) return PM_ERR_ARG;
  std::vector<uint2> h;
  int rc = fetch_pairs(c, true, h);
  if (rc) return rc;
  for (size_t i = 0; i < h.size(); ++i) { pairs_out[2 * i] = h[i].x; pairs_out[2 * i + 1] = h[i].y; }
  return 0;
}

int pm_get_subgraph_count(const pm_ctx* c, int pl, uint64_t* count_out, int* width_out) {
  if (!c || pl < 0 || pl >= (int)c->subgraph_count.size()) return PM_ERR_ARG;
  if (count_out) *count_out = c->subgraph_count[pl];
  if (width_out) *width_out = c->subgraph_width[pl];
  return 0;
}

int pm_get_subgraphs(const pm_ctx* c, int pl, uint32_t* rows_out) {
  if (!c || pl < 0 || pl >= (int)c->subgraphs.size() || !rows_out) return PM_ERR_ARG;
  if (c->subgraphs[pl].size() != c->subgraph_count[pl] * (uint64_t)c->subgraph_width[pl])
    return fail(const_cast<pm_ctx*>(c), PM_ERR_ARG, "subgraphs were not kept (keep_subgraphs = 0)");
  std::copy(c->subgraphs[pl].begin(), c->subgraphs[pl].end(), rows_out);
  return 0;
}

// result tree of beta.cpp:504-535, 713-717, 1375-1425 (row grammar: SURVEY A.5)
int pm_write_results(const pm_ctx* cc, const char* outdir) { return pm_write_results_ps(cc, outdir, 0); }

int pm_write_results_ps(const pm_ctx* cc, const char* outdir, int ps_index) {
  pm_ctx* c = const_cast<pm_ctx*>(cc);
  if (!c || !outdir || !c->state_ready || ps_index < 0) return PM_ERR_ARG;
  const std::string base(outdir), ps = base + "/" + std::to_string(ps_index);
  const std::string rk = std::to_string(c->rank);
  auto open = [&](const std::string& p, std::ofstream& f, bool append = false) -> bool {
    f.open(p, append ? std::ofstream::app : std::ofstream::out);
    if (!f) c->err = "cannot open " + p + " (the result tree must pre-exist, like the reference's)";
    return (bool)f;
  };
  if (c->rank == 0) {
    std::ofstream f_set, f_itr, f_step, f_ss;
    // one result_pattern_set per run (beta.cpp:413-414), one row per element of the set (:1375-1381)
    if (!open(base + "/result_pattern_set", f_set, ps_index > 0) || !open(ps + "/result_iteration", f_itr) ||
        !open(ps + "/result_step", f_step) || !open(ps + "/result_superstep", f_ss))
      return PM_ERR_IO;
    for (size_t i = 0; i < c->iter_seconds.size(); ++i) f_itr << i << ", " << c->iter_seconds[i] << "\n";
    for (auto& s : c->step_rows) f_step << s.first << ", LP, " << s.second << "\n";
    for (auto& r : c->rows) f_ss << r.itr << (r.kind == 0 ? ", LP, " : ", TP, ") << r.index << ", " << r.seconds << "\n";
    f_set << ps_index << ", " << c->n_ranks << ", " << c->summary.iterations << ", " << c->summary.search_seconds << ", "
          << c->pat.n_edges << ", " << c->pat.n_vertices << ", " << c->pat.constraints.size() << "\n";
  }
  std::ofstream fvc, fec, fv, fe, fm;
  if (!open(ps + "/all_ranks_active_vertices_count/active_vertices_" + rk, fvc) ||
      !open(ps + "/all_ranks_active_edges_count/active_edges_" + rk, fec) ||
      !open(ps + "/all_ranks_active_vertices/active_vertices_" + rk, fv) ||
      !open(ps + "/all_ranks_active_edges/active_edges_" + rk, fe) ||
      !open(ps + "/all_ranks_messages/messages_" + rk, fm))
    return PM_ERR_IO;
  for (auto& r : c->rows) {
    const char* kind = r.kind == 0 ? ", LP, " : ", TP, ";
    fvc << r.itr << kind << r.index << ", " << r.n_vertices << "\n";
    fec << r.itr << kind << r.index << ", " << r.n_edges << "\n";
    fm << r.itr << kind << r.index << ", 0\n";  // message counts are transport specific
  }
  std::vector<uint2> hv, he;
  int rc;
  if ((rc = fetch_pairs(c, false, hv))) return rc;
  if ((rc = fetch_pairs(c, true, he))) return rc;
  // labels of the vertices this rank owns (local row of v = v / n_ranks)
  std::vector<uint64_t> lab(c->nloc);
  PM_CUDA(c, cudaMemcpy(lab.data(), c->label, c->nloc * 8, cudaMemcpyDeviceToHost));
  // ... and of the hubs (their rows are written by the controller, which need not hold them): every rank contributes the
  // labels of the hubs it holds, a max-reduction hands everybody the whole table
  std::vector<uint64_t> hub_lab(c->hubs.size(), 0);
  if (hubs_on(c)) {
    for (size_t i = 0; i < c->hubs.size(); ++i)
      if ((int)(c->hubs[i] % (uint32_t)c->n_ranks) == c->rank) hub_lab[i] = lab[c->hubs[i] / (uint32_t)c->n_ranks];
    unsigned long long* d = nullptr;
    if ((rc = dev_alloc(c, &d, hub_lab.size() + 1))) return rc;
    cudaMemcpyAsync(d, hub_lab.data(), hub_lab.size() * 8, cudaMemcpyHostToDevice, c->stream);
    ncclResult_t nr = ncclAllReduce(d, d, hub_lab.size(), ncclUint64, ncclMax, comm_of(c), c->stream);
    cudaError_t e = cudaMemcpyAsync(hub_lab.data(), d, hub_lab.size() * 8, cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    dev_free(d);
    if (nr != ncclSuccess) return fail(c, PM_ERR_COMM, ncclGetErrorString(nr));
    if (e != cudaSuccess) return fail(c, PM_ERR_CUDA, cudaGetErrorString(e));
  }
  auto label_of = [&](uint32_t v) -> uint64_t {
    if (hubs_on(c)) {
      auto it = std::lower_bound(c->hubs.begin(), c->hubs.end(), v);
      if (it != c->hubs.end() && *it == v) return hub_lab[it - c->hubs.begin()];
    }
    return lab[v / c->n_ranks];
  };
  for (auto& p : hv) fv << c->rank << ", " << p.x << ", 0, " << label_of(p.x) << ", " << bitset16(p.y) << "\n";
  for (auto& p : he) fe << c->rank << ", " << p.x << ", " << p.y << "\n";
  for (size_t pl = 0; pl < c->pat.constraints.size(); ++pl) {
    std::ofstream fs;
    if (!open(ps + "/all_ranks_subgraphs/subgraphs_" + std::to_string(pl) + "_" + rk, fs)) return PM_ERR_IO;
    const int w = c->subgraph_width[pl];
    const auto& sg = c->subgraphs[pl];
    for (size_t i = 0; w && i + w <= sg.size(); i += w) {
      fs << "[" << c->rank << "], ";
      for (int x = 0; x < w; ++x) fs << sg[i + x] << ", ";
      fs << "[" << sg[i + w - 1] << "]\n";
    }
  }
  return 0;
}

}  // extern "C"
