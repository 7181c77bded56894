// pm_common.cuh — context, device-side constants and small helpers shared by the
// kernels of libpmgpu.so (sm_100a only; no CPU fallback, no multi-backend dispatch).
#pragma once
#include <chrono>
#include <unordered_map>
#include <mutex>
#include <list>

#include <cuda_runtime.h>
#include <stdint.h>

#include <map>
#include <string>
#include <vector>

#include "../../include/pmgpu.h"
#include "pm_pattern.hpp"

#define PM_SENTINEL 0xFFFFFFFFu  // padding slot in the adjacency arrays
#define PM_IDMASK 0x7FFFFFFFu    // bit 31 of a working-adjacency slot = "flag set outside LCC" 
#define PM_NOCLASS 16            // class id of a label no template vertex carries
#define PM_MAX_RANKS 8           // GPUs of one NVSwitch box

// degree bins (current active degree) -> kernel shape
#define PM_MID_MAX 4096u    // <= 4096 slots: one warp per vertex
                            // larger: one CTA per vertex

namespace pm {

// Template (pattern) constants, uploaded once per pattern.
struct PatConst {
  uint16_t N[16];      // N[p]: template neighbours of template vertex p over mandatory edges
  uint16_t No[16];     // ... over optional edges (approximate matching; all zero for an exact pattern)
  uint8_t min_opt[16]; // vertex_min_optional_edge_count (0: no requirement)
  int approx;          // approximate local constraint (approximate_pattern_matching/local_constraint_checking.hpp)
  uint32_t never;      // bit p: the minimum optional edge count of p exceeds its optional edges (p can never stay)
  // typed slot -> compact id table (one rank, every label class has at most two template vertices): the T_state a
  // survivor ends the first superstep with is one of at most three subsets of its class — tsub[class][1..3], [0] = 0
  // ("no survivor") — and its 2-bit number rides in the table next to the compact id, so the second superstep needs
  // ONE gather per kept neighbour (k_lcc_first_fused, k_lcc_scan<XLATE> typed) instead of a table and a mask gather
  uint16_t tsub[17][4];
  int typed;           // every class has at most two template vertices: the typed table can be used
  uint16_t LMc[17];    // class -> bitmask of template vertices carrying that label; [16] = 0
  uint64_t clabel[16]; // class -> label value
  int ncls;
  uint8_t cls_of_label[64];  // small-label mode: label value -> class (PM_NOCLASS if none)
  // small-label mode, first superstep from the neighbour-label signature sig[v]:
  unsigned long long req[16];  // template vertex p survives iff (sig & req[p]) == req[p]: req[p] = labels of N(p)
  unsigned long long rl[17];   // class c heard a valid neighbour iff sig & rl[c] != 0
};

struct NlcConst {      // one non-local constraint (walk)
  uint8_t cls[16];     // class of P[h]  (PM_NOCLASS if the label is not in the template)
  uint8_t I[16];       // template vertex id at hop h
  uint8_t e[16];       // enumeration index (TDS history rule)
  uint8_t lab[16];     // label value of P[h] when labels are bytes < 64, else 255
  int n;               // walk length = C + 2
  int C;               // max itr_count
  int valid_cycle;
};

struct DevCounters {
  uint32_t fr_n[2][4];          // frontier sizes [buffer][bin]; [.][3] unused
  unsigned long long filtered_init;  // candidates tested by the fused init filter
  uint32_t nf;                  // a vertex left the vertex_state_map in this LCC call
  uint32_t found;               // NLCC: a walk completed
  uint32_t deleted;             // NLCC: a source failed
  uint32_t overflow;            // NLCC: token pool / hash set exhausted
  uint32_t n_src;               // NLCC: number of sources
  uint32_t nf_init;             // the fused init filter removed a vertex that had entered the map
  uint32_t match_drop;          // TDS: a completed walk did not fit the match list
  uint32_t ticket;              // k_lcc_commit: blocks finished (the last one re-arms the frontier counters it consumed)
  unsigned long long pool_n;    // NLCC: tokens in the pool
  unsigned long long matches;   // TDS: completed walks
  unsigned long long fanout;    // NLCC: adjacency slots walked by tokens
  unsigned long long hash_n;    // NLCC: keys in the (vertex, source) set
  unsigned long long lvl[20];   // NLCC: token pool level bounds, level h = [lvl[h], lvl[h+1])
  // multi-GPU (all zero with one rank)
  unsigned long long out_n[PM_MAX_RANKS];  // tokens / walks this rank stored in the inbox region of rank g during the current hop
  uint32_t ndelta;              // mask changes this rank published in the current step
  uint32_t pad2;
  unsigned long long peak_out;  // largest inbox region fill of any hop (sizes the next run's inboxes)
  unsigned long long ce_n;      // closing-edge keys filed in the hash set
  uint32_t n_c;                 // survivors of the first-superstep filter on this rank = compact ids it owns
  uint32_t pad3;
};

// what every rank contributes to the per-step all-gather (= the barrier)
struct StepMsg {
  unsigned long long out_n[PM_MAX_RANKS];
  uint32_t ndelta, nf, found, deleted, overflow, pad;  // pad: nf_init
  unsigned long long accepted;   // tokens accepted so far (pool_n)
  unsigned long long peak;       // fullest inbox region this rank has written in the current constraint
  unsigned long long ce_n;       // closing-edge keys this rank filed
  uint32_t n_c, pad1;            // compact ids this rank owns (published once per pattern)
  uint32_t seq;                  // step number: written LAST, polled by the receiver
  uint32_t timeout;              // a peer never arrived (the step barrier gave up)
};

// Multi-GPU: one process per GPU, 1-D vertex partition owner(v) = v mod G like the reference
// (impl/delegate_partitioned_graph.ipp:1682-1697).  On the device every vertex is named by its
// SLOT, slot(v) = (v mod G) * nlmax + v / G, so that rank r owns the contiguous slot range
// [r * nlmax, (r + 1) * nlmax): adjacency arrays store slots, replicated per-vertex arrays (S, cls,
// lab8, ok) are indexed by slot, rank-local arrays by slot - base.  With G = 1, slot(v) = v.
// Peers exchange data by storing straight into each other's memory over NVLink (CUDA IPC mappings of
// cudaMalloc'd buffers); NCCL carries only the bootstrap, bulk all-gathers at initialisation and the
// small per-step all-gather that doubles as the barrier.
struct PeerTab {
  int G, rank;
  uint32_t nlmax;       // slots per rank
  uint32_t base;        // first slot of this rank
  uint32_t dcap;        // delta inbox: entries per sender region
  unsigned long long tcap;  // token inbox: tokens per sender region
  const uint32_t* rowblk[PM_MAX_RANKS];  // rank-local arrays of every rank (index: slot - owner * nlmax)
  const uint32_t* adeg[PM_MAX_RANKS];
  uint32_t* colw[PM_MAX_RANKS];
  uint8_t* ok[PM_MAX_RANKS];             // replicated-size array of every rank (index: slot); truth lives at the owner
  uint2* din[2][PM_MAX_RANKS];           // delta inbox of rank g: G regions of dcap (slot, mask) pairs, double buffered
  uint2* tin[2][PM_MAX_RANKS];           // token inbox of rank g: G regions of tcap tokens, double buffered
  StepMsg* sync_in[PM_MAX_RANKS];        // step mailbox of rank g: [2][G] StepMsg, slot [seq & 1][sender]
  // compact ids of the current pattern (see "Compact ids" in pm_lcc.cuh): rank g owns [off[g], off[g+1])
  uint32_t off[PM_MAX_RANKS + 1];
};

struct RowStat {                // one result row, accumulated on the device
  unsigned long long nv, ne;
  unsigned long long scanned[3];   // adjacency slots walked, per degree bin
  unsigned long long verts[3];     // live vertices whose row was walked, per degree bin
  unsigned long long filtered;     // candidates settled or passed on by the signature filter
  // several ranks with a delegate threshold: hubs this rank holds, counted for their CONTROLLER rank
  // (delegate_id % G, delegate_partitioned_graph.hpp:231-233) instead of into nv / ne
  unsigned long long hub_nv[PM_MAX_RANKS], hub_ne[PM_MAX_RANKS];
};

}  // namespace pm

struct pm_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  std::string err;
  uint64_t launches = 0;
  int rank = 0, n_ranks = 1;
  void* comm = nullptr;            // ncclComm_t
  uint64_t nlmax = 0;              // slots per rank (multiple of 16); slot base of this rank = rank * nlmax
  pm::PeerTab peers{};             // host copy of c_peer
  std::vector<void*> ipc_open;     // peer mappings currently open
  pm::StepMsg* step_msg = nullptr; // device: [1 + G] (mine, then everyone's)
  pm::StepMsg* sync_in = nullptr;  // device: [2][G] step mailbox the peers store into
  uint32_t step_seq = 0;           // steps taken so far (same on every rank)
  bool step_nccl = false;          // PM_COMM_NCCL=1: use an NCCL all-gather for the step barrier instead
  unsigned long long* d_scratch = nullptr;  // device: 8 words for small all-reduces
  pm::StepMsg* h_step = nullptr;   // pinned: [G]
  uint2* din[2] = {nullptr, nullptr};   // delta inboxes (G regions of dcap)
  uint2* tin[2] = {nullptr, nullptr};   // token inboxes (G regions of tcap)
  uint64_t dcap = 0, tcap = 0;
  int step_parity = 0;             // which delta inbox the next publication uses

  // ---- graph store (device) -------------------------------------------------
  uint64_t V = 0, nloc = 0, E_multi = 0, E = 0, Epad = 0, max_deg = 0, graph_bytes = 0;
  uint32_t* rowblk = nullptr;  // [V+1] row start in units of 8 slots (32-byte sectors)
  uint32_t* deg = nullptr;     // [V] distinct-neighbour degree
  uint32_t* degm = nullptr;    // [V] multigraph out-degree (duplicates + self loops) -> labels
  uint32_t* col0 = nullptr;    // [Epad] pristine adjacency: sorted, distinct, rows padded to 8 with PM_SENTINEL
  uint32_t* colw = nullptr;    // [colw_cap] DENSE working adjacency (per pattern): row of local compact id i starts at sector
                               // rowc[i] and has room for deg(v) slots; its first adeg[i] slots = keys(E_v)
  uint64_t colw_cap = 0;       // slots allocated (grow only)
  void* scan_tmp = nullptr;    // cub scratch of the row-start prefix
  size_t scan_tmp_bytes = 0;
  uint32_t* h_misc = nullptr;  // pinned: small read-backs ([0] = sectors the dense working adjacency needs)
  uint64_t* label = nullptr;   // [V] vertex labels
  uint8_t* lab8 = nullptr;     // [V] labels as bytes when every label is < 64 (degree labels always are)
  unsigned long long* sig = nullptr;  // [V] bit l set iff some distinct neighbour carries label l (labels < 64)
  uint8_t* lab0 = nullptr;     // [Epad] label of the neighbour stored in col0 (labels < 64); with packed labels only the
                               // run_fuzzy path asks for it (built on demand)
  uint32_t col_shift = 0;      // != 0: PACKED labels — a col0 slot is (label of the neighbour << col_shift) | neighbour id,
                               // chosen when id bits + label bits <= 32: the first scan streams ONE array
  bool labels_small = false;   // lab8 / sig are valid
  bool has_graph = false, has_labels = false;
  // delegates (pm_graph_set_delegate_threshold): vertices whose multigraph out-degree reaches the threshold, ascending
  // (their position is the delegate id, ipp:501-512, 681); hub_ctl[local vertex] = controller rank + 1, 0 for the rest
  uint64_t delegate_threshold = 0;
  std::vector<uint32_t> hubs;
  uint8_t* hub_ctl = nullptr;
  uint8_t* hubc = nullptr;     // [nloc] the same by local compact id (per pattern)

  // ---- pattern ----------------------------------------------------------------
  pm::Pattern pat;
  pm::PatConst pc;
  bool has_pattern = false;

  // ---- per-pattern state (device) -----------------------------------------------
  uint16_t* S = nullptr;    // [V] template_vertices[v] (T_arr) while v is active and in the map, else 0
                            // (vertex_state.template_vertices, T_state, lives in the frontier entries)
  uint32_t* adeg = nullptr; // [nloc + 1] |E_v|, by LOCAL compact id (cid - off[rank])
  uint32_t* rowc = nullptr; // [nloc + 1] row start (sectors) in colw, by local compact id
  uint32_t* vid = nullptr;  // [Vs] compact id -> slot (replicated)
  uint8_t* clsc = nullptr;  // [Vs] label class by compact id (replicated)
  uint32_t* fw = nullptr;   // [Vs / 16] per 16 slots: survivors before them in their tile << 16 | survivor bits (replicated)
  uint32_t* tb = nullptr;   // [tiles] compact id of every 4096-slot tile's first survivor (replicated)
  uint2* fwx = nullptr;     // [Vs / 16] {compact id of the word's first survivor, survivor bits}: what cid_of_slot reads
  bool typed = false;       // fwx[w].y holds 2-bit T_state numbers (0 = no survivor) instead of survivor bits
  bool fused01 = false;     // the current pattern's first LCC call runs supersteps 0 and 1 in one pass
  uint32_t cid_off[PM_MAX_RANKS + 1] = {0};  // host copy of the compact id ranges
  uint8_t* cls = nullptr;   // [V] label class (index into the template's distinct labels, PM_NOCLASS = none)
  uint4* fr[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};  // frontier entry lists [buffer][main, big rows]
  int cur = 0;              // which frontier buffer is current
  bool bin_live[2] = {true, true};  // main list / big-row list non-empty at the last host sync
  bool filter_done = false;     // the fused init filter ran: the first superstep skips its own filter
  float init_ms = 0;            // its device time (accounted to LP superstep 0)
  uint64_t init_candidates = 0;  // bin b or a larger one was non-empty at the last host sync
  pm::DevCounters* cnt = nullptr;      // device
  pm::DevCounters* h_cnt = nullptr;    // pinned host mirror
  pm::RowStat* rowstat = nullptr;      // device, [diameter + 1]
  pm::RowStat* h_rowstat = nullptr;    // pinned
  int rowstat_cap = 0;                 // rows rowstat / h_rowstat can hold
  bool state_ready = false;
  bool fuzzy_ids = false;   // the frontier entries of the last run name vertices, not compact ids (pm_run_fuzzy)

  // ---- NLCC scratch -----------------------------------------------------------------
  uint8_t* ok = nullptr;            // [V] token_source_map value of source s
  uint32_t* src_list = nullptr;     // [V] sources of the current constraint
  unsigned long long* hset = nullptr;  // (vertex, source) set, open addressing
  uint64_t hset_cap = 0;
  uint64_t hset_use = 0;            // power-of-two part of the table the current constraint uses
  std::vector<uint64_t> pool_seen;  // per constraint: most tokens ever stored (sizes the next run's pool / inbox regions)
  std::vector<uint64_t> keys_seen;  // per constraint: most keys ever held by the hash set (sizes the next run's table)
  std::map<std::string, std::vector<uint64_t>> keys_cache;
  std::string pat_dir;              // pattern directory of the loaded pattern
  std::string pat_key;              // (pattern directory, graph, labels, path) the sizes belong to
  uint64_t labels_version = 0;      // 0: degree labels (a function of the graph); else the pm_labels_set call they came from
  uint64_t labels_counter = 0;
  std::map<std::string, std::vector<uint64_t>> pool_cache;
  uint2* pool = nullptr;            // token pool: nem_1 (vertex, source); TDS (parent index, vertex)
  uint64_t pool_cap = 0;
  uint32_t* match_rows = nullptr;   // TDS: materialised walks of the last run of each constraint
  std::vector<std::vector<uint32_t>> subgraphs;  // host copies per constraint
  std::vector<int> subgraph_width;
  std::vector<uint64_t> subgraph_count;
  bool keep_subgraphs = true;       // materialise enumerated walks on the host

  // ---- run bookkeeping -----------------------------------------------------------------
  std::vector<pm_row_t> rows;
  std::vector<std::pair<uint64_t, double>> step_rows;  // result_step
  std::vector<double> iter_seconds;                    // result_iteration
  uint64_t itr = 0;
  pm_run_summary_t summary{};
  std::vector<cudaEvent_t> events;
  // CUDA-event timing of the first-superstep scan kernels (the dominant kernels), per bin
  std::vector<cudaEvent_t> kev2;   // per superstep: before main scan, after it, after the big-row scan
  std::vector<int> kev2_cls;       // per superstep: kernel class of the main scan (0 first, 1 later)
  std::vector<int> kev2_big;       // per superstep: the CTA-per-row scan ran (its end event was recorded)
  cudaEvent_t kev[4][2] = {{nullptr, nullptr}, {nullptr, nullptr}, {nullptr, nullptr}, {nullptr, nullptr}};
  pm_kernel_stats_t kstat[5] = {};  // [3]: the first-superstep signature filter, [4]: the renaming scan
};

namespace pm {

inline int fail(pm_ctx* c, int code, const std::string& msg) {
  if (c) c->err = msg;
  return code;
}

#define PM_CUDA(ctx, call)                                                              \
  do {                                                                                  \
    cudaError_t e_ = (call);                                                            \
    if (e_ != cudaSuccess)                                                              \
      return pm::fail((ctx), PM_ERR_CUDA,                                               \
                      std::string(#call) + ": " + cudaGetErrorString(e_) + " (" +       \
                          __FILE__ + ":" + std::to_string(__LINE__) + ")");             \
  } while (0)

#define PM_LAUNCH_CHECK(ctx)                                \
  do {                                                      \
    (ctx)->launches++;                                      \
    PM_CUDA((ctx), cudaGetLastError());                     \
  } while (0)

// Device blocks are recycled by exact size: re-opening a graph or loading the next pattern asks for the
// same array sizes again, and cudaMalloc / cudaFree of multi-GB blocks (a device-wide synchronisation each)
// would otherwise cost tens of milliseconds per call.  Freed blocks wait in a per-process list, oldest
// evicted first above kCacheCap bytes; everything is released when an allocation fails or a context dies.
struct DevBlockCache {
  static constexpr size_t kCacheCap = 64ull << 30;
  struct Block { void* p; int dev; size_t bytes; };
  std::mutex mu;
  std::unordered_map<void*, std::pair<int, size_t>> live;  // blocks handed out: device, size
  std::list<Block> idle;                                    // freed blocks, oldest first
  size_t idle_bytes = 0;

  void* take(int dev, size_t bytes) {
    std::lock_guard<std::mutex> g(mu);
    for (auto it = idle.begin(); it != idle.end(); ++it)
      if (it->dev == dev && it->bytes == bytes) {
        void* p = it->p;
        idle_bytes -= bytes;
        idle.erase(it);
        live[p] = {dev, bytes};
        return p;
      }
    return nullptr;
  }
  void adopt(void* p, int dev, size_t bytes) {
    std::lock_guard<std::mutex> g(mu);
    live[p] = {dev, bytes};
  }
  void give_back(void* p) {
    std::lock_guard<std::mutex> g(mu);
    auto it = live.find(p);
    if (it == live.end()) { cudaFree(p); return; }
    const Block b{p, it->second.first, it->second.second};
    live.erase(it);
    if (b.bytes > kCacheCap) { cudaFree(p); return; }
    while (!idle.empty() && idle_bytes + b.bytes > kCacheCap) {
      cudaFree(idle.front().p);
      idle_bytes -= idle.front().bytes;
      idle.pop_front();
    }
    idle.push_back(b);
    idle_bytes += b.bytes;
  }
  void flush(int dev) {  // dev < 0: every device
    std::lock_guard<std::mutex> g(mu);
    for (auto it = idle.begin(); it != idle.end();) {
      if (dev < 0 || it->dev == dev) {
        cudaFree(it->p);
        idle_bytes -= it->bytes;
        it = idle.erase(it);
      } else {
        ++it;
      }
    }
  }
};
inline DevBlockCache& dev_cache() {
  static DevBlockCache* cache = new DevBlockCache();  // never destroyed: the driver may be gone at exit
  return *cache;
}

template <class T>
inline int dev_alloc(pm_ctx* c, T** p, uint64_t n, uint64_t* tally = nullptr) {
  *p = nullptr;
  if (n == 0) n = 1;
  const size_t bytes = (n * sizeof(T) + 255) / 256 * 256;
  void* q = dev_cache().take(c->device, bytes);
  if (!q) {
    cudaError_t e = cudaMalloc(&q, bytes);
    if (e != cudaSuccess) {  // make room: drop everything that is only cached
      (void)cudaGetLastError();
      dev_cache().flush(c->device);
      q = nullptr;
      PM_CUDA(c, cudaMalloc(&q, bytes));
    }
    dev_cache().adopt(q, c->device, bytes);
  }
  *p = (T*)q;
  if (tally) *tally += n * sizeof(T);
  return 0;
}
template <class T>
inline void dev_free(T*& p) {
  if (p) dev_cache().give_back((void*)p);
  p = nullptr;
}

// persistent-style launch geometry: a multiple of the 148 SMs
static const int kBlock = 256;
static const int kGridPerSM = 8;
inline int grid_for(uint64_t work_items_per_thread_hint = 0) {
  (void)work_items_per_thread_hint;
  return 148 * kGridPerSM;
}

inline double wall_s() {
  using namespace std::chrono;
  return duration<double>(steady_clock::now().time_since_epoch()).count();
}

}  // namespace pm
