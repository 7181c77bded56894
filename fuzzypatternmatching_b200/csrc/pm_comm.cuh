// pm_comm.cuh — inter-GPU plumbing of libpmgpu.so (one process per GPU, <= 8 GPUs of one NVSwitch box).
//
// Replaces the reference's MPI mailbox + visitor queue exchange
// (/root/reference/include/havoqgt/new_mailbox.hpp:289-428, visitor_queue.hpp:395-434), its
// termination detection (termination_detection.hpp:97-330) and vertex_data::all_{min,max}_reduce
// (impl/vertex_data.hpp:114-127).  Design:
//   * data path: kernels store mask deltas and tokens STRAIGHT into the owner's inbox over NVLink
//     (peer pointers obtained with CUDA IPC; see PeerTab in pm_common.cuh) — no pack / send / unpack;
//   * control path: one small ncclAllGather of a StepMsg per superstep / hop on the compute stream.
//     It is the barrier ("every rank has finished the kernel that wrote into my inbox"), the
//     termination detection (counts and flags) and the delegate reduce in one;
//   * bulk initial state (labels, classes, first masks): ncclAllGather of contiguous slot ranges.
#pragma once

#include <dlfcn.h>
#include <nccl.h>

#include "pm_common.cuh"

namespace pm {

// NCCL is bound at run time, on the first multi-GPU call: a single-GPU user needs no NCCL at all, and
// a process that also runs PyTorch must end up with ONE libnccl (the loader hands back the copy torch
// already mapped under the same soname instead of a second, older one).
namespace dyn {
struct Api {
  void* lib = nullptr;
  decltype(&::ncclGetUniqueId) GetUniqueId = nullptr;
  decltype(&::ncclCommInitRank) CommInitRank = nullptr;
  decltype(&::ncclCommDestroy) CommDestroy = nullptr;
  decltype(&::ncclAllGather) AllGather = nullptr;
  decltype(&::ncclAllReduce) AllReduce = nullptr;
  decltype(&::ncclSend) Send = nullptr;
  decltype(&::ncclRecv) Recv = nullptr;
  decltype(&::ncclGroupStart) GroupStart = nullptr;
  decltype(&::ncclGroupEnd) GroupEnd = nullptr;
  decltype(&::ncclGetErrorString) GetErrorString = nullptr;
  bool ok = false;
};
inline Api& api() {
  static Api a;
  if (a.lib) return a;
  a.lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!a.lib) a.lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!a.lib) return a;
#define PM_SYM(name) a.name = (decltype(a.name))dlsym(a.lib, "nccl" #name)
  PM_SYM(GetUniqueId); PM_SYM(CommInitRank); PM_SYM(CommDestroy); PM_SYM(AllGather); PM_SYM(AllReduce);
  PM_SYM(Send); PM_SYM(Recv); PM_SYM(GroupStart); PM_SYM(GroupEnd); PM_SYM(GetErrorString);
#undef PM_SYM
  a.ok = a.GetUniqueId && a.CommInitRank && a.CommDestroy && a.AllGather && a.AllReduce && a.Send && a.Recv &&
         a.GroupStart && a.GroupEnd && a.GetErrorString;
  return a;
}
}  // namespace dyn
#define ncclGetUniqueId pm::dyn::api().GetUniqueId
#define ncclCommInitRank pm::dyn::api().CommInitRank
#define ncclCommDestroy pm::dyn::api().CommDestroy
#define ncclAllGather pm::dyn::api().AllGather
#define ncclAllReduce pm::dyn::api().AllReduce
#define ncclSend pm::dyn::api().Send
#define ncclRecv pm::dyn::api().Recv
#define ncclGroupStart pm::dyn::api().GroupStart
#define ncclGroupEnd pm::dyn::api().GroupEnd
#define ncclGetErrorString pm::dyn::api().GetErrorString

__constant__ PeerTab c_peer;

// owner of a compact id: rank g owns [off[g], off[g+1])
__device__ __forceinline__ uint32_t cid_owner(uint32_t c) {
  uint32_t o = 0;
  for (int g = 1; g < c_peer.G; ++g) o += c >= c_peer.off[g] ? 1u : 0u;
  return o;
}

#define PM_NCCL(ctx, call)                                                                          \
  do {                                                                                              \
    ncclResult_t r_ = (call);                                                                       \
    if (r_ != ncclSuccess)                                                                          \
      return pm::fail((ctx), PM_ERR_COMM, std::string(#call) + ": " + ncclGetErrorString(r_));      \
  } while (0)

inline ncclComm_t comm_of(pm_ctx* c) { return (ncclComm_t)c->comm; }

// slot <-> global vertex id (host side; kernels only ever see slots)
inline uint64_t slot_of(const pm_ctx* c, uint64_t v) {
  return c->n_ranks == 1 ? v : (v % c->n_ranks) * c->nlmax + v / c->n_ranks;
}
inline uint64_t vertex_of(const pm_ctx* c, uint64_t slot) {
  return c->n_ranks == 1 ? slot : (slot % c->nlmax) * c->n_ranks + slot / c->nlmax;
}

inline int comm_upload_peers(pm_ctx* c) {
  PM_CUDA(c, cudaMemcpyToSymbolAsync(c_peer, &c->peers, sizeof(PeerTab), 0, cudaMemcpyHostToDevice, c->stream));
  return 0;
}

// Makes `mine` (a cudaMalloc'd allocation base, the same buffer on every rank) addressable from
// every peer: out[g] = pointer valid on THIS device to rank g's buffer.  Collective.
inline int comm_share(pm_ctx* c, void* mine, void** out) {
  const int G = c->n_ranks;
  for (int g = 0; g < G; ++g) out[g] = nullptr;
  out[c->rank] = mine;
  if (G == 1) return 0;
  cudaIpcMemHandle_t h;
  PM_CUDA(c, cudaIpcGetMemHandle(&h, mine));
  cudaIpcMemHandle_t* d = nullptr;
  PM_CUDA(c, cudaMalloc((void**)&d, sizeof(h) * G));
  PM_CUDA(c, cudaMemcpyAsync(d + c->rank, &h, sizeof(h), cudaMemcpyHostToDevice, c->stream));
  PM_NCCL(c, ncclAllGather(d + c->rank, d, sizeof(h), ncclChar, comm_of(c), c->stream));
  std::vector<cudaIpcMemHandle_t> all(G);
  PM_CUDA(c, cudaMemcpyAsync(all.data(), d, sizeof(h) * G, cudaMemcpyDeviceToHost, c->stream));
  PM_CUDA(c, cudaStreamSynchronize(c->stream));
  cudaFree(d);
  for (int g = 0; g < G; ++g) {
    if (g == c->rank) continue;
    void* p = nullptr;
    PM_CUDA(c, cudaIpcOpenMemHandle(&p, all[g], cudaIpcMemLazyEnablePeerAccess));
    c->ipc_open.push_back(p);
    out[g] = p;
  }
  return 0;
}

// Drops every peer mapping this rank holds.  Collective when several ranks cooperate: nobody frees a
// shared buffer before every peer has unmapped it.
inline void comm_close_all(pm_ctx* c) {
  if (c->ipc_open.empty() && c->n_ranks == 1) return;
  cudaStreamSynchronize(c->stream);
  for (void* p : c->ipc_open) cudaIpcCloseMemHandle(p);
  c->ipc_open.clear();
  if (c->n_ranks > 1 && c->comm) {
    int* d = nullptr;
    if (cudaMalloc((void**)&d, 4) == cudaSuccess) {
      cudaMemsetAsync(d, 0, 4, c->stream);
      ncclAllReduce(d, d, 1, ncclInt, ncclSum, comm_of(c), c->stream);
      cudaStreamSynchronize(c->stream);
      cudaFree(d);
    }
  }
}

// fills step_msg[0] from the device counters
// (and re-arms the per-step counters)
__global__ void k_step_msg(DevCounters* cnt, StepMsg* msg) {
  if (threadIdx.x < PM_MAX_RANKS) {
    const unsigned long long n = cnt->out_n[threadIdx.x];
    msg->out_n[threadIdx.x] = n;
    cnt->out_n[threadIdx.x] = 0;
    atomicMax(&cnt->peak_out, n);
  }
  if (threadIdx.x == 0) {
    msg->ndelta = cnt->ndelta;
    cnt->ndelta = 0;
    msg->nf = cnt->nf;
    msg->pad = cnt->nf_init;
    msg->found = cnt->found;
    msg->deleted = cnt->deleted;
    msg->overflow = cnt->overflow;
    msg->accepted = cnt->pool_n;
    msg->peak = cnt->peak_out;
    msg->ce_n = cnt->ce_n;
    msg->n_c = cnt->n_c;
    msg->pad1 = 0;
    msg->seq = 0;
    msg->timeout = 0;
  }
}

// The step barrier over peer memory, one tiny kernel: every rank stores its StepMsg straight into the
// step mailbox of every peer (NVLink stores, sequence number last behind a system fence) and then polls
// its own mailbox until all peers' messages of this step have arrived.  Stores a rank issued before the
// barrier (mask deltas, tokens) are ordered before its sequence number, so they are visible afterwards.
// Mailbox slots alternate with the step parity: a rank is never more than one step ahead of a peer
// that has not yet read the previous message.  A peer that never arrives trips a time-out instead of a hang.
__global__ void k_step_sync(DevCounters* cnt, StepMsg* msg, StepMsg* all_out, uint32_t seq) {
  const int G = c_peer.G, me = c_peer.rank;
  const int lane = threadIdx.x;
  // 1. my message (and re-arm the per-step counters)
  if (lane < PM_MAX_RANKS) {
    const unsigned long long n = cnt->out_n[lane];
    msg->out_n[lane] = n;
    cnt->out_n[lane] = 0;
    atomicMax(&cnt->peak_out, n);
  }
  if (lane == 0) {
    msg->ndelta = cnt->ndelta;
    cnt->ndelta = 0;
    msg->nf = cnt->nf;
    msg->pad = cnt->nf_init;
    msg->found = cnt->found;
    msg->deleted = cnt->deleted;
    msg->overflow = cnt->overflow;
    msg->accepted = cnt->pool_n;
    msg->peak = cnt->peak_out;
    msg->ce_n = cnt->ce_n;
    msg->n_c = cnt->n_c;
    msg->pad1 = 0;
    msg->seq = seq;
    msg->timeout = 0;
  }
  __syncwarp();
  __threadfence_system();  // everything this GPU stored into peers' inboxes before this step is out
  // 2. lane g delivers it to rank g
  const uint32_t slot = (seq & 1u) * (uint32_t)G + (uint32_t)me;
  if (lane < G) {
    volatile unsigned long long* dst = reinterpret_cast<volatile unsigned long long*>(c_peer.sync_in[lane] + slot);
    const unsigned long long* srcw = reinterpret_cast<const unsigned long long*>(msg);
    constexpr int W = sizeof(StepMsg) / 8;
    for (int i = 0; i < W - 1; ++i) dst[i] = srcw[i];
    __threadfence_system();
    dst[W - 1] = srcw[W - 1];  // {seq, timeout}: the arrival flag
  }
  // 3. lane g waits for rank g's message of this step
  bool late = false;
  if (lane < G) {
    const volatile StepMsg* in = c_peer.sync_in[me] + (seq & 1u) * (uint32_t)G + lane;
    unsigned long long t0 = 0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    while (in->seq != seq) {
      unsigned long long t1;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if (t1 - t0 > 20000000000ull) { late = true; break; }  // 20 s
    }
    __threadfence_system();
    const unsigned long long* srcw = reinterpret_cast<const unsigned long long*>(const_cast<const StepMsg*>(in));
    unsigned long long* dstw = reinterpret_cast<unsigned long long*>(all_out + lane);
    for (int i = 0; i < (int)(sizeof(StepMsg) / 8); ++i) dstw[i] = srcw[i];
    if (late) { all_out[lane].timeout = 1; cnt->overflow = 1u; }
  }
}

// The per-step barrier: every rank contributes its StepMsg; afterwards step_msg[1 + g] holds rank
// g's on every rank, and all stores rank g issued before the call are visible here.
inline int comm_step(pm_ctx* c) {
  if (c->n_ranks == 1) return 0;
  if (!c->step_nccl) {
    c->step_seq++;
    k_step_sync<<<1, 32, 0, c->stream>>>(c->cnt, c->step_msg, c->step_msg + 1, c->step_seq);
    PM_LAUNCH_CHECK(c);
    return 0;
  }
  k_step_msg<<<1, 32, 0, c->stream>>>(c->cnt, c->step_msg);
  PM_LAUNCH_CHECK(c);
  PM_NCCL(c, ncclAllGather(c->step_msg, c->step_msg + 1, sizeof(StepMsg), ncclChar, comm_of(c), c->stream));
  return 0;
}

// copies everyone's StepMsg to the host (call after comm_step; synchronises the stream)
inline int comm_step_fetch(pm_ctx* c) {
  PM_CUDA(c, cudaMemcpyAsync(c->h_step, c->step_msg + 1, sizeof(StepMsg) * c->n_ranks, cudaMemcpyDeviceToHost, c->stream));
  PM_CUDA(c, cudaStreamSynchronize(c->stream));
  for (int g = 0; g < c->n_ranks; ++g)
    if (c->h_step[g].timeout) return fail(c, PM_ERR_COMM, "step barrier timed out waiting for rank " + std::to_string(g));
  return 0;
}

// all-gathers the contiguous per-rank slot ranges of a replicated array in place
template <class T>
inline int comm_allgather_slots(pm_ctx* c, T* replicated) {
  if (c->n_ranks == 1) return 0;
  PM_NCCL(c, ncclAllGather(replicated + (uint64_t)c->rank * c->nlmax, replicated, c->nlmax * sizeof(T), ncclChar,
                           comm_of(c), c->stream));
  return 0;
}

// all-gathers equally sized per-rank segments of an array in place (segment g = [g * n, (g + 1) * n))
template <class T>
inline int comm_allgather_seg(pm_ctx* c, T* array, uint64_t n) {
  if (c->n_ranks == 1 || n == 0) return 0;
  PM_NCCL(c, ncclAllGather(array + (uint64_t)c->rank * n, array, n * sizeof(T), ncclChar, comm_of(c), c->stream));
  return 0;
}

inline int comm_allreduce_u64(pm_ctx* c, uint64_t* host_values, int n, ncclRedOp_t op) {
  if (c->n_ranks == 1) return 0;
  if (n > 8 || !c->d_scratch) return fail(c, PM_ERR_ARG, "comm_allreduce_u64: bad size");
  unsigned long long* d = c->d_scratch;
  PM_CUDA(c, cudaMemcpyAsync(d, host_values, 8 * n, cudaMemcpyHostToDevice, c->stream));
  PM_NCCL(c, ncclAllReduce(d, d, n, ncclUint64, op, comm_of(c), c->stream));
  PM_CUDA(c, cudaMemcpyAsync(host_values, d, 8 * n, cudaMemcpyDeviceToHost, c->stream));
  PM_CUDA(c, cudaStreamSynchronize(c->stream));
  return 0;
}
inline int comm_allreduce_max_u64(pm_ctx* c, uint64_t* host_value) { return comm_allreduce_u64(c, host_value, 1, ncclMax); }
inline int comm_allreduce_sum_u64(pm_ctx* c, uint64_t* host_values, int n) { return comm_allreduce_u64(c, host_values, n, ncclSum); }

}  // namespace pm
