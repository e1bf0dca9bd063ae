// reduce.cu — sgc_reduce_counts: the one exchange step of the count path.
//
// The reference collects one Counter per sample (count.rs:136) and never splits a sample.  Here a
// large sample is cut into read shards, one sgc_counter per (GPU, shard); the per-guide count
// vectors u64[n_guides + 2] (counts, total_reads, matched_reads) are the only state that crosses
// GPUs.  Shards held by ONE process are summed here: same-device shards by a fold kernel,
// different devices by one ncclReduce over NVLink/NVSwitch, enqueued on the counters' own streams.
// (One process per GPU — torchrun, MPI — sums through its own collective on sgc_counter_state.)
//
// NCCL is not a link-time dependency: libnccl.so.2 is opened on first use, so the library loads
// (and every single-GPU path works) on a machine without it, and inside a process that already
// carries an NCCL (PyTorch) the same copy is shared instead of a second one being mapped.
#include <dlfcn.h>

#include <algorithm>
#include <atomic>
#include <map>
#include <mutex>
#include <vector>

#include "internal.h"

namespace sgc {
namespace {

// the part of nccl.h this file needs (stable across NCCL 2.x)
using ncclComm_t = struct ncclComm*;
enum { kNcclSuccess = 0 };
enum { kNcclUint64 = 5 };  // ncclDataType_t: ncclUint64
enum { kNcclSum = 0 };     // ncclRedOp_t: ncclSum

struct Nccl {
  void* handle = nullptr;
  int (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  int (*Reduce)(const void*, void*, size_t, int, int, int, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  std::string error;
};

Nccl& nccl() {
  static Nccl n;
  static std::once_flag once;
  std::call_once(once, [] {
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
      n.handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (n.handle) break;
    }
    if (!n.handle) {
      n.error = std::string("cannot load libnccl.so.2: ") + (dlerror() ? dlerror() : "not found");
      return;
    }
    auto sym = [&](const char* s) {
      void* p = dlsym(n.handle, s);
      if (!p && n.error.empty()) n.error = std::string("libnccl.so.2 lacks ") + s;
      return p;
    };
    n.CommInitAll = reinterpret_cast<decltype(n.CommInitAll)>(sym("ncclCommInitAll"));
    n.CommDestroy = reinterpret_cast<decltype(n.CommDestroy)>(sym("ncclCommDestroy"));
    n.GroupStart = reinterpret_cast<decltype(n.GroupStart)>(sym("ncclGroupStart"));
    n.GroupEnd = reinterpret_cast<decltype(n.GroupEnd)>(sym("ncclGroupEnd"));
    n.Reduce = reinterpret_cast<decltype(n.Reduce)>(sym("ncclReduce"));
    n.GetErrorString = reinterpret_cast<decltype(n.GetErrorString)>(sym("ncclGetErrorString"));
  });
  return n;
}

int nccl_error(const char* what, int rc) {
  Nccl& n = nccl();
  return set_error(SGC_ERR_NCCL, std::string(what) + ": " + (n.GetErrorString ? n.GetErrorString(rc) : "NCCL error"));
}

// communicators of one device set, in the order of the (sorted) device list; kept for the life
// of the process (loading NCCL and ncclCommInitAll take seconds: sgc_reduce_prepare)
struct CommSet {
  std::vector<int> devices;
  std::vector<ncclComm_t> comms;
};
std::mutex g_comm_mu;
std::map<std::vector<int>, CommSet> g_comms;

// the communicators of a (sorted, duplicate-free) device list, created on first use; call with
// g_comm_mu held
int comms_for(const std::vector<int>& devices, CommSet** out) {
  Nccl& n = nccl();
  if (!n.error.empty()) return set_error(SGC_ERR_NCCL, n.error);
  CommSet& cs = g_comms[devices];
  if (cs.comms.empty()) {
    cs.devices = devices;
    cs.comms.resize(devices.size());
    const int rc = n.CommInitAll(cs.comms.data(), (int)devices.size(), devices.data());
    if (rc != kNcclSuccess) {
      g_comms.erase(devices);
      return nccl_error("ncclCommInitAll", rc);
    }
  }
  *out = &cs;
  return SGC_OK;
}

__global__ void add_state_kernel(unsigned long long* __restrict__ dst, const unsigned long long* __restrict__ src,
                                 size_t n) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] += src[i];
}

std::atomic<bool> g_force_nccl{false};  // sgc_reduce_prepare has been called: the caller wants NCCL, wait for it

// state[root] += state of every other leader, through a scratch vector on the root's device
int peer_reduce(const std::map<int, sgc_counter*>& leader, sgc_counter* root, size_t words) {
  DeviceGuard guard(root->device);
  unsigned long long* scratch = nullptr;
  SGC_CUDA_TRY(cudaMalloc(&scratch, words * sizeof(uint64_t)));
  struct Free {
    void* p;
    cudaStream_t s;
    ~Free() {
      cudaStreamSynchronize(s);
      cudaFree(p);
    }
  } free_scratch{scratch, root->stream};
  for (const auto& kv : leader) {
    sgc_counter* l = kv.second;
    if (l == root) continue;
    cudaEvent_t ev;
    {
      DeviceGuard src(l->device);
      SGC_CUDA_TRY(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
      const cudaError_t e = cudaEventRecord(ev, l->stream);
      if (e != cudaSuccess) {
        cudaEventDestroy(ev);
        SGC_CUDA_TRY(e);
      }
    }
    cudaError_t e = cudaStreamWaitEvent(root->stream, ev, 0);  // the shard's counting has finished
    if (e == cudaSuccess)
      e = cudaMemcpyPeerAsync(scratch, root->device, l->d_state, l->device, words * sizeof(uint64_t), root->stream);
    cudaEventDestroy(ev);
    SGC_CUDA_TRY(e);
    add_state_kernel<<<(unsigned)((words + 255) / 256), 256, 0, root->stream>>>(root->d_state, scratch, words);
    SGC_CUDA_TRY(cudaGetLastError());
  }
  return SGC_OK;
}

}  // namespace
}  // namespace sgc

using namespace sgc;

extern "C" int sgc_reduce_counts(sgc_counter* const* shards, int n_shards, int root) {
  if (!shards || n_shards < 1 || root < 0 || root >= n_shards) return set_error(SGC_ERR_INVALID_ARG, "bad shard list");
  for (int i = 0; i < n_shards; ++i) {
    if (!shards[i]) return set_error(SGC_ERR_INVALID_ARG, "shard is NULL");
    if (shards[i]->lib->n != shards[0]->lib->n || shards[i]->lib->k != shards[0]->lib->k)
      return set_error(SGC_ERR_INVALID_ARG, "shards were built from different libraries");
    for (int j = 0; j < i; ++j)
      if (shards[j] == shards[i]) return set_error(SGC_ERR_INVALID_ARG, "a shard is listed twice");
  }
  if (n_shards == 1) return SGC_OK;
  const size_t words = (size_t)shards[0]->lib->n + 2;

  // 1. one leader per device (the root on its device); the other shards of the device are added
  //    into it on the leader's stream, after their own streams have drained
  std::map<int, sgc_counter*> leader;
  leader[shards[root]->lib->device] = shards[root];
  for (int i = 0; i < n_shards; ++i) leader.emplace(shards[i]->lib->device, shards[i]);
  for (int i = 0; i < n_shards; ++i) {
    sgc_counter* c = shards[i];
    sgc_counter* l = leader[c->lib->device];
    if (c == l) continue;
    DeviceGuard guard(c->lib->device);
    cudaEvent_t ev;
    SGC_CUDA_TRY(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    cudaError_t e = cudaEventRecord(ev, c->stream);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(l->stream, ev, 0);
    cudaEventDestroy(ev);
    SGC_CUDA_TRY(e);
    add_state_kernel<<<(unsigned)((words + 255) / 256), 256, 0, l->stream>>>(l->d_state, c->d_state, words);
    SGC_CUDA_TRY(cudaGetLastError());
  }
  if (leader.size() == 1) return SGC_OK;

  // 2. one rank per device
  std::vector<int> devices;
  for (auto& kv : leader) devices.push_back(kv.first);  // sorted by the map
  // NCCL when its communicators for these devices exist (sgc_reduce_prepare made them, or an
  // earlier reduce did).  Creating them takes seconds — longer than counting a whole sample of
  // tens of millions of reads — so a reduce that finds none (and nobody creating them) copies the
  // n_guides + 2 words of every other device straight into a buffer on the root's device
  // (cudaMemcpyPeerAsync: NVLink peer-to-peer) and adds them there.
  std::unique_lock<std::mutex> lk(g_comm_mu, std::try_to_lock);  // communicators are shared: one reduce at a time
  const bool nccl_ready = lk.owns_lock() && g_comms.count(devices) != 0;
  if (!nccl_ready && !g_force_nccl) {
    if (lk.owns_lock()) lk.unlock();
    return peer_reduce(leader, shards[root], words);
  }
  if (!lk.owns_lock()) lk.lock();
  CommSet* csp = nullptr;
  int rc = comms_for(devices, &csp);
  if (rc) return rc;
  CommSet& cs = *csp;
  Nccl& n = nccl();
  int root_rank = 0;
  for (size_t r = 0; r < devices.size(); ++r)
    if (devices[r] == shards[root]->lib->device) root_rank = (int)r;
  rc = n.GroupStart();
  if (rc != kNcclSuccess) return nccl_error("ncclGroupStart", rc);
  for (size_t r = 0; r < devices.size(); ++r) {
    sgc_counter* l = leader[devices[r]];
    DeviceGuard guard(devices[r]);
    rc = n.Reduce(l->d_state, l->d_state, words, kNcclUint64, kNcclSum, root_rank, cs.comms[r], l->stream);
    if (rc != kNcclSuccess) {
      n.GroupEnd();
      return nccl_error("ncclReduce", rc);
    }
  }
  rc = n.GroupEnd();
  if (rc != kNcclSuccess) return nccl_error("ncclGroupEnd", rc);
  return SGC_OK;
}

// Creates the NCCL communicators sgc_reduce_counts will need for counters on these devices
// (ncclCommInitAll takes seconds: a host starts this on a side thread while it builds its tables).
extern "C" int sgc_reduce_prepare(const int* devices, int n_devices) {
  if (!devices || n_devices < 1) return set_error(SGC_ERR_INVALID_ARG, "bad device list");
  std::vector<int> sorted(devices, devices + n_devices);
  std::sort(sorted.begin(), sorted.end());
  sorted.erase(std::unique(sorted.begin(), sorted.end()), sorted.end());
  if (sorted.size() < 2) return SGC_OK;
  g_force_nccl = true;
  std::lock_guard<std::mutex> lk(g_comm_mu);
  CommSet* cs = nullptr;
  return comms_for(sorted, &cs);
}
