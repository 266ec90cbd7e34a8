// Data-parallel entry points of the C ABI (scope row D1 / SURVEY 8b last row): the gradient all-reduce over NCCL, taking the
// caller's ncclComm_t and stream.  The reference is single-GPU; this is the only collective of the path.
//
// The library does not link against NCCL: the communicator a caller hands in belongs to ONE NCCL instance (in a PyTorch process
// the one libtorch_cuda loaded), so the symbols are resolved at first use from that already-loaded libnccl.so.2 (dlopen NOLOAD),
// falling back to a plain dlopen for callers that create their own communicator.
#include <dlfcn.h>
#include <string.h>

#include "../../include/qavit_b200.h"
#include "kernels.h"

namespace {

typedef int (*AllReduceFn)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef int (*GroupFn)();
typedef const char* (*ErrStrFn)(int);
typedef int (*CountFn)(void*, int*);

struct Nccl {
  void* h = nullptr;
  AllReduceFn all_reduce = nullptr;
  GroupFn group_start = nullptr, group_end = nullptr;
  ErrStrFn err = nullptr;
  CountFn count = nullptr, rank = nullptr;
};

Nccl* nccl() {
  static Nccl n;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* h = dlopen("libnccl.so.2", RTLD_LAZY | RTLD_NOLOAD);
    if (!h) h = dlopen("libnccl.so.2", RTLD_LAZY | RTLD_GLOBAL);
    if (h) {
      n.h = h;
      n.all_reduce = reinterpret_cast<AllReduceFn>(dlsym(h, "ncclAllReduce"));
      n.group_start = reinterpret_cast<GroupFn>(dlsym(h, "ncclGroupStart"));
      n.group_end = reinterpret_cast<GroupFn>(dlsym(h, "ncclGroupEnd"));
      n.err = reinterpret_cast<ErrStrFn>(dlsym(h, "ncclGetErrorString"));
      n.count = reinterpret_cast<CountFn>(dlsym(h, "ncclCommCount"));
      n.rank = reinterpret_cast<CountFn>(dlsym(h, "ncclCommUserRank"));
    }
  }
  return (n.h && n.all_reduce && n.group_start && n.group_end) ? &n : nullptr;
}

constexpr int kNcclFloat32 = 7, kNcclSum = 0;   // ncclDataType_t / ncclRedOp_t values of nccl.h (stable since NCCL 2.0)

}  // namespace

extern "C" int qavit_dp_available(void) { return nccl() ? 1 : 0; }

extern "C" int qavit_dp_comm_info(void* comm, int* world, int* rank) {
  Nccl* n = nccl();
  QV_CHECK(n && n->count && n->rank, "qavit_dp: libnccl.so.2 is not loadable in this process");
  QV_CHECK(comm, "qavit_dp: null communicator");
  int w = 0, r = 0;
  QV_CHECK(n->count(comm, &w) == 0 && n->rank(comm, &r) == 0, "qavit_dp: ncclCommCount / ncclCommUserRank failed");
  if (world) *world = w;
  if (rank) *rank = r;
  return 0;
}

// SUM all-reduce, in place, of `n_buckets` fp32 device ranges as ONE NCCL group on `stream` (a bucket = a contiguous range of the
// optimizer's flat gradient buffer; the 1 / world of the mean is folded into qavit_clip_grads_scaled).
extern "C" int qavit_dp_allreduce_sum(void* comm, float* const* bufs, const size_t* counts, int n_buckets, void* stream) {
  Nccl* n = nccl();
  QV_CHECK(n, "qavit_dp: libnccl.so.2 is not loadable in this process");
  QV_CHECK(comm && bufs && counts && n_buckets >= 1, "qavit_dp_allreduce_sum: null argument");
  int rc = n->group_start();
  for (int i = 0; i < n_buckets && rc == 0; ++i) {
    if (counts[i] == 0) continue;
    rc = n->all_reduce(bufs[i], bufs[i], counts[i], kNcclFloat32, kNcclSum, comm, static_cast<cudaStream_t>(stream));
  }
  const int rc2 = n->group_end();
  if (rc == 0) rc = rc2;
  QV_CHECK(rc == 0, "qavit_dp_allreduce_sum: NCCL error %d (%s)", rc, n->err ? n->err(rc) : "?");
  return 0;
}
