// comm.cu -- C-ABI of the one collective of the path: the sum over ranks of the packed partial sums of a
// sharded expectation(+gradient) step (include/qhbm_b200.h, "collective").  NCCL is bound at run time
// (dlopen of libnccl.so.2: the copy already loaded in the process, e.g. torch's, else the system one), so the
// library itself links against nothing but the CUDA runtime and single-GPU users never load NCCL.
#include <dlfcn.h>
#include <nccl.h>  // types and prototypes only; no symbol of it is linked

#include <cstring>
#include <mutex>
#include <string>

#include "common.h"
#include "../../include/qhbm_b200.h"

namespace {

struct NcclApi {
  void* handle = nullptr;
  decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
  decltype(&ncclCommInitRank) CommInitRank = nullptr;
  decltype(&ncclCommDestroy) CommDestroy = nullptr;
  decltype(&ncclAllReduce) AllReduce = nullptr;
  decltype(&ncclGetErrorString) GetErrorString = nullptr;
  decltype(&ncclCommCount) CommCount = nullptr;
  decltype(&ncclCommUserRank) CommUserRank = nullptr;
  decltype(&ncclGetVersion) GetVersion = nullptr;
};

template <class F>
void bind(void* h, const char* name, F& fn) {
  fn = reinterpret_cast<F>(dlsym(h, name));
  if (!fn) throw std::runtime_error(std::string("libnccl.so.2 has no symbol ") + name);
}

const NcclApi& nccl() {
  static NcclApi api;
  static std::once_flag once;
  static std::string failure;
  std::call_once(once, [] {
    const char* override_path = std::getenv("QHBM_NCCL_LIB");
    void* h = override_path ? dlopen(override_path, RTLD_NOW | RTLD_GLOBAL) : nullptr;
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);  // already in the process (torch's copy)
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) {
      failure = std::string("cannot load libnccl.so.2 (set QHBM_NCCL_LIB): ") + dlerror();
      return;
    }
    try {
      api.handle = h;
      bind(h, "ncclGetUniqueId", api.GetUniqueId);
      bind(h, "ncclCommInitRank", api.CommInitRank);
      bind(h, "ncclCommDestroy", api.CommDestroy);
      bind(h, "ncclAllReduce", api.AllReduce);
      bind(h, "ncclGetErrorString", api.GetErrorString);
      bind(h, "ncclCommCount", api.CommCount);
      bind(h, "ncclCommUserRank", api.CommUserRank);
      bind(h, "ncclGetVersion", api.GetVersion);
    } catch (const std::exception& e) {
      failure = e.what();
      api.handle = nullptr;
    }
  });
  if (!api.handle) throw std::runtime_error(failure);
  return api;
}

void nccl_check(ncclResult_t r, const char* what) {
  if (r != ncclSuccess) throw std::runtime_error(std::string(what) + ": " + nccl().GetErrorString(r));
}

}  // namespace

struct qhbm_comm {
  ncclComm_t comm = nullptr;
  bool owned = false;
  int rank = 0, nranks = 1;
};

using namespace qhbm;

extern "C" {

int qhbm_comm_unique_id(uint8_t* out, int32_t out_bytes) {
  return guarded([&] {
    if (!out || out_bytes < (int32_t)sizeof(ncclUniqueId))
      throw std::runtime_error("unique id buffer must hold QHBM_COMM_ID_BYTES bytes");
    ncclUniqueId id;
    nccl_check(nccl().GetUniqueId(&id), "ncclGetUniqueId");
    std::memcpy(out, &id, sizeof(id));
  });
}

int qhbm_comm_create(const uint8_t* id, int32_t rank, int32_t nranks, qhbm_comm_t** out) {
  return guarded([&] {
    if (!id || !out) throw std::runtime_error("null argument");
    if (nranks < 1 || rank < 0 || rank >= nranks) throw std::runtime_error("rank out of range");
    ncclUniqueId uid;
    std::memcpy(&uid, id, sizeof(uid));
    auto* c = new qhbm_comm;
    ncclResult_t r = nccl().CommInitRank(&c->comm, nranks, uid, rank);  // on the calling thread's current device
    if (r != ncclSuccess) {
      delete c;
      nccl_check(r, "ncclCommInitRank");
    }
    c->owned = true;
    c->rank = rank;
    c->nranks = nranks;
    *out = c;
  });
}

int qhbm_comm_adopt(void* nccl_comm, qhbm_comm_t** out) {
  return guarded([&] {
    if (!nccl_comm || !out) throw std::runtime_error("null argument");
    auto* c = new qhbm_comm;
    c->comm = static_cast<ncclComm_t>(nccl_comm);
    ncclResult_t r = nccl().CommCount(c->comm, &c->nranks);
    if (r == ncclSuccess) r = nccl().CommUserRank(c->comm, &c->rank);
    if (r != ncclSuccess) {
      delete c;
      nccl_check(r, "ncclCommCount / ncclCommUserRank");
    }
    *out = c;
  });
}

int qhbm_comm_info(const qhbm_comm_t* c, int32_t* rank, int32_t* nranks, int32_t* nccl_version) {
  return guarded([&] {
    if (!c) throw std::runtime_error("null communicator");
    if (rank) *rank = c->rank;
    if (nranks) *nranks = c->nranks;
    if (nccl_version) {
      int v = 0;
      nccl_check(nccl().GetVersion(&v), "ncclGetVersion");
      *nccl_version = v;
    }
  });
}

void qhbm_comm_destroy(qhbm_comm_t* c) {
  if (!c) return;
  if (c->owned && c->comm) {
    try {
      nccl().CommDestroy(c->comm);
    } catch (...) {
    }
  }
  delete c;
}

int qhbm_allreduce(qhbm_comm_t* c, void* d_buf, int64_t count, int32_t dtype, void* stream) {
  return guarded([&] {
    if (!c) throw std::runtime_error("null communicator");
    if (count < 0 || (count > 0 && !d_buf)) throw std::runtime_error("bad buffer");
    if (dtype != QHBM_F32 && dtype != QHBM_F64) throw std::runtime_error("dtype must be QHBM_F32 or QHBM_F64");
    if (count == 0) return;
    nccl_check(nccl().AllReduce(d_buf, d_buf, (size_t)count, dtype == QHBM_F32 ? ncclFloat32 : ncclFloat64, ncclSum,
                                c->comm, (cudaStream_t)stream),
               "ncclAllReduce");
  });
}

}  // extern "C"
