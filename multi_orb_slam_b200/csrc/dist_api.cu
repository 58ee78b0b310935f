// orbd_*: the one collective of the path behind the C ABI (include/orb_b200.h) — the all-gather of the per-camera
// keypoint / descriptor blocks that gives every GPU the descriptors of all cameras of a rig-frame, the multi-GPU
// analogue of Frame::mDescriptors_total (reference src/Frame.cc:170,191-194), consumed by the per-camera loops of
// src/ORBmatcher.cc:628,2030,2269,3582.
//
// NCCL is bound at run time (dlopen of libnccl.so.2: the host process decides which NCCL it carries — a C++ host
// links its own, the Python harness already holds torch's), so liborb_b200.so itself keeps depending on libcudart
// only and single-GPU users never touch NCCL.  The handful of NCCL entry points used are declared here from NCCL's
// public C API (nccl.h: ncclGetUniqueId, ncclCommInitRank, ncclAllGather, ncclCommDestroy, ncclGetErrorString).
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdint.h>

#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>

#include "../../include/orb_b200.h"
#include "device_guard.h"

namespace {

typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;  // NCCL_UNIQUE_ID_BYTES
typedef int ncclResult_t;                             // ncclSuccess == 0
enum { kNcclUint8 = 1 };                              // ncclUint8 / ncclChar family: 1-byte elements

struct NcclApi {
  void* so = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*GetVersion)(int*) = nullptr;
};

std::mutex g_mu;
NcclApi g_nccl;
std::string g_err;

bool load_nccl() {
  std::lock_guard<std::mutex> lock(g_mu);
  if (g_nccl.so) return true;
  const char* names[] = {getenv("ORB_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
  void* so = nullptr;
  std::string why = "not found";
  for (const char* n : names) {
    if (!n || !*n) continue;
    so = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (so) break;
    if (const char* e = dlerror()) why = e;  // dlerror() clears itself: read it once
  }
  if (!so) {
    g_err = "cannot load NCCL (set ORB_NCCL_LIB): " + why;
    return false;
  }
  NcclApi a;
  a.so = so;
  a.GetUniqueId = reinterpret_cast<decltype(a.GetUniqueId)>(dlsym(so, "ncclGetUniqueId"));
  a.CommInitRank = reinterpret_cast<decltype(a.CommInitRank)>(dlsym(so, "ncclCommInitRank"));
  a.AllGather = reinterpret_cast<decltype(a.AllGather)>(dlsym(so, "ncclAllGather"));
  a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(dlsym(so, "ncclCommDestroy"));
  a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(dlsym(so, "ncclGetErrorString"));
  a.GetVersion = reinterpret_cast<decltype(a.GetVersion)>(dlsym(so, "ncclGetVersion"));
  if (!a.GetUniqueId || !a.CommInitRank || !a.AllGather || !a.CommDestroy || !a.GetErrorString) {
    g_err = "the NCCL library lacks a required symbol";
    dlclose(so);
    return false;
  }
  g_nccl = a;
  return true;
}

}  // namespace

struct orbd_comm {
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1, device = 0;
  std::string err;
  bool ok(ncclResult_t r, const char* what) {
    if (r == 0) return true;
    err = std::string(what) + ": " + g_nccl.GetErrorString(r);
    return false;
  }
};

extern "C" {
#pragma GCC visibility push(default)

const char* orbd_last_error(const orbd_comm* c) { return c ? c->err.c_str() : g_err.c_str(); }

int orbd_nccl_version(void) {
  int v = 0;
  if (!load_nccl() || !g_nccl.GetVersion || g_nccl.GetVersion(&v) != 0) return ORBX_E_STATE;
  return v;
}

int orbd_get_unique_id(uint8_t* id128) {
  if (!id128) return ORBX_E_INVALID;
  if (!load_nccl()) return ORBX_E_STATE;
  ncclUniqueId id;
  const ncclResult_t r = g_nccl.GetUniqueId(&id);
  if (r != 0) { g_err = std::string("ncclGetUniqueId: ") + g_nccl.GetErrorString(r); return ORBX_E_CUDA; }
  std::memcpy(id128, id.internal, ORBD_UNIQUE_ID_BYTES);
  return ORBX_OK;
}

int orbd_comm_create(int rank, int world, const uint8_t* id128, int device, orbd_comm** out) {
  if (!out) return ORBX_E_INVALID;
  *out = nullptr;
  if (!id128 || world < 1 || rank < 0 || rank >= world) { g_err = "invalid rank / world / id"; return ORBX_E_INVALID; }
  if (!load_nccl()) return ORBX_E_STATE;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { g_err = "no CUDA device (this library has no CPU fallback)"; return ORBX_E_CUDA; }
  orbd_comm* c = new orbd_comm();
  c->rank = rank;
  c->world = world;
  if (device < 0) cudaGetDevice(&c->device); else c->device = device;
  OrbDeviceGuard guard(c->device);
  ncclUniqueId id;
  std::memcpy(id.internal, id128, ORBD_UNIQUE_ID_BYTES);
  if (!c->ok(g_nccl.CommInitRank(&c->comm, world, id, rank), "ncclCommInitRank")) {
    g_err = c->err;
    delete c;
    return ORBX_E_CUDA;
  }
  *out = c;
  return ORBX_OK;
}

void orbd_comm_destroy(orbd_comm* c) {
  if (!c) return;
  if (c->comm) {
    OrbDeviceGuard guard(c->device);
    g_nccl.CommDestroy(c->comm);
  }
  delete c;
}

int orbd_rank(const orbd_comm* c) { return c ? c->rank : ORBX_E_INVALID; }
int orbd_world(const orbd_comm* c) { return c ? c->world : ORBX_E_INVALID; }

int orbd_allgather_inplace(orbd_comm* c, void* d_buf, size_t bytes_per_rank, void* cuda_stream) {
  if (!c || !d_buf) return ORBX_E_INVALID;
  if (bytes_per_rank == 0 || c->world == 1) return ORBX_OK;  // a single rank already holds everything
  OrbDeviceGuard guard(c->device);
  // in place: NCCL takes sendbuff == recvbuff + rank * sendcount
  const uint8_t* send = static_cast<const uint8_t*>(d_buf) + (size_t)c->rank * bytes_per_rank;
  return c->ok(g_nccl.AllGather(send, d_buf, bytes_per_rank, kNcclUint8, c->comm, (cudaStream_t)cuda_stream), "ncclAllGather")
             ? ORBX_OK : ORBX_E_CUDA;
}

#pragma GCC visibility pop
}
