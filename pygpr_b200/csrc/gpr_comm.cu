// pygpr_b200 — communicator entry points of the C ABI (gpr_comm_init / allreduce_f64 / destroy).
//
// The Python package shards experts with torch.distributed (pygpr_b200/dist.py); these entry points give a consumer WITHOUT torch the
// same two collectives of the path — all-reduce(sum, f64) of the [3, m] partial sums of GRBCM.aggregate (gr_bcm.py:133,143-144) and of
// the [1+P] NLML / gradient scalars of a shared-vector training step — on the device buffers the other entry points produce.
// Thin wrappers over NCCL (ring / tree / NVLS over NVLink 5 is NCCL's choice: the messages are KBs, latency bound).  NCCL is bound at
// run time (dlopen "libnccl.so.2"): the library loads and every other entry point works where NCCL is absent; gpr_comm_* then return
// GPR_ERR_UNSUPPORTED.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <string.h>

#include <mutex>

#include "../../include/pygpr_b200.h"

extern "C" int gpr_internal_fail(int code, const char* fmt, const char* a, long long v);   // gpr_kernels.cu

namespace {

// the part of nccl.h these wrappers need (stable since NCCL 2.0)
struct NcclUniqueId { char internal[128]; };
typedef void* NcclComm;
constexpr int NCCL_FLOAT64 = 8, NCCL_SUM = 0;

struct Nccl {
    void* handle = nullptr;
    int (*GetUniqueId)(NcclUniqueId*) = nullptr;
    int (*CommInitRank)(NcclComm*, int, NcclUniqueId, int) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    int (*CommDestroy)(NcclComm) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    bool ok = false;
};

Nccl& nccl() {
    static Nccl n;
    static std::once_flag once;
    std::call_once(once, [] {
        for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
            n.handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (n.handle) break;
        }
        if (!n.handle) return;
        n.GetUniqueId = reinterpret_cast<decltype(n.GetUniqueId)>(dlsym(n.handle, "ncclGetUniqueId"));
        n.CommInitRank = reinterpret_cast<decltype(n.CommInitRank)>(dlsym(n.handle, "ncclCommInitRank"));
        n.AllReduce = reinterpret_cast<decltype(n.AllReduce)>(dlsym(n.handle, "ncclAllReduce"));
        n.CommDestroy = reinterpret_cast<decltype(n.CommDestroy)>(dlsym(n.handle, "ncclCommDestroy"));
        n.GetErrorString = reinterpret_cast<decltype(n.GetErrorString)>(dlsym(n.handle, "ncclGetErrorString"));
        n.ok = n.GetUniqueId && n.CommInitRank && n.AllReduce && n.CommDestroy;
    });
    return n;
}

int nccl_fail(const char* what, int rc) {
    Nccl& n = nccl();
    return gpr_internal_fail(GPR_ERR_CUDA, what, (n.GetErrorString ? n.GetErrorString(rc) : "NCCL error"), rc);
}
int no_nccl() { return gpr_internal_fail(GPR_ERR_UNSUPPORTED, "NCCL (libnccl.so.2) could not be loaded%s", "", 0); }

}  // namespace

extern "C" int gpr_comm_unique_id_bytes(void) { return (int)sizeof(NcclUniqueId); }

extern "C" int gpr_comm_unique_id(void* id) {
    if (!id) return gpr_internal_fail(GPR_ERR_BAD_ARG, "bad argument: gpr_comm_unique_id%s", "", 0);
    Nccl& n = nccl();
    if (!n.ok) return no_nccl();
    const int rc = n.GetUniqueId(static_cast<NcclUniqueId*>(id));
    return rc ? nccl_fail("ncclGetUniqueId: %s (%lld)", rc) : GPR_OK;
}

extern "C" int gpr_comm_init(void** comm, int world, int rank, const void* id) {
    if (!(comm && id && world >= 1 && rank >= 0 && rank < world)) return gpr_internal_fail(GPR_ERR_BAD_ARG, "bad argument: gpr_comm_init%s (rank=%lld)", "", rank);
    Nccl& n = nccl();
    if (!n.ok) return no_nccl();
    NcclUniqueId uid;
    memcpy(&uid, id, sizeof(uid));
    NcclComm c = nullptr;
    const int rc = n.CommInitRank(&c, world, uid, rank);   // collective: every rank calls it with rank 0's id, on its own current device
    if (rc) return nccl_fail("ncclCommInitRank: %s (%lld)", rc);
    *comm = c;
    return GPR_OK;
}

extern "C" int gpr_comm_allreduce_f64(void* comm, double* buf, long long count, void* stream) {
    if (!(comm && buf && count > 0)) return gpr_internal_fail(GPR_ERR_BAD_ARG, "bad argument: gpr_comm_allreduce_f64%s (count=%lld)", "", count);
    Nccl& n = nccl();
    if (!n.ok) return no_nccl();
    const int rc = n.AllReduce(buf, buf, (size_t)count, NCCL_FLOAT64, NCCL_SUM, comm, (cudaStream_t)stream);   // in place, stream ordered
    return rc ? nccl_fail("ncclAllReduce: %s (%lld)", rc) : GPR_OK;
}

extern "C" int gpr_comm_destroy(void* comm) {
    if (!comm) return GPR_OK;
    Nccl& n = nccl();
    if (!n.ok) return no_nccl();
    const int rc = n.CommDestroy(comm);
    return rc ? nccl_fail("ncclCommDestroy: %s (%lld)", rc) : GPR_OK;
}
