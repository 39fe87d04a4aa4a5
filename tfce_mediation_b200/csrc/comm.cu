// Collection of the per-shuffle maxima across GPUs: tmb_allgather_max.
//
// The reference's "collective" is N worker processes appending lines to shared CSV files
// (STEP_2_tfce_randomise_parallel.py:139-157, pyfunc.py:119).  Here every rank (one process per GPU) owns a contiguous
// slice of the permutation range and the maxima -- a few kilobytes to a few megabytes per job -- are collected with
// ONE ncclAllGather over NVLink at the end of the job.  There is no exchange inside a shuffle, so nothing to fuse with
// a compute kernel: the cost is one launch.
//
// NCCL is bound at run time (dlopen of libnccl.so.2: inside a PyTorch process that is the copy torch already loaded,
// otherwise the system library), so libtfce_b200.so has no link-time dependency on it and single-GPU users never load it.
// The 128-byte ncclUniqueId travels between the ranks by whatever the host side has (torch.distributed broadcast in
// parallel.py; MPI or a file would do as well).
#include "common.cuh"
#include "../../include/tfce_b200.h"

#include <dlfcn.h>
#include <cstring>

namespace {

typedef struct ncclComm *ncclComm_t;
struct NcclId { char internal[128]; };
typedef int ncclResult_t;
constexpr int kNcclFloat = 7; // ncclFloat32

struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(NcclId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, NcclId, int) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*GetVersion)(int *) = nullptr;
};

NcclApi g_nccl;

int load_nccl() {
    if (g_nccl.handle) return 0;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL); // already in the process (PyTorch)?
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) {
        tmb::set_error("tmb_comm: cannot load libnccl.so.2 (%s)", dlerror());
        return 1;
    }
    NcclApi a;
    a.handle = h;
    a.GetUniqueId = reinterpret_cast<decltype(a.GetUniqueId)>(dlsym(h, "ncclGetUniqueId"));
    a.CommInitRank = reinterpret_cast<decltype(a.CommInitRank)>(dlsym(h, "ncclCommInitRank"));
    a.AllGather = reinterpret_cast<decltype(a.AllGather)>(dlsym(h, "ncclAllGather"));
    a.CommDestroy = reinterpret_cast<decltype(a.CommDestroy)>(dlsym(h, "ncclCommDestroy"));
    a.GetErrorString = reinterpret_cast<decltype(a.GetErrorString)>(dlsym(h, "ncclGetErrorString"));
    a.GetVersion = reinterpret_cast<decltype(a.GetVersion)>(dlsym(h, "ncclGetVersion"));
    if (!a.GetUniqueId || !a.CommInitRank || !a.AllGather || !a.CommDestroy || !a.GetErrorString) {
        tmb::set_error("tmb_comm: libnccl.so.2 lacks an expected symbol");
        return 1;
    }
    g_nccl = a;
    return 0;
}

#define TMB_NCCL(call)                                                                             \
    do {                                                                                           \
        ncclResult_t r__ = (call);                                                                 \
        if (r__ != 0) {                                                                            \
            ::tmb::set_error("%s failed: %s", #call, g_nccl.GetErrorString(r__));                  \
            return 1;                                                                              \
        }                                                                                          \
    } while (0)

} // namespace

struct tmb_comm {
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1, device = 0;
};

using namespace tmb;

extern "C" int tmb_comm_unique_id(void *id_out) {
    TMB_REQUIRE(id_out, "tmb_comm_unique_id: null pointer");
    if (load_nccl()) return 1;
    NcclId id;
    TMB_NCCL(g_nccl.GetUniqueId(&id));
    memcpy(id_out, &id, sizeof(id));
    return 0;
}

extern "C" int tmb_comm_create(const void *id, int rank, int world, int device, tmb_comm **out) {
    TMB_REQUIRE(id && out && world >= 1 && rank >= 0 && rank < world, "tmb_comm_create: bad arguments");
    if (load_nccl()) return 1;
    TMB_ON_DEVICE(device);
    NcclId nid;
    memcpy(&nid, id, sizeof(nid));
    tmb_comm *c = new tmb_comm();
    c->rank = rank; c->world = world; c->device = device;
    ncclResult_t r = g_nccl.CommInitRank(&c->comm, world, nid, rank);
    if (r != 0) {
        set_error("ncclCommInitRank failed: %s", g_nccl.GetErrorString(r));
        delete c;
        return 1;
    }
    *out = c;
    return 0;
}

extern "C" int tmb_comm_destroy(tmb_comm *c) {
    if (!c) return 0;
    DeviceGuard guard(c->device);
    if (c->comm) g_nccl.CommDestroy(c->comm);
    delete c;
    return 0;
}

// global_dev[r * count + i] = rank r's local_dev[i]: every rank contributes `count` floats (pad short slices with zeros;
// the host knows every rank's real count from the shard rule, so no size exchange is needed).
extern "C" int tmb_allgather_max(tmb_comm *c, const float *local_dev, int64_t count, float *global_dev, void *stream) {
    TMB_REQUIRE(c && local_dev && global_dev && count >= 0, "tmb_allgather_max: bad arguments");
    if (count == 0) return 0;
    TMB_ON_DEVICE(c->device);
    TMB_NCCL(g_nccl.AllGather(local_dev, global_dev, (size_t)count, kNcclFloat, c->comm, (cudaStream_t)stream));
    count_launch();
    return 0;
}
