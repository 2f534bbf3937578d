// comm.cpp — the multi-GPU layer of the library (the reference is single-GPU and has no counterpart).
//
// One process per GPU.  Perturbation directions are independent given the real state, which is deterministic and therefore
// identical on every rank: each rank carries its share of the derivative components (for a Hessian batch: every first-order
// component and a share of the second-order pairs) and the ranks exchange only the per-frame pose records - one NCCL
// all-gather of (1 + ncomp_max) x 16 floats per rank, queued by the frame loop itself on a stream of its own behind the
// frame's record upload (kinfu.cpp), so it runs beside the next frame's kernels over NVLink.
//
// NCCL is bound at run time (dlopen of libnccl.so.2: the copy PyTorch has already loaded in a Python process, the system
// one in the C++ driver), so libxslam_b200.so has no link-time dependency on a particular NCCL build; the five entry points
// used are declared here with their published signatures (nccl.h, NCCL 2.x).  Without NCCL every call fails loudly.
// (A process that also imports PyTorch must import it before the first call here, so that the soname resolves to PyTorch's
// bundled NCCL rather than an older system copy: x-slam_b200/parallel.py does.)
#include "../../include/xslam_b200.h"

#include <cuda_runtime_api.h>
#include <dlfcn.h>

#include <cstring>
#include <string>

namespace xs {
void set_error(const std::string &msg);
}
using xs::set_error;

namespace {

typedef struct {
    char internal[128];
} nccl_unique_id;  // ncclUniqueId, NCCL_UNIQUE_ID_BYTES = 128
typedef void *nccl_comm;
constexpr int NCCL_FLOAT = 7;  // ncclFloat32

struct NcclApi {
    int (*GetUniqueId)(nccl_unique_id *) = nullptr;
    int (*CommInitRank)(nccl_comm *, int, nccl_unique_id, int) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, nccl_comm, cudaStream_t) = nullptr;
    int (*CommDestroy)(nccl_comm) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    bool ok = false;
};

const NcclApi &nccl() {
    static NcclApi api;
    static bool tried = false;
    if (tried) return api;
    tried = true;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return api;
    api.GetUniqueId = (int (*)(nccl_unique_id *)) dlsym(h, "ncclGetUniqueId");
    api.CommInitRank = (int (*)(nccl_comm *, int, nccl_unique_id, int)) dlsym(h, "ncclCommInitRank");
    api.AllGather = (int (*)(const void *, void *, size_t, int, nccl_comm, cudaStream_t)) dlsym(h, "ncclAllGather");
    api.CommDestroy = (int (*)(nccl_comm)) dlsym(h, "ncclCommDestroy");
    api.GetErrorString = (const char *(*) (int) ) dlsym(h, "ncclGetErrorString");
    api.ok = api.GetUniqueId && api.CommInitRank && api.AllGather && api.CommDestroy && api.GetErrorString;
    return api;
}

int nccl_fail(const char *what, int rc) {
    set_error(std::string(what) + ": " + (nccl().GetErrorString ? nccl().GetErrorString(rc) : "NCCL error"));
    return XS_ERR_NCCL;
}

}  // namespace

struct xs_comm {
    nccl_comm comm = nullptr;
    int rank = 0, world = 1;
};

extern "C" {

int xs_set_device(int device) {
    if (cudaSetDevice(device) != cudaSuccess) {
        set_error("xs_set_device: no such CUDA device");
        return XS_ERR_CUDA;
    }
    return XS_OK;
}

int xs_comm_unique_id(unsigned char id_out[128]) {
    if (!id_out) return XS_ERR_ARG;
    if (!nccl().ok) {
        set_error("xs_comm: libnccl.so.2 is not available");
        return XS_ERR_NCCL;
    }
    nccl_unique_id id;
    const int rc = nccl().GetUniqueId(&id);
    if (rc != 0) return nccl_fail("ncclGetUniqueId", rc);
    std::memcpy(id_out, id.internal, 128);
    return XS_OK;
}

xs_comm *xs_comm_create(int rank, int world, const unsigned char id[128]) {
    if (!id || world < 1 || rank < 0 || rank >= world) {
        set_error("xs_comm_create: bad arguments");
        return nullptr;
    }
    if (!nccl().ok) {
        set_error("xs_comm: libnccl.so.2 is not available");
        return nullptr;
    }
    nccl_unique_id uid;
    std::memcpy(uid.internal, id, 128);
    xs_comm *c = new xs_comm();
    c->rank = rank;
    c->world = world;
    const int rc = nccl().CommInitRank(&c->comm, world, uid, rank);  // on the calling thread's current device
    if (rc != 0) {
        nccl_fail("ncclCommInitRank", rc);
        delete c;
        return nullptr;
    }
    return c;
}

void xs_comm_destroy(xs_comm *c) {
    if (!c) return;
    if (c->comm && nccl().ok) nccl().CommDestroy(c->comm);
    delete c;
}

int xs_comm_rank(const xs_comm *c) { return c ? c->rank : 0; }
int xs_comm_world(const xs_comm *c) { return c ? c->world : 1; }

// all-gather of `floats` floats per rank on `stream` (device buffers: send [floats], recv [world][floats])
int xs_comm_all_gather(xs_comm *c, const float *d_send, float *d_recv, long floats, void *stream) {
    if (!c || !d_send || !d_recv || floats <= 0) return XS_ERR_ARG;
    const int rc = nccl().AllGather(d_send, d_recv, (size_t) floats, NCCL_FLOAT, c->comm, (cudaStream_t) stream);
    return rc == 0 ? XS_OK : nccl_fail("ncclAllGather", rc);
}

}  // extern "C"
