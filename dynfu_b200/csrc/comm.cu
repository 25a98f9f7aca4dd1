// NCCL plumbing for data-parallel solves: a communicator created from C so that the solver can issue its
// all-reduces itself (no host-language callback inside the PCG loop).  libnccl is opened at run time
// (dlopen "libnccl.so.2" -- the copy PyTorch already loaded when present), so the library has no link-time
// dependency on NCCL and single-GPU users never need it.
#include <dlfcn.h>
#include <string.h>

#include "dfu_internal.h"

namespace {

typedef struct { char internal[128]; } nccl_unique_id;  // ncclUniqueId (nccl.h:37-38)
typedef void* nccl_comm_t;
typedef int (*fn_get_unique_id)(nccl_unique_id*);
typedef int (*fn_comm_init_rank)(nccl_comm_t*, int, nccl_unique_id, int);
typedef int (*fn_comm_destroy)(nccl_comm_t);
typedef int (*fn_all_reduce)(const void*, void*, size_t, int, int, nccl_comm_t, cudaStream_t);
typedef const char* (*fn_error_string)(int);

struct NcclApi {
    void* handle = nullptr;
    fn_get_unique_id get_unique_id = nullptr;
    fn_comm_init_rank comm_init_rank = nullptr;
    fn_comm_destroy comm_destroy = nullptr;
    fn_all_reduce all_reduce = nullptr;
    fn_error_string error_string = nullptr;
};

NcclApi* nccl() {
    static NcclApi api;
    static bool tried = false;
    if (!tried) {
        tried = true;
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (api.handle) break;
        }
        if (api.handle) {
            api.get_unique_id = (fn_get_unique_id) dlsym(api.handle, "ncclGetUniqueId");
            api.comm_init_rank = (fn_comm_init_rank) dlsym(api.handle, "ncclCommInitRank");
            api.comm_destroy = (fn_comm_destroy) dlsym(api.handle, "ncclCommDestroy");
            api.all_reduce = (fn_all_reduce) dlsym(api.handle, "ncclAllReduce");
            api.error_string = (fn_error_string) dlsym(api.handle, "ncclGetErrorString");
        }
    }
    const bool ok = api.handle && api.get_unique_id && api.comm_init_rank && api.comm_destroy && api.all_reduce;
    return ok ? &api : nullptr;
}

const char* nccl_err(NcclApi* a, int rc) { return a->error_string ? a->error_string(rc) : "NCCL error"; }

}  // namespace

struct dfu_comm {
    nccl_comm_t comm = nullptr;
    int rank = 0, world = 1;
};

extern "C" {

int dfu_comm_unique_id(char id_host[128]) {
    NcclApi* a = nccl();
    DFU_REQUIRE(a != nullptr, DFU_ERR_INVALID, "libnccl.so.2 could not be loaded");
    DFU_REQUIRE(id_host != nullptr, DFU_ERR_INVALID, "NULL argument");
    nccl_unique_id id;
    const int rc = a->get_unique_id(&id);
    if (rc != 0) {
        dfu_set_error("ncclGetUniqueId: %s", nccl_err(a, rc));
        return DFU_ERR_CUDA;
    }
    memcpy(id_host, id.internal, 128);
    return DFU_OK;
}

int dfu_comm_create(dfu_comm** out, const char id_host[128], int rank, int world) {
    NcclApi* a = nccl();
    DFU_REQUIRE(a != nullptr, DFU_ERR_INVALID, "libnccl.so.2 could not be loaded");
    DFU_REQUIRE(out && id_host && world >= 1 && rank >= 0 && rank < world, DFU_ERR_INVALID, "bad argument");
    nccl_unique_id id;
    memcpy(id.internal, id_host, 128);
    dfu_comm* c = new dfu_comm();
    c->rank = rank;
    c->world = world;
    const int rc = a->comm_init_rank(&c->comm, world, id, rank);  // uses the calling thread's current device
    if (rc != 0) {
        dfu_set_error("ncclCommInitRank: %s", nccl_err(a, rc));
        delete c;
        return DFU_ERR_CUDA;
    }
    *out = c;
    return DFU_OK;
}

int dfu_comm_destroy(dfu_comm* c) {
    if (!c) return DFU_OK;
    NcclApi* a = nccl();
    if (a && c->comm) a->comm_destroy(c->comm);
    delete c;
    return DFU_OK;
}

// sum-all-reduce of count floats in place; has the signature of dfu_allreduce_fn with ctx = dfu_comm*
int dfu_comm_allreduce(float* buf, size_t count, void* ctx, dfu_stream stream) {
    NcclApi* a = nccl();
    dfu_comm* c = reinterpret_cast<dfu_comm*>(ctx);
    if (!a || !c || !c->comm) return 1;
    const int rc = a->all_reduce(buf, buf, count, /*ncclFloat32*/ 7, /*ncclSum*/ 0, c->comm, reinterpret_cast<cudaStream_t>(stream));
    if (rc != 0) dfu_set_error("ncclAllReduce: %s", nccl_err(a, rc));
    return rc;
}

int dfu_solver_set_comm(dfu_solver* s, dfu_comm* c) {
    DFU_REQUIRE(s != nullptr, DFU_ERR_INVALID, "NULL argument");
    if (c == nullptr || c->world <= 1) return dfu_solver_set_allreduce(s, nullptr, nullptr);
    return dfu_solver_set_allreduce(s, dfu_comm_allreduce, c);
}

}  // extern "C"
