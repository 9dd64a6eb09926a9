// Multi-GPU side of the C-ABI: the scene is replicated, tiles or sample ranges are partitioned, and the
// per-pixel accumulators of all ranks are summed into rank 0 with ONE ncclReduce per buffer over NVLink / NVSwitch.
// Two ways to form the communicator: one process per GPU (the launcher carries the ncclUniqueId to every rank:
// spcu_comm_unique_id + spcu_comm_init_rank; bench.py under torchrun) or one process driving several contexts
// (spcu_comm_init_all; sp::CudaIntegrator with SPCU_DEVICES=N, one host thread per device).  There is no exchange during
// tracing and the payload is 33 MB per 1080p frame (133 MB at 4K): a fused compute+collective kernel has nothing to overlap.
//
// libnccl.so.2 is bound at the FIRST communicator call, not at load time, and the copy a host process has already mapped wins
// (RTLD_NOLOAD first): a Python process that imports torch gets torch's bundled NCCL, and a process that loads this library
// before torch does not pin the system's older libnccl under the same soname (torch's libtorch_cuda.so would then fail to
// resolve its newer symbols).  Only the eight entry points below are used; their signatures are stable across NCCL 2.x.
#include "ctx.h"

#include <dlfcn.h>
#include <nccl.h>

using namespace spcu;

static_assert(SPCU_NCCL_ID_BYTES == NCCL_UNIQUE_ID_BYTES, "spcu.h must carry a whole ncclUniqueId");

namespace {

struct Nccl
{
    decltype(&ncclGetUniqueId)    GetUniqueId    = nullptr;
    decltype(&ncclCommInitRank)   CommInitRank   = nullptr;
    decltype(&ncclCommInitAll)    CommInitAll    = nullptr;
    decltype(&ncclCommDestroy)    CommDestroy    = nullptr;
    decltype(&ncclReduce)         Reduce         = nullptr;
    decltype(&ncclGroupStart)     GroupStart     = nullptr;
    decltype(&ncclGroupEnd)       GroupEnd       = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    std::string                   error;
    bool                          ok = false;
};

const Nccl& nccl()
{
    static const Nccl api = [] {
        Nccl  a;
        void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
        if (!h) {
            h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
        }
        if (!h) {
            a.error = std::string("libnccl.so.2 cannot be loaded: ") + dlerror();
            return a;
        }
#define SPCU_NCCL_SYM(name)                                                            \
    a.name = reinterpret_cast<decltype(a.name)>(dlsym(h, "nccl" #name));               \
    if (!a.name) {                                                                     \
        a.error = "libnccl.so.2 lacks nccl" #name;                                     \
        return a;                                                                      \
    }
        SPCU_NCCL_SYM(GetUniqueId)
        SPCU_NCCL_SYM(CommInitRank)
        SPCU_NCCL_SYM(CommInitAll)
        SPCU_NCCL_SYM(CommDestroy)
        SPCU_NCCL_SYM(Reduce)
        SPCU_NCCL_SYM(GroupStart)
        SPCU_NCCL_SYM(GroupEnd)
        SPCU_NCCL_SYM(GetErrorString)
#undef SPCU_NCCL_SYM
        a.ok = true;
        return a;
    }();
    return api;
}

ncclComm_t comm_of(const spcu_ctx* c) { return static_cast<ncclComm_t>(c->nccl_comm); }

} // namespace

#define NEED_NCCL(ctx)                                                          \
    if (!nccl().ok) {                                                           \
        return spcu::fail((ctx), SPCU_ERR_CUDA, "%s", nccl().error.c_str());    \
    }

#define NK(ctx, call)                                                                                                        \
    do {                                                                                                                     \
        const ncclResult_t r_ = (call);                                                                                      \
        if (r_ != ncclSuccess) {                                                                                             \
            return spcu::fail((ctx), SPCU_ERR_CUDA, "%s: %s (%s:%d)", #call, nccl().GetErrorString(r_), __FILE__, __LINE__); \
        }                                                                                                                    \
    } while (0)

extern "C" {

int spcu_comm_unique_id(uint8_t id[SPCU_NCCL_ID_BYTES])
{
    if (!id) {
        return fail(nullptr, SPCU_ERR_INVALID, "id is NULL");
    }
    ncclUniqueId u;
    NEED_NCCL(nullptr);
    NK(nullptr, nccl().GetUniqueId(&u));
    std::memcpy(id, &u, SPCU_NCCL_ID_BYTES);
    return SPCU_OK;
}

void spcu_comm_destroy(spcu_ctx* c)
{
    if (c && c->nccl_comm) {
        cudaSetDevice(c->device);
        nccl().CommDestroy(comm_of(c));
        c->nccl_comm   = nullptr;
        c->comm_rank   = 0;
        c->comm_nranks = 1;
    }
}

int spcu_comm_init_rank(spcu_ctx* c, int nranks, int rank, const uint8_t id[SPCU_NCCL_ID_BYTES])
{
    if (!c || !id || nranks < 1 || rank < 0 || rank >= nranks) {
        return fail(c, SPCU_ERR_INVALID, "spcu_comm_init_rank: bad argument (rank %d of %d)", rank, nranks);
    }
    NEED_NCCL(c);
    spcu_comm_destroy(c);
    CK(c, cudaSetDevice(c->device));
    ncclUniqueId u;
    std::memcpy(&u, id, SPCU_NCCL_ID_BYTES);
    ncclComm_t comm = nullptr;
    NK(c, nccl().CommInitRank(&comm, nranks, u, rank));
    c->nccl_comm   = comm;
    c->comm_rank   = rank;
    c->comm_nranks = nranks;
    return SPCU_OK;
}

int spcu_comm_init_all(spcu_ctx** ctxs, int n)
{
    if (!ctxs || n < 1) {
        return fail(nullptr, SPCU_ERR_INVALID, "spcu_comm_init_all: bad argument");
    }
    NEED_NCCL(nullptr);
    std::vector<int>        devs(n);
    std::vector<ncclComm_t> comms(n, nullptr);
    for (int i = 0; i < n; ++i) {
        if (!ctxs[i]) {
            return fail(nullptr, SPCU_ERR_INVALID, "spcu_comm_init_all: context %d is NULL", i);
        }
        spcu_comm_destroy(ctxs[i]);
        devs[i] = ctxs[i]->device;
    }
    NK(ctxs[0], nccl().CommInitAll(comms.data(), n, devs.data()));
    for (int i = 0; i < n; ++i) {
        ctxs[i]->nccl_comm   = comms[i];
        ctxs[i]->comm_rank   = i;
        ctxs[i]->comm_nranks = n;
    }
    return SPCU_OK;
}

int spcu_comm_rank(const spcu_ctx* c) { return c ? c->comm_rank : -1; }
int spcu_comm_size(const spcu_ctx* c) { return c ? c->comm_nranks : -1; }

int spcu_reduce_to_root(spcu_ctx* c, float* d_rgb_sum, float* d_lum_sumsq, void* stream)
{
    if (int rc = need_scene(c); rc != SPCU_OK) return rc;
    if (!d_rgb_sum) {
        return fail(c, SPCU_ERR_INVALID, "d_rgb_sum is NULL");
    }
    if (c->comm_nranks == 1) {
        return SPCU_OK; // one rank: the sums are already complete
    }
    if (!c->nccl_comm) {
        return fail(c, SPCU_ERR_INVALID, "no communicator: call spcu_comm_init_rank / spcu_comm_init_all first");
    }
    const size_t       n_pixels = static_cast<size_t>(c->ds.width) * c->ds.height;
    const cudaStream_t st       = static_cast<cudaStream_t>(stream);
    // in place: rank 0 receives the sums, the other ranks' buffers are left as they are
    NK(c, nccl().GroupStart());
    NK(c, nccl().Reduce(d_rgb_sum, d_rgb_sum, n_pixels * 3, ncclFloat32, ncclSum, 0, comm_of(c), st));
    if (d_lum_sumsq) {
        NK(c, nccl().Reduce(d_lum_sumsq, d_lum_sumsq, n_pixels, ncclFloat32, ncclSum, 0, comm_of(c), st));
    }
    NK(c, nccl().GroupEnd());
    return SPCU_OK;
}

} // extern "C"
