// bh_mg.cu — multi-GPU host driver of the Morton-slice mode behind the C ABI (include/bh.h "bh_mg_*").
//
// The reference's frame loop (nbody_v5_bench.cu:353-367) calls simulationStep() on one GPU.  Here one PROCESS per
// GPU holds the full state; each step every rank sorts and builds on all bodies (deterministic: identical trees
// without a broadcast), traverses + integrates only its own Morton slice, and the ranks all-gather the updated
// slices in place over NCCL / NVLink (posm 16 B, vel 16 B, ids 4 B per body).  Everything is asynchronous:
//   compute stream   [wait posm] keys + sort, tree, centre of mass, traversal   [wait vel, ids] update   [record done]
//   comm stream      [wait done] all-gather posm [record]  all-gather vel, ids [record]
// Only the kick-drift update reads velocities and ids, so their gathers (20 of the 36 B/body) hide behind the whole
// next step up to its last kernel, and no host thread ever blocks.
//
// NCCL is loaded lazily with dlopen (RTLD_NOLOAD first: a host process that already carries an NCCL — e.g. the
// one bundled with PyTorch — keeps exactly that one); libbh.so itself has no link-time dependency on it.
#include "bh_common.cuh"

#include <dlfcn.h>
#include <nccl.h>

#include <cstdio>
#include <new>

extern "C" int bh_set_slice(bh_ctx* ctx, int rank, int world);

namespace {

struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId*);
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
    ncclResult_t (*CommDestroy)(ncclComm_t);
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t);
    ncclResult_t (*GroupStart)();
    ncclResult_t (*GroupEnd)();
    const char* (*GetErrorString)(ncclResult_t);
    bool ok = false;
};

NcclApi g_nccl;

int load_nccl() {
    if (g_nccl.ok) return 0;
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return BH_E_UNSUPPORTED;
#define BH_SYM(field, name) g_nccl.field = reinterpret_cast<decltype(g_nccl.field)>(dlsym(h, name)); if (!g_nccl.field) return BH_E_UNSUPPORTED
    BH_SYM(GetUniqueId, "ncclGetUniqueId"); BH_SYM(CommInitRank, "ncclCommInitRank"); BH_SYM(CommDestroy, "ncclCommDestroy");
    BH_SYM(AllGather, "ncclAllGather"); BH_SYM(GroupStart, "ncclGroupStart"); BH_SYM(GroupEnd, "ncclGroupEnd");
    BH_SYM(GetErrorString, "ncclGetErrorString");
#undef BH_SYM
    g_nccl.ok = true;
    return 0;
}

}  // namespace

struct bh_mg {
    bh_ctx* ctx = nullptr;
    int rank = 0, world = 1, device = 0;
    ncclComm_t comm = nullptr;
    cudaStream_t comm_stream = nullptr;
    cudaEvent_t ev_compute = nullptr, ev_posm = nullptr, ev_rest = nullptr;
    bool in_flight = false;        // all-gathers of the last step not yet waited for
    int64_t n = 0, per = 0;        // bodies, padded slice length (whole 32-body chunks)
    float4 *posm = nullptr, *vel = nullptr;
    int32_t* ids = nullptr;
};

#define BH_NCCL_TRY(expr)                                                                       \
    do {                                                                                        \
        ncclResult_t _r = (expr);                                                               \
        if (_r != ncclSuccess) {                                                                \
            fprintf(stderr, "bh_mg: %s failed: %s\n", #expr, g_nccl.GetErrorString(_r));        \
            return BH_E_DEVICE;                                                                 \
        }                                                                                       \
    } while (0)

extern "C" {

int bh_mg_unique_id(void* id128) {
    if (!id128) return BH_E_INVAL;
    static_assert(sizeof(ncclUniqueId) == BH_MG_ID_BYTES, "BH_MG_ID_BYTES");
    int e = load_nccl();
    if (e) return e;
    BH_NCCL_TRY(g_nccl.GetUniqueId(reinterpret_cast<ncclUniqueId*>(id128)));
    return 0;
}

// Refreshes the slice geometry from the context (after bh_import_soa* changed n).
static int mg_refresh(bh_mg* m) {
    void *p = nullptr, *v = nullptr, *i = nullptr;
    int64_t n = 0, first = 0, count = 0;
    int e = bh_state_ptrs(m->ctx, &p, &v, &i, &n, &first, &count);
    if (e) return e;
    const int64_t groups = (n + BH_GROUP - 1) / BH_GROUP;
    m->per = (groups + m->world - 1) / m->world * BH_GROUP;   // == default_slice() in bh_engine.cu
    m->n = n;
    m->posm = (float4*)p; m->vel = (float4*)v; m->ids = (int32_t*)i;
    return 0;
}

int bh_mg_create(bh_mg** out, bh_ctx* ctx, const void* id128, int rank, int world, int device) {
    if (!out || !ctx || !id128 || world < 1 || rank < 0 || rank >= world) return BH_E_INVAL;
    *out = nullptr;
    int e = load_nccl();
    if (e) return e;
    BH_CUDA_TRY(cudaSetDevice(device));
    bh_mg* m = new (std::nothrow) bh_mg();
    if (!m) return BH_E_NOMEM;
    m->ctx = ctx; m->rank = rank; m->world = world; m->device = device;
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    ncclResult_t r = g_nccl.CommInitRank(&m->comm, world, id, rank);
    if (r != ncclSuccess) { fprintf(stderr, "bh_mg: ncclCommInitRank: %s\n", g_nccl.GetErrorString(r)); delete m; return BH_E_DEVICE; }
    cudaError_t ce = cudaStreamCreateWithFlags(&m->comm_stream, cudaStreamNonBlocking);
    if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&m->ev_compute, cudaEventDisableTiming);
    if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&m->ev_posm, cudaEventDisableTiming);
    if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&m->ev_rest, cudaEventDisableTiming);
    if (ce != cudaSuccess) { bh_mg_destroy(m); return (int)ce; }
    e = bh_set_slice(ctx, rank, world);
    if (e) { bh_mg_destroy(m); return e; }
    {   // NCCL sets its channels up on the first collective (~1 s): pay that here, not in the first frame
        int* warm = nullptr;
        ce = cudaMalloc(&warm, sizeof(int) * (size_t)world);
        if (ce == cudaSuccess) {
            ncclResult_t r2 = g_nccl.AllGather(warm + rank, warm, 1, ncclInt32, m->comm, m->comm_stream);
            ce = cudaStreamSynchronize(m->comm_stream);
            cudaFree(warm);
            if (r2 != ncclSuccess) { bh_mg_destroy(m); return BH_E_DEVICE; }
        }
        if (ce != cudaSuccess) { bh_mg_destroy(m); return (int)ce; }
    }
    *out = m;
    return 0;
}

void bh_mg_destroy(bh_mg* m) {
    if (!m) return;
    cudaSetDevice(m->device);
    cudaDeviceSynchronize();
    if (m->comm) g_nccl.CommDestroy(m->comm);
    if (m->comm_stream) cudaStreamDestroy(m->comm_stream);
    if (m->ev_compute) cudaEventDestroy(m->ev_compute);
    if (m->ev_posm) cudaEventDestroy(m->ev_posm);
    if (m->ev_rest) cudaEventDestroy(m->ev_rest);
    delete m;
}

// in-place all-gather of this rank's rows [rank*per, (rank+1)*per) of the three state arrays
static int mg_gather(bh_mg* m, cudaStream_t compute) {
    BH_CUDA_TRY(cudaEventRecord(m->ev_compute, compute));
    BH_CUDA_TRY(cudaStreamWaitEvent(m->comm_stream, m->ev_compute, 0));
    const size_t per = (size_t)m->per;
    BH_NCCL_TRY(g_nccl.AllGather(m->posm + per * m->rank, m->posm, per * 4, ncclFloat, m->comm, m->comm_stream));
    BH_CUDA_TRY(cudaEventRecord(m->ev_posm, m->comm_stream));
    BH_NCCL_TRY(g_nccl.GroupStart());
    BH_NCCL_TRY(g_nccl.AllGather(m->vel + per * m->rank, m->vel, per * 4, ncclFloat, m->comm, m->comm_stream));
    BH_NCCL_TRY(g_nccl.AllGather(m->ids + per * m->rank, m->ids, per, ncclInt32, m->comm, m->comm_stream));
    BH_NCCL_TRY(g_nccl.GroupEnd());
    BH_CUDA_TRY(cudaEventRecord(m->ev_rest, m->comm_stream));
    m->in_flight = true;
    return 0;
}

int bh_mg_step(bh_mg* m, int nsteps, void* stream) {
    if (!m || nsteps < 0) return BH_E_INVAL;
    BH_CUDA_TRY(cudaSetDevice(m->device));
    cudaStream_t st = (cudaStream_t)stream;
    int e = mg_refresh(m);
    if (e) return e;
    for (int s = 0; s < nsteps; ++s) {
        if (m->in_flight) BH_CUDA_TRY(cudaStreamWaitEvent(st, m->ev_posm, 0));   // positions + masses of the previous step are complete
        e = bh_step_part(m->ctx, 0, st);                                          // cube, keys, radix sort
        if (!e) e = bh_step_part(m->ctx, 1, st);                                  // tree, centre of mass, traversal: no velocity needed
        if (e) return e;
        if (m->in_flight) BH_CUDA_TRY(cudaStreamWaitEvent(st, m->ev_rest, 0));   // velocities and ids: only the update reads them
        e = bh_step_part(m->ctx, 2, st);
        if (e) return e;
        if (m->world > 1) { e = mg_gather(m, st); if (e) return e; }
    }
    return 0;
}

int bh_mg_finish(bh_mg* m, void* stream) {
    if (!m) return BH_E_INVAL;
    if (m->in_flight) {
        BH_CUDA_TRY(cudaSetDevice(m->device));
        BH_CUDA_TRY(cudaStreamWaitEvent((cudaStream_t)stream, m->ev_posm, 0));
        BH_CUDA_TRY(cudaStreamWaitEvent((cudaStream_t)stream, m->ev_rest, 0));
        m->in_flight = false;
    }
    return 0;
}

int bh_mg_info(bh_mg* m, int* rank, int* world, int64_t* slice_first, int64_t* slice_count, int64_t* padded_per_rank) {
    if (!m) return BH_E_INVAL;
    int e = mg_refresh(m);
    if (e) return e;
    int64_t n = 0, first = 0, count = 0;
    bh_state_ptrs(m->ctx, nullptr, nullptr, nullptr, &n, &first, &count);
    if (rank) *rank = m->rank;
    if (world) *world = m->world;
    if (slice_first) *slice_first = first;
    if (slice_count) *slice_count = count;
    if (padded_per_rank) *padded_per_rank = m->per;
    return 0;
}

}  // extern "C"
