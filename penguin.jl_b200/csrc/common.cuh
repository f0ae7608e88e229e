// common.cuh -- context, slab-partitioned padded grid, device fields, deterministic fused reductions, NCCL plumbing.
//
// Layout (DESIGN.md "Data layout in HBM"): every per-cell quantity is one contiguous SoA array of `nloc` doubles on
// the reference's padded grid n = prod(n_i+1), x fastest (/root/reference/src/capacity.jl:167-175).  The slowest
// dimension is slab-partitioned over the ranks; every local array carries ONE ghost plane on each side of the slab
// (zeros outside the global domain), so that idx +- stride is always in bounds and -- because every coefficient
// that would couple across a domain border is exactly zero -- stencil kernels need no border branches.
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdint.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/penguin_b200.h"

#define PB_MAXD 3

static thread_local std::string g_last_error;

struct pb200_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    int rank = 0, nranks = 1;
    void *comm = nullptr;  // ncclComm_t
    std::string err;
    int64_t launches = 0;
    int sm_count = 148;
    // reduction scratch
    double *d_partials = nullptr;  // [RED_MAXK][RED_MAXBLOCKS]
    double *d_results = nullptr;   // [RED_SLOTS]
    unsigned *d_counter = nullptr;
    double *h_results = nullptr;   // pinned mirror
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr;
    // second stream of the fused Krylov iteration (fold2.cuh): ghost-class tiles + halo exchange run beside the interior class; kernels
    // on it that reduce use their own scratch (they are concurrent with reducing kernels of the main stream)
    cudaStream_t stream2 = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    double *d_partials2 = nullptr;
    unsigned *d_counter2 = nullptr;
    // one PROCESS driving several GPUs (pb200_init_multi): the team handle (is_team) fans every call out to its member contexts, one host thread
    // per GPU; members know their team so that the peer-memory exchange maps the mailboxes directly (CUDA IPC does not work inside one process)
    struct Team *team = nullptr;
    bool is_team = false;
    // peer-memory exchange (p2p.cuh): every rank's mailbox is mapped into every other rank with CUDA IPC; halos and the Krylov
    // scalar reductions are then plain kernels that store into the peers' mailboxes over NVLink and spin on sequence flags
    struct P2PState *p2p = nullptr;
    int p2p_gen = 0;   // bumped whenever the mailboxes are re-mapped (captured graphs that hold the old pointers are rebuilt)
    // optional per-launch timing of the operator apply (pb200_set_profiling)
    bool profile = false;
    std::vector<cudaEvent_t> pev;   // event pairs
    std::vector<int> ptag;
    size_t pev_used = 0;
    int64_t apply_launches = 0;
};

// event bracket around one launch (no-op unless profiling): prof_mark(ctx, tag); kernel<<<>>>; prof_mark(ctx, tag);
#define PB_PROF_APPLY PB200_K_APPLY
#define PB_PROF_UPDATE PB200_K_UPDATE
#define PB_PROF_PUPD PB200_K_PUPD
#define PB_PROF_BAPPLY PB200_K_BAND_APPLY
#define PB_PROF_BPREC PB200_K_BAND_PREC
#define PB_PROF_XCHG PB200_K_EXCHANGE
#define PB_PROF_PROLOGUE PB200_K_PROLOGUE
#define PB_PROF_EPILOGUE PB200_K_EPILOGUE
#define PB_PROF_NTAG PB200_K_NCLASS
static inline void prof_mark(pb200_ctx *ctx, int tag)
{
    if (!ctx->profile) return;
    if (ctx->pev_used == ctx->pev.size()) {
        cudaEvent_t e;
        cudaEventCreate(&e);
        ctx->pev.push_back(e);
        ctx->ptag.push_back(0);
    }
    ctx->ptag[ctx->pev_used] = tag;
    cudaEventRecord(ctx->pev[ctx->pev_used++], ctx->stream);
}
// summed elapsed time of the recorded pairs per tag (stream must be synchronised); resets the pool
static inline void prof_collect(pb200_ctx *ctx, double ms[PB_PROF_NTAG], int64_t n[PB_PROF_NTAG])
{
    for (int t = 0; t < PB_PROF_NTAG; ++t) { ms[t] = 0.0; n[t] = 0; }
    for (size_t i = 0; i + 1 < ctx->pev_used; i += 2) {
        float t = 0.f;
        cudaEventElapsedTime(&t, ctx->pev[i], ctx->pev[i + 1]);
        ms[ctx->ptag[i]] += t; n[ctx->ptag[i]]++;
    }
    ctx->pev_used = 0;
}

static int set_err(pb200_ctx *ctx, int code, const std::string &msg)
{
    g_last_error = msg;
    if (ctx) ctx->err = msg;
    return code;
}

#define CUDA_TRY(ctx, call)                                                                                          \
    do {                                                                                                             \
        cudaError_t e__ = (call);                                                                                    \
        if (e__ != cudaSuccess)                                                                                      \
            return set_err(ctx, PB200_ECUDA,                                                                         \
                           std::string(#call) + ": " + cudaGetErrorString(e__) + " (" + __FILE__ + ":" +            \
                               std::to_string(__LINE__) + ")");                                                     \
    } while (0)

#define LAUNCH_CHECK(ctx)                                                                                            \
    do {                                                                                                             \
        (ctx)->launches++;                                                                                           \
        cudaError_t e__ = cudaGetLastError();                                                                        \
        if (e__ != cudaSuccess)                                                                                      \
            return set_err(ctx, PB200_ECUDA, std::string("kernel launch: ") + cudaGetErrorString(e__) + " (" +      \
                                                 __FILE__ + ":" + std::to_string(__LINE__) + ")");                   \
    } while (0)

// ------------------------------------------------------------------------------------------------------------
// Grid
// ------------------------------------------------------------------------------------------------------------
struct Grid {
    int N;
    int nc[PB_MAXD];   // real cells per dim (1 for unused dims)
    int pd[PB_MAXD];   // padded dims (n_i + 1; 1 for unused dims)
    double x0[PB_MAXD], h[PB_MAXD], L[PB_MAXD];
    int sd;            // slab dimension = N-1
    int k0, k1;        // owned padded planes [k0,k1) of dim sd (global index)
    int lz;            // local planes incl. the two ghost planes
    int64_t plane;     // cells per plane of dim sd
    int64_t nown;      // owned cells
    int64_t nloc;      // local cells incl. ghosts
    int64_t ntot;      // global padded cells
    int64_t stride[PB_MAXD];
    int rank, nranks;
};

static int make_grid(pb200_ctx *ctx, int ndim, const int *n, const double *x0, const double *L, Grid *g)
{
    if (ndim < 1 || ndim > 3) return set_err(ctx, PB200_EINVAL, "ndim must be 1, 2 or 3");
    memset(g, 0, sizeof(*g));
    g->N = ndim;
    for (int d = 0; d < PB_MAXD; ++d) {
        g->nc[d] = 1; g->pd[d] = 1; g->x0[d] = 0; g->h[d] = 1; g->L[d] = 1;
    }
    for (int d = 0; d < ndim; ++d) {
        if (n[d] < 1) return set_err(ctx, PB200_EINVAL, "n[d] must be >= 1");
        g->nc[d] = n[d]; g->pd[d] = n[d] + 1; g->x0[d] = x0[d]; g->L[d] = L[d];
        g->h[d] = L[d] / n[d];   // same expression as src/mesh.jl:49-50
    }
    g->sd = ndim - 1;
    g->rank = ctx->rank; g->nranks = ctx->nranks;
    int planes = g->pd[g->sd];
    if (planes < ctx->nranks) return set_err(ctx, PB200_EINVAL, "fewer planes than ranks");
    int base = planes / ctx->nranks, rem = planes % ctx->nranks;
    g->k0 = ctx->rank * base + (ctx->rank < rem ? ctx->rank : rem);
    g->k1 = g->k0 + base + (ctx->rank < rem ? 1 : 0);
    g->lz = g->k1 - g->k0 + 2;
    g->plane = 1;
    for (int d = 0; d < g->sd; ++d) g->plane *= g->pd[d];
    g->nown = (int64_t)(g->k1 - g->k0) * g->plane;
    g->nloc = (int64_t)g->lz * g->plane;
    g->ntot = (int64_t)planes * g->plane;
    g->stride[0] = 1; g->stride[1] = g->pd[0]; g->stride[2] = (int64_t)g->pd[0] * g->pd[1];
    return PB200_OK;
}

// coordinates (global) of owned-cell ordinal t (0 <= t < nown); local linear index is t + plane
__device__ __forceinline__ void cell_coords(const Grid &g, int64_t t, int c[PB_MAXD])
{
    int64_t q = t;
    c[0] = (int)(q % g.pd[0]); q /= g.pd[0];
    c[1] = (int)(q % g.pd[1]); q /= g.pd[1];
    c[2] = (int)q;
    c[g.sd] += g.k0;
}

// ------------------------------------------------------------------------------------------------------------
// Device fields
// ------------------------------------------------------------------------------------------------------------
static int dev_alloc(pb200_ctx *ctx, double **p, int64_t n)
{
    CUDA_TRY(ctx, cudaMalloc((void **)p, sizeof(double) * (size_t)n));
    CUDA_TRY(ctx, cudaMemsetAsync(*p, 0, sizeof(double) * (size_t)n, ctx->stream));
    return PB200_OK;
}
static void dev_free(double *&p)
{
    if (p) cudaFree(p);
    p = nullptr;
}
// host (owned cells) -> device (skip the lower ghost plane)
static int upload_owned(pb200_ctx *ctx, const Grid &g, double *dst, const double *src)
{
    if (!src) {
        CUDA_TRY(ctx, cudaMemsetAsync(dst, 0, sizeof(double) * (size_t)g.nloc, ctx->stream));
        return PB200_OK;
    }
    CUDA_TRY(ctx, cudaMemcpyAsync(dst + g.plane, src, sizeof(double) * (size_t)g.nown, cudaMemcpyHostToDevice, ctx->stream));
    return PB200_OK;
}
static int download_owned(pb200_ctx *ctx, const Grid &g, double *dst, const double *src)
{
    if (!dst) return PB200_OK;
    CUDA_TRY(ctx, cudaMemcpyAsync(dst, src + g.plane, sizeof(double) * (size_t)g.nown, cudaMemcpyDeviceToHost, ctx->stream));
    return PB200_OK;
}

// ------------------------------------------------------------------------------------------------------------
// NCCL, loaded lazily with dlopen so that single-GPU use has no libnccl dependency
// ------------------------------------------------------------------------------------------------------------
typedef struct { char internal[128]; } pb_ncclUniqueId;
struct NcclApi {
    void *h = nullptr;
    int (*GetUniqueId)(pb_ncclUniqueId *) = nullptr;
    int (*CommInitRank)(void **, int, pb_ncclUniqueId, int) = nullptr;
    int (*CommDestroy)(void *) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, void *, cudaStream_t) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, void *, cudaStream_t) = nullptr;
    int (*Send)(const void *, size_t, int, int, void *, cudaStream_t) = nullptr;
    int (*Recv)(void *, size_t, int, int, void *, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
};
static NcclApi g_nccl;
static const int PB_NCCL_FLOAT64 = 8, PB_NCCL_SUM = 0, PB_NCCL_MAX = 2, PB_NCCL_UINT8_T = 1;

static int nccl_load(pb200_ctx *ctx)
{
    if (g_nccl.h) return PB200_OK;
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *nm : names) {
        g_nccl.h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (g_nccl.h) break;
    }
    if (!g_nccl.h) return set_err(ctx, PB200_ENCCL, std::string("cannot dlopen libnccl.so.2: ") + dlerror());
#define PB_SYM(field, name)                                                            \
    *(void **)(&g_nccl.field) = dlsym(g_nccl.h, name);                                 \
    if (!g_nccl.field) return set_err(ctx, PB200_ENCCL, std::string("missing symbol ") + name);
    PB_SYM(GetUniqueId, "ncclGetUniqueId")
    PB_SYM(CommInitRank, "ncclCommInitRank")
    PB_SYM(CommDestroy, "ncclCommDestroy")
    PB_SYM(AllReduce, "ncclAllReduce")
    PB_SYM(AllGather, "ncclAllGather")
    PB_SYM(Send, "ncclSend")
    PB_SYM(Recv, "ncclRecv")
    PB_SYM(GroupStart, "ncclGroupStart")
    PB_SYM(GroupEnd, "ncclGroupEnd")
    PB_SYM(GetErrorString, "ncclGetErrorString")
#undef PB_SYM
    return PB200_OK;
}
#define NCCL_TRY(ctx, call)                                                                                          \
    do {                                                                                                             \
        int r__ = (call);                                                                                            \
        if (r__ != 0)                                                                                                \
            return set_err(ctx, PB200_ENCCL, std::string(#call) + ": " + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r__) : "?")); \
    } while (0)

// exchange the ghost planes of `nf` fields with the slab neighbours (one grouped send/recv pair per neighbour)
static int halo_exchange(pb200_ctx *ctx, const Grid &g, double *const *fields, int nf)
{
    if (ctx->nranks == 1) return PB200_OK;
    NCCL_TRY(ctx, g_nccl.GroupStart());
    for (int f = 0; f < nf; ++f) {
        double *p = fields[f];
        if (!p) continue;
        size_t cnt = (size_t)g.plane;
        if (ctx->rank > 0) {
            NCCL_TRY(ctx, g_nccl.Send(p + g.plane, cnt, PB_NCCL_FLOAT64, ctx->rank - 1, ctx->comm, ctx->stream));
            NCCL_TRY(ctx, g_nccl.Recv(p, cnt, PB_NCCL_FLOAT64, ctx->rank - 1, ctx->comm, ctx->stream));
        }
        if (ctx->rank < ctx->nranks - 1) {
            NCCL_TRY(ctx, g_nccl.Send(p + (int64_t)(g.lz - 2) * g.plane, cnt, PB_NCCL_FLOAT64, ctx->rank + 1, ctx->comm, ctx->stream));
            NCCL_TRY(ctx, g_nccl.Recv(p + (int64_t)(g.lz - 1) * g.plane, cnt, PB_NCCL_FLOAT64, ctx->rank + 1, ctx->comm, ctx->stream));
        }
    }
    NCCL_TRY(ctx, g_nccl.GroupEnd());
    return PB200_OK;
}

// ------------------------------------------------------------------------------------------------------------
// Deterministic fused reductions: per-thread partials -> warp shuffle -> block -> last block sums the block
// partials in a fixed order and publishes results[slot .. slot+K).  One launch, no atomics on doubles.
// ------------------------------------------------------------------------------------------------------------
#define RED_MAXK 4
#define RED_MAXBLOCKS 4096
#define RED_SLOTS 32
#define RED_THREADS 256

template <int K>
__device__ __forceinline__ void block_reduce_publish(double (&v)[K], double *__restrict__ partials, double *__restrict__ results,
                                                     unsigned *__restrict__ counter, bool accumulate = false)
{   // accumulate: results[k] += sum (a later launch of the same stream adds its share of one dot product: the order stays fixed)
    __shared__ double sm[RED_MAXK][32];
    __shared__ bool is_last;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int k = 0; k < K; ++k) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_down_sync(0xffffffffu, v[k], o);
        if (lane == 0) sm[k][wid] = v[k];
    }
    __syncthreads();
    if (wid == 0) {
#pragma unroll
        for (int k = 0; k < K; ++k) {
            double s = lane < nw ? sm[k][lane] : 0.0;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
            if (lane == 0) partials[(size_t)k * RED_MAXBLOCKS + blockIdx.x] = s;
        }
    }
    if (threadIdx.x == 0) {
        __threadfence();
        unsigned t = atomicAdd(counter, 1u);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last) {
        __threadfence();
#pragma unroll
        for (int k = 0; k < K; ++k) {
            // four independent accumulators: the loads of a thread are in flight together (the summation order stays fixed)
            double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
            const double *__restrict__ pk = partials + (size_t)k * RED_MAXBLOCKS;
            unsigned i = threadIdx.x;
            for (; i + 3 * blockDim.x < gridDim.x; i += 4 * blockDim.x) {
                s0 += __ldcg(pk + i); s1 += __ldcg(pk + i + blockDim.x); s2 += __ldcg(pk + i + 2 * blockDim.x); s3 += __ldcg(pk + i + 3 * blockDim.x);
            }
            for (; i < gridDim.x; i += blockDim.x) s0 += __ldcg(pk + i);
            double s = (s0 + s1) + (s2 + s3);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
            __syncthreads();
            if (lane == 0) sm[k][wid] = s;
            __syncthreads();
            if (wid == 0) {
                double r = lane < nw ? sm[k][lane] : 0.0;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) r += __shfl_down_sync(0xffffffffu, r, o);
                if (lane == 0) results[k] = accumulate ? results[k] + r : r;
            }
        }
        if (threadIdx.x == 0) *counter = 0u;
    }
}

static inline int red_grid(pb200_ctx *ctx, int64_t n, int per_thread = 1)
{
    int64_t b = (n + (int64_t)RED_THREADS * per_thread - 1) / ((int64_t)RED_THREADS * per_thread);
    int64_t cap = (int64_t)ctx->sm_count * 8;
    if (b > cap) b = cap;
    if (b > RED_MAXBLOCKS) b = RED_MAXBLOCKS;
    if (b < 1) b = 1;
    return (int)b;
}

static int p2p_allreduce(pb200_ctx *ctx, int slot, int K, bool *done);
// sum the K results of slot over the ranks (in stream order)
static int allreduce_results(pb200_ctx *ctx, int slot, int K)
{
    if (ctx->nranks == 1) return PB200_OK;
    bool done = false;
    int rc = p2p_allreduce(ctx, slot, K, &done);
    if (rc || done) return rc;
    NCCL_TRY(ctx, g_nccl.AllReduce(ctx->d_results + slot, ctx->d_results + slot, (size_t)K, PB_NCCL_FLOAT64, PB_NCCL_SUM, ctx->comm, ctx->stream));
    return PB200_OK;
}
// copy results[slot..slot+K) to the host (synchronises the stream)
static int fetch_results(pb200_ctx *ctx, int slot, int K, double *out)
{
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->h_results + slot, ctx->d_results + slot, sizeof(double) * K, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    for (int k = 0; k < K; ++k) out[k] = ctx->h_results[slot + k];
    return PB200_OK;
}
