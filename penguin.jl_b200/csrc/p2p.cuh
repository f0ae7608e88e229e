// p2p.cuh -- halo exchange and Krylov scalar reductions through PEER MEMORY (NVLink / NVSwitch) instead of NCCL launches.
//
// One process per GPU.  Every rank owns a MAILBOX in device memory; at set-up the mailboxes are mapped into all other ranks with CUDA IPC
// (the 64-byte handles travel through one ncclAllGather).  An exchange is then ONE small kernel per rank:
//     halo   : store my boundary planes (and the ghost ranges of the compact interface array) into the neighbour's landing zone,
//              __threadfence_system, store the sequence number into the neighbour's flag, spin on my own flag, unpack my landing zone
//              into my ghost planes;
//     reduce : store my K partial sums into slot [my rank] of every peer, flag, spin until all peers have flagged, add the N values in
//              rank order (deterministic, identical on all ranks).
// Sequence numbers live in device memory and are advanced by the kernels themselves, so the kernels can be captured in a CUDA graph and
// replayed; landing zones and reduction slots are double-buffered by sequence parity (a peer can run at most one exchange ahead).
// Latency: one NVLink round trip (~2-4 us) instead of an NCCL launch (~15-20 us) -- the Krylov iterations of a 2048^2 slab take ~100 us.
// A spin gives up after ~2 s and raises an error flag instead of hanging the GPU.  PB200_NO_P2P=1 keeps everything on NCCL.
#pragma once
#include "common.cuh"

#define P2P_MAXR 8
#define P2P_REDK 8
#define P2P_HDR 4096   // bytes reserved for flags, counters and reduction slots at the start of a mailbox

struct P2PHeader {                                   // at offset 0 of every mailbox
    unsigned long long flag_halo[2];                 // [0] written by my LOWER neighbour, [1] by my UPPER neighbour
    unsigned long long flag_red[P2P_MAXR];           // written by peer r
    unsigned long long seq_halo, seq_red;            // my own sequence counters (advanced by my kernels)
    unsigned int done[2];                            // block counters of the halo kernel: [0] blocks that have packed, [1] blocks that have left
    int err;                                         // set when a spin timed out
    int pad;
    double red[2][P2P_MAXR][P2P_REDK];               // [parity][source rank][k]
};
static_assert(sizeof(P2PHeader) <= P2P_HDR, "mailbox header too large");

struct P2PState {
    bool on = false;
    int n = 1, rank = 0;
    char *mbox = nullptr;
    char *peer[P2P_MAXR] = {};
    size_t zone_doubles = 0;   // capacity of ONE landing zone; the mailbox holds 2 (parity) x 2 (direction) of them after the header
    size_t bytes = 0;
};

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// spin until *flag >= seq; false on time-out (~2 s)
__device__ __forceinline__ bool p2p_wait(const unsigned long long *flag, unsigned long long seq)
{
    const long long t0 = clock64();
    while (ld_acquire_sys(flag) < seq) {
        if (clock64() - t0 > 4000000000ll) return false;
        __nanosleep(100);
    }
    return true;
}

struct HaloSeg { const double *src; double *dst; int n; };
struct HaloDirArgs {
    int active;
    int nseg;
    HaloSeg send[3];              // src = my boundary data (dst unused)
    HaloSeg recv[3];              // dst = my ghost storage (src unused)
    char *remote;                 // neighbour's mailbox
    int remote_dir;               // which of ITS zones / flags I write: 1 if I am its lower... see p2p_halo
    int local_dir;                // which of MY zones / flags the neighbour writes
};

// both directions in one launch: lo = exchange with the LOWER neighbour, hi = with the UPPER neighbour
__global__ void k_p2p_halo(char *mbox, size_t zone_doubles, HaloDirArgs lo, HaloDirArgs hi)
{
    P2PHeader *me = (P2PHeader *)mbox;
    // the sequence number of this exchange: every block reads it before anyone advances it (the LAST block to leave advances it)
    const unsigned long long seq = me->seq_halo + 1;
    const int par = (int)(seq & 1);
    const int gtid = blockIdx.x * blockDim.x + threadIdx.x, gsz = gridDim.x * blockDim.x;
#pragma unroll
    for (int dir = 0; dir < 2; ++dir) {
        const HaloDirArgs &D = dir == 0 ? lo : hi;
        if (!D.active) continue;
        double *rz = (double *)(D.remote + P2P_HDR) + ((size_t)par * 2 + D.remote_dir) * zone_doubles;
        size_t off = 0;
        for (int q = 0; q < D.nseg; ++q) {
            const double *__restrict__ src = D.send[q].src;
            for (int i = gtid; i < D.send[q].n; i += gsz) rz[off + i] = src[i];
            off += D.send[q].n;
        }
    }
    __threadfence_system();
    __syncthreads();
    __shared__ int s_ok;
    if (threadIdx.x == 0) {
        const unsigned t = atomicAdd(&me->done[0], 1u);
        if (t == gridDim.x - 1) {   // every block's stores are fenced: tell the neighbours that my data has landed
            if (lo.active) st_release_sys(&((P2PHeader *)lo.remote)->flag_halo[lo.remote_dir], seq);
            if (hi.active) st_release_sys(&((P2PHeader *)hi.remote)->flag_halo[hi.remote_dir], seq);
        }
        s_ok = 1;
        if (lo.active && !p2p_wait(&me->flag_halo[lo.local_dir], seq)) s_ok = 0;
        if (hi.active && !p2p_wait(&me->flag_halo[hi.local_dir], seq)) s_ok = 0;
        if (!s_ok) me->err = 1;
    }
    __syncthreads();
    if (s_ok) {
#pragma unroll
        for (int dir = 0; dir < 2; ++dir) {
            const HaloDirArgs &D = dir == 0 ? lo : hi;
            if (!D.active) continue;
            const double *lz = (const double *)(mbox + P2P_HDR) + ((size_t)par * 2 + D.local_dir) * zone_doubles;
            size_t off = 0;
            for (int q = 0; q < D.nseg; ++q) {
                double *__restrict__ dst = D.recv[q].dst;
                for (int i = gtid; i < D.recv[q].n; i += gsz) dst[i] = __ldcv(lz + off + i);
                off += D.recv[q].n;
            }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        // a SEPARATE counter for the way out: a fast block may leave before a slow one has packed, so one shared counter would let a
        // leaving block take the ticket that releases the flags (or never reach it)
        const unsigned t = atomicAdd(&me->done[1], 1u);
        if (t == gridDim.x - 1) { me->done[0] = 0; me->done[1] = 0; __threadfence(); me->seq_halo = seq; }   // last block out: re-arm and advance
    }
}

// results[slot .. slot+K) <- sum over the ranks (rank order), one block
__global__ void k_p2p_allreduce(char *mbox, char *p0, char *p1, char *p2, char *p3, char *p4, char *p5, char *p6, char *p7, int n, int rank, double *results,
                                int slot, int K)
{
    char *peers[P2P_MAXR] = {p0, p1, p2, p3, p4, p5, p6, p7};
    P2PHeader *me = (P2PHeader *)mbox;
    const unsigned long long seq = me->seq_red + 1;
    const int par = (int)(seq & 1);
    const int tid = threadIdx.x;
    __syncthreads();
    if (tid < n * K) {
        const int r = tid / K, k = tid - r * K;
        ((P2PHeader *)peers[r])->red[par][rank][k] = results[slot + k];
    }
    __threadfence_system();
    __syncthreads();
    __shared__ int s_ok;
    if (tid == 0) s_ok = 1;
    __syncthreads();
    if (tid < n) {
        st_release_sys(&((P2PHeader *)peers[tid])->flag_red[rank], seq);
        if (!p2p_wait(&me->flag_red[tid], seq)) { s_ok = 0; me->err = 1; }
    }
    __syncthreads();
    if (tid < K && s_ok) {
        double sum = 0.0;
        for (int r = 0; r < n; ++r) sum += __ldcv(&me->red[par][r][tid]);
        results[slot + tid] = sum;
    }
    if (tid == 0) me->seq_red = seq;
}

struct Team {
    int n = 0;
    std::vector<pb200_ctx *> ctx;      // member contexts, rank r on devices[r]
    std::vector<char *> mbox;          // mailboxes of the members (same process: mapped directly with peer access)
};

static void p2p_free(pb200_ctx *ctx)
{
    P2PState *P = ctx->p2p;
    if (!P) return;
    if (!ctx->team) for (int r = 0; r < P->n; ++r) if (r != P->rank && P->peer[r]) cudaIpcCloseMemHandle(P->peer[r]);
    if (P->mbox) cudaFree(P->mbox);
    delete P;
    ctx->p2p = nullptr;
}

// (re)create the mailboxes with landing zones of at least `zone_doubles` and map them into every rank.  COLLECTIVE.
static int p2p_setup(pb200_ctx *ctx, size_t zone_doubles)
{
    if (ctx->nranks == 1 || ctx->nranks > P2P_MAXR || getenv("PB200_NO_P2P")) return PB200_OK;
    if (ctx->p2p && !ctx->p2p->on) return PB200_OK;   // tried before, not available (the same on every rank)
    // the zone size must be the same everywhere and so must the decision to (re)build: take the maximum of the requests first
    double want = (double)zone_doubles;
    double *d_tmp = nullptr;
    unsigned char *d_h = nullptr;
    struct Guard { double *&a; unsigned char *&b; ~Guard() { if (a) cudaFree(a); if (b) cudaFree(b); } } guard{d_tmp, d_h};   // also on the error returns
    CUDA_TRY(ctx, cudaMalloc((void **)&d_tmp, sizeof(double)));
    CUDA_TRY(ctx, cudaMemcpyAsync(d_tmp, &want, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    NCCL_TRY(ctx, g_nccl.AllReduce(d_tmp, d_tmp, 1, PB_NCCL_FLOAT64, PB_NCCL_MAX, ctx->comm, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(&want, d_tmp, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->p2p && ctx->p2p->on && (double)ctx->p2p->zone_doubles >= want) return PB200_OK;
    // generous first allocation (>= 1 Mi doubles per zone, twice the request) so that later, larger problems rarely force a re-mapping
    zone_doubles = (size_t)(2.0 * want) + 64;
    if (zone_doubles < (1u << 20)) zone_doubles = 1u << 20;
    p2p_free(ctx);
    ctx->p2p_gen++;
    P2PState *P = new P2PState();
    ctx->p2p = P;
    P->n = ctx->nranks; P->rank = ctx->rank; P->zone_doubles = zone_doubles;
    P->bytes = P2P_HDR + sizeof(double) * 4 * zone_doubles;
    int ok = 1;
    cudaIpcMemHandle_t mine;
    memset(&mine, 0, sizeof(mine));
    if (cudaMalloc((void **)&P->mbox, P->bytes) != cudaSuccess) ok = 0;
    if (ok && cudaMemset(P->mbox, 0, P->bytes) != cudaSuccess) ok = 0;
    if (ok && cudaDeviceSynchronize() != cudaSuccess) ok = 0;   // (default-stream memset vs the non-blocking streams that use the mailbox next)
    if (ok && !ctx->team && cudaIpcGetMemHandle(&mine, P->mbox) != cudaSuccess) ok = 0;
    cudaGetLastError();
    if (ctx->team) {   // one process: publish the raw pointer (the all-gather below is the barrier), enable peer access to the other members' devices
        ctx->team->mbox[P->rank] = ok ? P->mbox : nullptr;
        for (int r = 0; r < P->n && ok; ++r) {
            if (r == P->rank) continue;
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, ctx->device, ctx->team->ctx[r]->device) != cudaSuccess || !can) { ok = 0; break; }
            cudaError_t e = cudaDeviceEnablePeerAccess(ctx->team->ctx[r]->device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) ok = 0;
            cudaGetLastError();
        }
    }
    // all-gather the handles (64 bytes each) and the ok flags through NCCL
    const size_t HS = sizeof(cudaIpcMemHandle_t) + 8;
    std::vector<unsigned char> h_all(HS * P->n, 0);
    CUDA_TRY(ctx, cudaMalloc((void **)&d_h, HS * P->n));
    memcpy(h_all.data() + HS * P->rank, &mine, sizeof(mine));
    h_all[HS * P->rank + sizeof(mine)] = (unsigned char)ok;
    CUDA_TRY(ctx, cudaMemcpyAsync(d_h + HS * P->rank, h_all.data() + HS * P->rank, HS, cudaMemcpyHostToDevice, ctx->stream));
    NCCL_TRY(ctx, g_nccl.AllGather(d_h + HS * P->rank, d_h, HS, PB_NCCL_UINT8_T, ctx->comm, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(h_all.data(), d_h, HS * P->n, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    for (int r = 0; r < P->n; ++r) ok = ok && h_all[HS * r + sizeof(mine)];
    if (ok) {
        for (int r = 0; r < P->n && ok; ++r) {
            if (r == P->rank) { P->peer[r] = P->mbox; continue; }
            if (ctx->team) { P->peer[r] = ctx->team->mbox[r]; if (!P->peer[r]) ok = 0; continue; }
            cudaIpcMemHandle_t hr;
            memcpy(&hr, h_all.data() + HS * r, sizeof(hr));
            void *ptr = nullptr;
            if (cudaIpcOpenMemHandle(&ptr, hr, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = 0; cudaGetLastError(); }
            P->peer[r] = (char *)ptr;
        }
    }
    // everybody must have mapped everybody: one more agreement round
    double okd = ok ? 1.0 : 0.0;
    CUDA_TRY(ctx, cudaMemcpyAsync(d_tmp, &okd, sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    NCCL_TRY(ctx, g_nccl.AllReduce(d_tmp, d_tmp, 1, PB_NCCL_FLOAT64, PB_NCCL_SUM, ctx->comm, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(&okd, d_tmp, sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    P->on = okd > P->n - 0.5;
    if (getenv("PB200_DEBUG")) fprintf(stderr, "[pb200] rank %d: peer-memory exchange %s (zone %zu doubles)\n", ctx->rank, P->on ? "ON" : "off (NCCL)", zone_doubles);
    return PB200_OK;
}

static int p2p_allreduce(pb200_ctx *ctx, int slot, int K, bool *done)
{
    *done = false;
    P2PState *P = ctx->p2p;
    if (!P || !P->on || K > P2P_REDK) return PB200_OK;
    k_p2p_allreduce<<<1, 64, 0, ctx->stream>>>(P->mbox, P->peer[0], P->peer[1], P->peer[2], P->peer[3], P->peer[4], P->peer[5], P->peer[6], P->peer[7], P->n, P->rank,
                                               ctx->d_results, slot, K);
    LAUNCH_CHECK(ctx);
    *done = true;
    return PB200_OK;
}

// host check of the time-out flag (call at points that synchronise anyway)
static int p2p_check(pb200_ctx *ctx)
{
    P2PState *P = ctx->p2p;
    if (!P || !P->on) return PB200_OK;
    int err = 0;
    CUDA_TRY(ctx, cudaMemcpyAsync(&err, P->mbox + offsetof(P2PHeader, err), sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    if (err) {
        // report once: the flag is cleared so that a later solve (after the caller has dealt with the stalled rank) is not failed by it
        cudaMemsetAsync(P->mbox + offsetof(P2PHeader, err), 0, sizeof(int), ctx->stream);
        cudaStreamSynchronize(ctx->stream);
        return set_err(ctx, PB200_ENCCL, "peer-memory exchange timed out waiting for a neighbour rank");
    }
    return PB200_OK;
}
