// dist.cu — row-partitioned multi-GPU path: one process (rank) per GPU, contiguous row blocks, one exchange
// of the updated term slice per iteration.
//
// The reference has no distributed backend at all; its only partitioning precedent is the
// equal-row chunking of parallel_matrix_vector_multiply (ref src/simd_ops.rs:219) and StreamingMatrix
// (ref src/matrix/optimized.rs:485-517), which sb200_partition_rows reproduces.
//
// Per term, on every rank:
//   fused push kernel over the local rows (gathers from the full-length term vector, writes the rank's slice
//   of the next one in place, publishes the local ||t'||^2)
//   -> ncclAllReduce(sum) of the 2 norm doubles + in-place ncclAllGather of the slice (8 n / G bytes sent per rank)
//   -> 1-thread kernel that takes the loop decision from the global norm (identical on every rank).
// NCCL is bound at run time (dlopen) so the library has no link-time dependency on it and shares the copy
// already loaded by the host process (e.g. the one bundled with PyTorch).
#include <dlfcn.h>
#include <nccl.h>

#include <cmath>
#include <cstring>

#include "solver.hpp"

using namespace sb200;

namespace {

struct NcclApi {
    void *handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

NcclApi g_nccl;
std::mutex g_nccl_mu;

int32_t load_nccl() {
    std::lock_guard<std::mutex> lk(g_nccl_mu);
    if (g_nccl.handle) return SB200_OK;
    const char *names[] = {getenv("SUBLINEAR_B200_NCCL"), "libnccl.so.2", "libnccl.so"};
    void *h = nullptr;
    for (const char *nm : names) {
        if (!nm || !*nm) continue;
        h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (h) break;
    }
    if (!h) return fail(SB200_ERR_ALGORITHM, "cannot load NCCL (libnccl.so.2): %s", dlerror());
#define SB_SYM(field, name)                                                                    \
    g_nccl.field = reinterpret_cast<decltype(g_nccl.field)>(dlsym(h, name));                   \
    if (!g_nccl.field) return fail(SB200_ERR_ALGORITHM, "NCCL symbol %s not found", name)
    SB_SYM(GetUniqueId, "ncclGetUniqueId");
    SB_SYM(CommInitRank, "ncclCommInitRank");
    SB_SYM(CommDestroy, "ncclCommDestroy");
    SB_SYM(AllGather, "ncclAllGather");
    SB_SYM(AllReduce, "ncclAllReduce");
    SB_SYM(GroupStart, "ncclGroupStart");
    SB_SYM(GroupEnd, "ncclGroupEnd");
    SB_SYM(GetErrorString, "ncclGetErrorString");
#undef SB_SYM
    g_nccl.handle = h;
    return SB200_OK;
}

#define SB_NCCL(expr)                                                                                      \
    do {                                                                                                   \
        ncclResult_t _r = (expr);                                                                          \
        if (_r != ncclSuccess)                                                                             \
            return fail(SB200_ERR_ALGORITHM, "NCCL error at %s:%d: %s", __FILE__, __LINE__,                \
                        g_nccl.GetErrorString(_r));                                                        \
    } while (0)

}  // namespace

// Exchange arena: one cudaMalloc block per rank, mapped by every other rank through CUDA IPC.
//   [0, 256)        flags  : world x u64, flags[r] = last epoch signalled by rank r
//   [256, 1024)     slots  : [2 parities][world][2] doubles, partial sums of the current exchange
//   [1024, ...)     T0 | T1 | X : three full-length vectors (cap doubles each): term ping-pong + solution mirror
struct PeerArena {
    void *base = nullptr;
    uint64_t cap = 0;  // doubles per vector
    void *peer[kMaxPeers] = {nullptr};
    unsigned long long *flags(int p) const { return reinterpret_cast<unsigned long long *>(peer[p]); }
    double *slots(int p) const { return reinterpret_cast<double *>(static_cast<char *>(peer[p]) + 256); }
    double *vec(int p, int which) const {
        return reinterpret_cast<double *>(static_cast<char *>(peer[p]) + 1024) + (size_t)which * cap;
    }
};

struct sb200_comm {
    int rank = 0, world = 1, device = 0;
    ncclComm_t comm = nullptr;
    cudaStream_t stream = nullptr;
    bool p2p = false;  // exchange over IPC-mapped peer memory (default when world <= 8); else NCCL collectives
    PeerArena arena;
    unsigned long long epoch_base = 0;
};

namespace {

uint64_t rows_per_rank(uint64_t n, int world) { return (n + (uint64_t)world - 1) / (uint64_t)world; }

// The exchange after a push: global norms + the new term slice. `tfull` has world * per doubles.
int32_t exchange_term(sb200_comm *c, LoopCtl *ctl, double *tfull, uint64_t per, cudaStream_t st) {
    if (c->world == 1) return SB200_OK;
    SB_NCCL(g_nccl.GroupStart());
    SB_NCCL(g_nccl.AllReduce(ctl->red, ctl->red, 2, ncclDouble, ncclSum, c->comm, st));
    SB_NCCL(g_nccl.AllGather(tfull + (uint64_t)c->rank * per, tfull, per, ncclDouble, c->comm, st));
    SB_NCCL(g_nccl.GroupEnd());
    return SB200_OK;
}

struct DistPlan {
    uint64_t n = 0, per = 0, nloc = 0, row0 = 0;
};

int32_t dist_check(sb200_comm *c, const sb200_matrix *m, uint64_t nlocal, DistPlan &p) {
    if (!c || !m) return fail(SB200_ERR_INVALID_INPUT, "null comm or matrix");
    if (!m->distributed) return fail(SB200_ERR_INVALID_INPUT, "matrix was not created by sb200_dist_matrix_from_csr");
    p.n = m->n_global;
    p.per = rows_per_rank(p.n, c->world);
    p.nloc = m->nrows;
    p.row0 = m->row_base;
    if (p.row0 != std::min<uint64_t>(p.n, (uint64_t)c->rank * p.per))
        return fail(SB200_ERR_INVALID_INPUT, "row block does not start at this rank's partition boundary");
    if (nlocal != p.nloc)
        return fail(SB200_ERR_DIMENSION_MISMATCH, "expected %llu, actual %llu in neumann_initialization (local rows)",
                    (unsigned long long)p.nloc, (unsigned long long)nlocal);
    return SB200_OK;
}

// Every rank learns whether any rank failed a local check (codes are small positive ints: max-reduce).
int32_t agree_status(sb200_comm *c, int32_t local, Workspace &ws, cudaStream_t st) {
    if (c->world == 1) return local;
    double v = (double)local, *d = ws.partials.p;
    SB_CUDA(cudaMemcpyAsync(d, &v, 8, cudaMemcpyHostToDevice, st));
    SB_NCCL(g_nccl.AllReduce(d, d, 1, ncclDouble, ncclMax, c->comm, st));
    SB_CUDA(cudaMemcpyAsync(&v, d, 8, cudaMemcpyDeviceToHost, st));
    SB_CUDA(cudaStreamSynchronize(st));
    return (int32_t)v;
}


// (Re)allocate the exchange arena for vectors of `nfull` doubles and map every peer's. Collective.
int32_t ensure_arena(sb200_comm *c, uint64_t nfull) {
    PeerArena &A = c->arena;
    if (A.cap >= nfull && A.base) return SB200_OK;
    cudaStream_t st = c->stream;
    DevBuf<double> bar;
    SB_TRY(bar.alloc(1));
    if (A.base) {  // growth: unmap peers, make sure everybody did, then free
        for (int p = 0; p < c->world; p++)
            if (p != c->rank && A.peer[p]) cudaIpcCloseMemHandle(A.peer[p]);
        SB_NCCL(g_nccl.AllReduce(bar.p, bar.p, 1, ncclDouble, ncclSum, c->comm, st));
        SB_CUDA(cudaStreamSynchronize(st));
        cudaFree(A.base);
        A.base = nullptr;
        A.cap = 0;
    }
    const uint64_t cap = (nfull + 31) & ~31ull;
    const size_t bytes = 1024 + (size_t)3 * cap * sizeof(double);
    cudaError_t e = cudaMalloc(&A.base, bytes);
    if (e != cudaSuccess) {
        A.base = nullptr;
        cudaGetLastError();
    }
    {   // every rank learns whether any allocation failed before anybody waits in the handle exchange
        const double mine_failed = e != cudaSuccess ? 1.0 : 0.0;
        double any = 0.0;
        SB_CUDA(cudaMemcpyAsync(bar.p, &mine_failed, 8, cudaMemcpyHostToDevice, st));
        SB_NCCL(g_nccl.AllReduce(bar.p, bar.p, 1, ncclDouble, ncclMax, c->comm, st));
        SB_CUDA(cudaMemcpyAsync(&any, bar.p, 8, cudaMemcpyDeviceToHost, st));
        SB_CUDA(cudaStreamSynchronize(st));
        if (any != 0.0) {
            if (A.base) cudaFree(A.base);
            A.base = nullptr;
            return fail(SB200_ERR_MEMORY_ALLOCATION, "cudaMalloc of the %zu-byte exchange arena failed on %s: %s", bytes,
                        e != cudaSuccess ? "this rank" : "another rank", cudaGetErrorString(e));
        }
    }
    SB_CUDA(cudaMemset(A.base, 0, bytes));
    SB_CUDA(cudaDeviceSynchronize());
    cudaIpcMemHandle_t mine;
    SB_CUDA(cudaIpcGetMemHandle(&mine, A.base));
    DevBuf<char> hb;
    SB_TRY(hb.alloc((size_t)c->world * sizeof(cudaIpcMemHandle_t)));
    SB_CUDA(cudaMemcpyAsync(hb.p + (size_t)c->rank * sizeof(mine), &mine, sizeof(mine), cudaMemcpyHostToDevice, st));
    SB_NCCL(g_nccl.AllGather(hb.p + (size_t)c->rank * sizeof(mine), hb.p, sizeof(mine), ncclChar, c->comm, st));
    std::vector<cudaIpcMemHandle_t> all(c->world);
    SB_CUDA(cudaMemcpyAsync(all.data(), hb.p, (size_t)c->world * sizeof(mine), cudaMemcpyDeviceToHost, st));
    SB_CUDA(cudaStreamSynchronize(st));
    for (int p = 0; p < c->world; p++) {
        if (p == c->rank) {
            A.peer[p] = A.base;
            continue;
        }
        e = cudaIpcOpenMemHandle(&A.peer[p], all[p], cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            cudaGetLastError();
            return fail(SB200_ERR_ALGORITHM, "cudaIpcOpenMemHandle for rank %d failed: %s (set SUBLINEAR_B200_DIST=nccl)", p,
                        cudaGetErrorString(e));
        }
    }
    A.cap = cap;
    return SB200_OK;
}

PeerExchange make_px(const sb200_comm *c, int t_which, bool publish_x) {
    PeerExchange px{};
    px.world = c->world;
    px.rank = c->rank;
    px.epoch_base = c->epoch_base;
    for (int p = 0; p < c->world; p++) {
        px.t_out[p] = t_which >= 0 ? c->arena.vec(p, t_which) : nullptr;
        px.x_out[p] = publish_x ? c->arena.vec(p, 2) : nullptr;
        px.slots[p] = c->arena.slots(p);
        px.flags[p] = c->arena.flags(p);
    }
    return px;
}

// The P2P flavour of sb200_dist_push_iterations_dev / sb200_dist_solve: same control flow as the NCCL flavour below,
// the exchange is fused into the kernels (see PeerExchange in common.hpp).
int32_t push_iterations_p2p(sb200_comm *c, sb200_matrix *mm, const DistPlan &p, Workspace &ws, const double *b_local_dev,
                            uint64_t nterms, double *x, cudaStream_t st) {
    SB_TRY(ensure_arena(c, p.per * c->world));
    c->epoch_base += 1ull << 24;  // a fresh epoch range per call: flags left by earlier calls can never satisfy a wait
    const int cfg = -1;
    LoopCtl h{};
    h.res_norm = INFINITY;
    h.alive = 1;
    h.max_terms = h.max_iterations = 0xFFFFFFFFu;
    *ws.h_ctl = h;
    SB_CUDA(cudaMemcpyAsync(ws.ctl.p, ws.h_ctl, sizeof(LoopCtl), cudaMemcpyHostToDevice, st));
    InitArgs ia{};
    ia.b = b_local_dev;
    ia.dinv = mm->d_dinv[0].p;
    ia.t_out = c->arena.vec(c->rank, 0) + p.row0;
    ia.x_out = x;
    ia.n = (uint32_t)p.nloc;
    ia.ctl = ws.ctl.p;
    ia.partials = ws.partials.p;
    ia.row_base = (uint32_t)p.row0;
    ia.norm_log = ws.norm_log.p;
    ia.px = make_px(c, 0, false);
    SB_TRY(launch_init_state(ia, st));
    TileKernelArgs base{};
    fill_tile_args(mm, base);
    base.ctl = ws.ctl.p;
    base.partials = ws.partials.p;
    base.acc = ws.tmp.p;  // column-slab passes: partial row sums
    base.force = 1;
    base.sol = x;
    base.dinv = mm->d_dinv[0].p;
    base.norm_log = ws.norm_log.p;
    SB_CUDA(cudaEventRecord(ws.ev0, st));
    for (uint64_t it = 1; it <= nterms; it++) {
        TileKernelArgs a = base;
        a.xin = c->arena.vec(c->rank, (int)((it - 1) & 1));
        a.xin_own = a.xin + p.row0;
        a.out = c->arena.vec(c->rank, (int)(it & 1)) + p.row0;
        a.it = (uint32_t)it;
        a.px = make_px(c, (int)(it & 1), false);
        if (getenv("SUBLINEAR_B200_DEBUG_NOSTORE"))  // timing aid: compute + signalling only, no remote term stores
            for (int q = 0; q < c->world; q++) a.px.t_out[q] = nullptr;
        SB_TRY(launch_tile_kernel(cfg, EPI_PUSH, a, st));
    }
    SB_CUDA(cudaEventRecord(ws.ev1, st));
    return SB200_OK;
}

}  // namespace

extern "C" {

int32_t sb200_comm_unique_id(uint8_t id[SB200_UNIQUE_ID_BYTES]) {
    clear_error();
    if (!id) return fail(SB200_ERR_INVALID_INPUT, "id is null");
    SB_TRY(load_nccl());
    static_assert(sizeof(ncclUniqueId) == SB200_UNIQUE_ID_BYTES, "ncclUniqueId size");
    ncclUniqueId u;
    SB_NCCL(g_nccl.GetUniqueId(&u));
    memcpy(id, &u, sizeof(u));
    return SB200_OK;
}

int32_t sb200_comm_init(int32_t rank, int32_t world, const uint8_t id[SB200_UNIQUE_ID_BYTES], int32_t device,
                        sb200_comm **out) {
    clear_error();
    if (!out || !id) return fail(SB200_ERR_INVALID_INPUT, "null argument");
    *out = nullptr;
    if (world < 1 || rank < 0 || rank >= world) return fail(SB200_ERR_INVALID_INPUT, "bad rank %d / world %d", rank, world);
    SB_TRY(sb200_set_device(device));
    SB_TRY(load_nccl());
    std::unique_ptr<sb200_comm> c(new sb200_comm());
    c->rank = rank;
    c->world = world;
    c->device = device;
    ncclUniqueId u;
    memcpy(&u, id, sizeof(u));
    SB_NCCL(g_nccl.CommInitRank(&c->comm, world, u, rank));
    SB_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    const char *mode = getenv("SUBLINEAR_B200_DIST");  // "nccl" forces the collective-based exchange
    c->p2p = world > 1 && world <= kMaxPeers && !(mode && std::string(mode) == "nccl");
    *out = c.release();
    return SB200_OK;
}

void sb200_comm_free(sb200_comm *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    for (int p = 0; p < c->world; p++)
        if (p != c->rank && c->arena.peer[p]) cudaIpcCloseMemHandle(c->arena.peer[p]);
    if (c->arena.base) cudaFree(c->arena.base);
    if (c->stream) cudaStreamDestroy(c->stream);
    if (c->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm);
    delete c;
}

int32_t sb200_partition_rows(uint64_t nrows, int32_t world, int32_t rank, uint64_t *row0, uint64_t *row1) {
    if (world < 1 || rank < 0 || rank >= world || !row0 || !row1)
        return fail(SB200_ERR_INVALID_INPUT, "bad partition arguments");
    const uint64_t per = rows_per_rank(nrows, world);  // chunk_size = (rows + threads - 1) / threads (simd_ops.rs:219)
    *row0 = std::min<uint64_t>(nrows, (uint64_t)rank * per);
    *row1 = std::min<uint64_t>(nrows, (uint64_t)(rank + 1) * per);
    return SB200_OK;
}

int32_t sb200_dist_matrix_from_csr(sb200_comm *c, uint64_t n_global, uint64_t row0, uint64_t row1,
                                   const uint64_t *row_ptr, const uint32_t *col_indices, const double *values,
                                   sb200_matrix **out) {
    clear_error();
    if (!c || !out || !row_ptr) return fail(SB200_ERR_INVALID_INPUT, "null argument");
    *out = nullptr;
    uint64_t e0, e1;
    SB_TRY(sb200_partition_rows(n_global, c->world, c->rank, &e0, &e1));
    if (row0 != e0 || row1 != e1)
        return fail(SB200_ERR_INVALID_INPUT, "rank %d owns rows [%llu, %llu), got [%llu, %llu)", c->rank,
                    (unsigned long long)e0, (unsigned long long)e1, (unsigned long long)row0, (unsigned long long)row1);
    SB_TRY(sb200_set_device(c->device));
    const uint64_t nloc = row1 - row0;
    SB_TRY(matrix_from_host_csr(row_ptr, nullptr, col_indices, values, nloc, n_global, row_ptr[nloc], true, out));
    (*out)->distributed = true;
    (*out)->tile_cfg = -1;  // the fused exchange lives in the warp-stream kernel
    (*out)->row_base = row0;
    (*out)->n_global = n_global;
    return SB200_OK;
}

int32_t sb200_dist_push_iterations_dev(sb200_comm *c, const sb200_matrix *m, const double *b_local_dev, uint64_t nlocal,
                                       uint64_t nterms, double *x_local_dev, double *term_norms, float *elapsed_ms) {
    clear_error();
    DistPlan p;
    SB_TRY(dist_check(c, m, nlocal, p));
    DeviceGuard g(m->device);
    sb200_matrix *mm = const_cast<sb200_matrix *>(m);
    cudaStream_t st = c->stream;
    auto ws = matrix_acquire_ws(mm);
    struct Release {
        sb200_matrix *m;
        std::unique_ptr<Workspace> &ws;
        ~Release() { matrix_release_ws(m, std::move(ws)); }
    } rel{mm, ws};
    const int cfg = m->tile_cfg;
    const size_t npart = 2 * (size_t)std::max(tile_kernel_max_grid(cfg, EPI_PUSH), init_state_grid()) + 2;
    SB_TRY(ws->ensure(p.nloc, p.per * c->world, npart));
    if (ws->norm_log.n < nterms + 1) SB_TRY(ws->norm_log.alloc(nterms + 1));
    int32_t local_rc = matrix_analyse(mm, SB200_MODE_CORRECT, false);
    if (local_rc == SB200_OK && m->first_bad_diag[0] != kNone) local_rc = SB200_ERR_INVALID_SPARSE_MATRIX;
    const int32_t rc = agree_status(c, local_rc, *ws, st);
    if (rc != SB200_OK) return local_rc != SB200_OK ? local_rc : fail(rc, "another rank rejected its row block");

    double *x = x_local_dev ? x_local_dev : ws->x.p;
    if (c->p2p) {
        SB_TRY(push_iterations_p2p(c, mm, p, *ws, b_local_dev, nterms, x, st));
        std::vector<double> plog(nterms + 1);
        SB_CUDA(cudaMemcpyAsync(plog.data(), ws->norm_log.p, (nterms + 1) * 8, cudaMemcpyDeviceToHost, st));
        SB_CUDA(cudaStreamSynchronize(st));
        if (elapsed_ms) SB_CUDA(cudaEventElapsedTime(elapsed_ms, ws->ev0, ws->ev1));
        if (term_norms)
            for (uint64_t k = 0; k < nterms; k++) term_norms[k] = std::sqrt(plog[k + 1]);
        return SB200_OK;
    }
    LoopCtl h{};
    h.res_norm = INFINITY;
    h.alive = 1;
    h.max_terms = h.max_iterations = 0xFFFFFFFFu;
    *ws->h_ctl = h;
    SB_CUDA(cudaMemcpyAsync(ws->ctl.p, ws->h_ctl, sizeof(LoopCtl), cudaMemcpyHostToDevice, st));
    InitArgs ia{};
    ia.b = b_local_dev;
    ia.dinv = m->d_dinv[0].p;
    ia.t_out = ws->t[0].p + p.row0;
    ia.x_out = x;
    ia.n = (uint32_t)p.nloc;
    ia.ctl = ws->ctl.p;
    ia.partials = ws->partials.p;
    ia.defer_tail = c->world > 1;
    ia.norm_log = ws->norm_log.p;
    SB_TRY(launch_init_state(ia, st));
    SB_TRY(exchange_term(c, ws->ctl.p, ws->t[0].p, p.per, st));
    if (c->world > 1) SB_TRY(launch_dist_tail(ws->ctl.p, 1, 0, 0, 0, 1, ws->norm_log.p, st));
    TileKernelArgs base{};
    fill_tile_args(m, base);
    base.ctl = ws->ctl.p;
    base.partials = ws->partials.p;
    base.acc = ws->tmp.p;
    base.force = 1;
    base.sol = x;
    base.dinv = m->d_dinv[0].p;
    base.defer_tail = c->world > 1;
    base.norm_log = ws->norm_log.p;
    SB_CUDA(cudaEventRecord(ws->ev0, st));
    for (uint64_t it = 1; it <= nterms; it++) {
        TileKernelArgs a = base;
        a.xin = ws->t[(it - 1) & 1].p;
        a.xin_own = a.xin + p.row0;
        a.out = ws->t[it & 1].p + p.row0;
        a.it = (uint32_t)it;
        SB_TRY(launch_tile_kernel(cfg, EPI_PUSH, a, st));
        SB_TRY(exchange_term(c, ws->ctl.p, ws->t[it & 1].p, p.per, st));
        if (c->world > 1) SB_TRY(launch_dist_tail(ws->ctl.p, 1, (uint32_t)it, 0, 0, 1, ws->norm_log.p, st));
    }
    SB_CUDA(cudaEventRecord(ws->ev1, st));
    std::vector<double> log(nterms + 1);
    SB_CUDA(cudaMemcpyAsync(log.data(), ws->norm_log.p, (nterms + 1) * 8, cudaMemcpyDeviceToHost, st));
    SB_CUDA(cudaStreamSynchronize(st));
    if (elapsed_ms) SB_CUDA(cudaEventElapsedTime(elapsed_ms, ws->ev0, ws->ev1));
    if (term_norms)
        for (uint64_t k = 0; k < nterms; k++) term_norms[k] = std::sqrt(log[k + 1]);
    return SB200_OK;
}

// Distributed NeumannSolver::solve: the control flow of solver.cu::solve_device with the exchange after every
// term kernel and an allgather of the solution before every residual kernel (the residual A x - rhs needs all of x).
int32_t sb200_dist_solve(sb200_comm *c, const sb200_solver *s, const sb200_matrix *m, const double *b_local,
                         uint64_t nlocal, const sb200_options *opt, double *x_local, sb200_result *out) {
    clear_error();
    if (!out) return fail(SB200_ERR_INVALID_INPUT, "result is null");
    memset(out, 0, sizeof(*out));
    // without a communicator / matrix nothing can be agreed on: these are the only rank-local returns. Every other
    // precondition is folded into local_rc and returned only after agree_status(), so that a rank that rejects its
    // arguments never leaves the others blocked in a collective or spinning on its flags.
    if (!c || !m) return fail(SB200_ERR_INVALID_INPUT, "null comm or matrix");
    DeviceGuard g(m->device);
    sb200_matrix *mm = const_cast<sb200_matrix *>(m);
    cudaStream_t st = c->stream;
    DistPlan p;
    int32_t local_rc = SB200_OK;
    auto check = [&](int32_t rc) { if (local_rc == SB200_OK) local_rc = rc; };
    if (!s) check(fail(SB200_ERR_INVALID_INPUT, "null solver"));
    else check(validate_options(opt));
    if (local_rc == SB200_OK) check(dist_check(c, m, nlocal, p));
    if (local_rc == SB200_OK && nlocal && (!b_local || !x_local)) check(fail(SB200_ERR_INVALID_INPUT, "null vector"));
    if (local_rc == SB200_OK && opt->initial_guess && opt->initial_guess_len != nlocal)
        check(fail(SB200_ERR_DIMENSION_MISMATCH, "expected %llu, actual %llu in initial_guess (local rows)",
                   (unsigned long long)nlocal, (unsigned long long)opt->initial_guess_len));
    const uint64_t max_it = (opt && local_rc == SB200_OK) ? opt->max_iterations : 1, max_terms = s ? s->max_terms : 1;
    if (local_rc == SB200_OK && (max_it >= 0xFFFFFFFFull || max_terms >= 0xFFFFFFFFull || max_it == 0 || max_terms == 0))
        check(fail(SB200_ERR_INVALID_INPUT, "max_iterations / max_terms must be in [1, 2^32)"));
    const bool args_ok = local_rc == SB200_OK;
    const bool compat = args_ok && opt->mode == SB200_MODE_REF_COMPAT;
    const bool identity = args_ok && opt->residual_check == SB200_RESIDUAL_IDENTITY;
    const bool multi = c->world > 1;
    const bool p2p = multi && c->p2p;

    auto ws = matrix_acquire_ws(mm);
    struct Release {
        sb200_matrix *m;
        std::unique_ptr<Workspace> &ws;
        ~Release() { matrix_release_ws(m, std::move(ws)); }
    } rel{mm, ws};
    const int cfg = m->tile_cfg;
    const size_t npart = 2 * (size_t)std::max(std::max(tile_kernel_max_grid(cfg, EPI_PUSH), tile_kernel_max_grid(cfg, EPI_RESID)),
                                              init_state_grid()) + 2;
    if (args_ok) check(ws->ensure(p.nloc, p2p ? 1 : p.per * c->world, npart));
    else SB_TRY(ws->ensure(1, 1, 4));  // the agreement below only needs one scratch double
    {
        const int32_t rc = agree_status(c, local_rc, *ws, st);
        if (rc != SB200_OK) return local_rc != SB200_OK ? local_rc : fail(rc, "another rank rejected its arguments");
    }
    if (p2p) {
        SB_TRY(ensure_arena(c, p.per * c->world));  // collective; allocation failures are agreed on inside
        c->epoch_base += 1ull << 24;  // a fresh epoch range per solve
    }
    // term ping-pong and the full-length solution mirror: arena vectors (P2P) or workspace buffers (NCCL)
    double *const T[2] = {p2p ? c->arena.vec(c->rank, 0) : ws->t[0].p, p2p ? c->arena.vec(c->rank, 1) : ws->t[1].p};

    // local checks of NeumannState::new, then agree across ranks so that nobody blocks in a collective. Column
    // dominance (SB200_DOMINANCE_ROW_OR_COL, the PageRank systems) needs whole columns: the per-column sums of the row
    // blocks are all-reduced before the test, and "row dominant" has to hold on every rank.
    const bool cols = opt->dominance == SB200_DOMINANCE_ROW_OR_COL;
    ColReduce reduce_cols;
    if (cols && multi)
        reduce_cols = [&](double *cd, double *co, uint64_t nc, cudaStream_t s2) -> int32_t {
            SB_NCCL(g_nccl.GroupStart());
            SB_NCCL(g_nccl.AllReduce(cd, cd, nc, ncclDouble, ncclSum, c->comm, s2));
            SB_NCCL(g_nccl.AllReduce(co, co, nc, ncclDouble, ncclSum, c->comm, s2));
            SB_NCCL(g_nccl.GroupEnd());
            return SB200_OK;
        };
    local_rc = matrix_analyse(mm, opt->mode, cols, reduce_cols);
    if (cols) {
        // 2 = a row of this block violates row dominance; the column verdict is already global
        const int32_t any_bad_row = agree_status(c, (local_rc == SB200_OK && m->first_bad_dd != kNone) ? 2 : 0, *ws, st);
        if (local_rc == SB200_OK && any_bad_row != 0 && m->first_bad_col != kNone)
            local_rc = fail(SB200_ERR_MATRIX_NOT_DIAGONALLY_DOMINANT,
                            "matrix is neither row nor column diagonally dominant (first violating column %llu)",
                            (unsigned long long)m->first_bad_col);
    } else if (local_rc == SB200_OK && m->first_bad_dd != kNone) {
        local_rc = fail(SB200_ERR_MATRIX_NOT_DIAGONALLY_DOMINANT, "matrix is not diagonally dominant (first violating row %llu)",
                        (unsigned long long)(m->first_bad_dd + p.row0));
    }
    if (local_rc == SB200_OK && m->first_bad_diag[opt->mode] != kNone)
        local_rc = fail(SB200_ERR_INVALID_SPARSE_MATRIX, "Missing or near-zero diagonal element at position %llu",
                        (unsigned long long)(m->first_bad_diag[opt->mode] + p.row0));
    DevBuf<double> x0;
    if (local_rc == SB200_OK && opt->initial_guess) {
        local_rc = x0.alloc(p.nloc);
        if (local_rc == SB200_OK) local_rc = copy_h2d(x0.p, opt->initial_guess, p.nloc * 8, st);
    }
    {
        const int32_t rc = agree_status(c, local_rc, *ws, st);
        if (rc != SB200_OK) return local_rc != SB200_OK ? local_rc : fail(rc, "another rank rejected its row block");
    }

    SB_TRY(copy_h2d(ws->b.p, b_local, p.nloc * 8, st));
    LoopCtl h{};
    h.res_norm = INFINITY;
    h.tolerance = opt->tolerance;
    h.series_tolerance = s->series_tolerance;
    h.max_terms = (uint32_t)max_terms;
    h.max_iterations = (uint32_t)max_it;
    h.alive = 1;
    *ws->h_ctl = h;
    SB_CUDA(cudaMemcpyAsync(ws->ctl.p, ws->h_ctl, sizeof(LoopCtl), cudaMemcpyHostToDevice, st));

    uint64_t launches = 0;
    const double *dinv = m->d_dinv[opt->mode].p;
    const double *resid_rhs = compat ? ws->c.p : ws->b.p;
    TileKernelArgs base{};
    fill_tile_args(m, base);
    base.ctl = ws->ctl.p;
    base.partials = ws->partials.p;
    base.acc = ws->tmp.p;
    base.identity_res = identity;
    base.defer_tail = multi && !p2p;

    // residual over the local rows. NCCL: allgather x into the term buffer that is dead at this point.
    // P2P: x was published into every rank's solution mirror by the kernel that produced it (x_published), or is
    // published here by a copy kernel (after the loop ended / in the spin phase).
    auto enqueue_resid = [&](uint64_t it, double *scratch_full, int last, int force, bool x_published) -> int32_t {
        const double *xfull = ws->x.p - p.row0;
        if (p2p) {
            if (!x_published) {
                double *dst[kMaxPeers];
                for (int q = 0; q < c->world; q++) dst[q] = c->arena.vec(q, 2);
                SB_TRY(launch_peer_publish(ws->x.p, p.nloc, p.row0, dst, ws->ctl.p, make_px(c, -1, false), force, st));
                launches++;
            }
            xfull = c->arena.vec(c->rank, 2);
        } else if (multi) {
            SB_CUDA(cudaMemcpyAsync(scratch_full + p.row0, ws->x.p, p.nloc * 8, cudaMemcpyDeviceToDevice, st));
            SB_NCCL(g_nccl.AllGather(scratch_full + (uint64_t)c->rank * p.per, scratch_full, p.per, ncclDouble, c->comm, st));
            xfull = scratch_full;
        }
        TileKernelArgs a = base;
        a.xin = xfull;
        a.xin_own = ws->x.p;
        a.rhs = resid_rhs;
        a.it = (uint32_t)it;
        a.last_in_iter = last;
        a.force = force;
        a.identity_res = 0;
        if (p2p) a.px = make_px(c, -1, false);
        launches++;
        SB_TRY(launch_tile_kernel(cfg, EPI_RESID, a, st));
        if (!p2p && multi) {
            SB_NCCL(g_nccl.AllReduce(ws->ctl.p->red, ws->ctl.p->red, 1, ncclDouble, ncclSum, c->comm, st));
            SB_TRY(launch_dist_tail(ws->ctl.p, 2, (uint32_t)it, last, 0, force, nullptr, st));
        }
        return SB200_OK;
    };
    auto read_ctl = [&]() -> int32_t {
        SB_CUDA(cudaMemcpyAsync(ws->h_ctl, ws->ctl.p, sizeof(LoopCtl), cudaMemcpyDeviceToHost, st));
        SB_CUDA(cudaStreamSynchronize(st));
        return SB200_OK;
    };

    // per-launch events of the push kernels (options.enable_profiling), as in solver.cu
    struct ProfEv {
        cudaEvent_t e0, e1;
        uint64_t it;
    };
    std::vector<ProfEv> prof;
    struct ProfCleanup {
        std::vector<ProfEv> &v;
        ~ProfCleanup() {
            for (auto &q : v) { cudaEventDestroy(q.e0); cudaEventDestroy(q.e1); }
        }
    } prof_cleanup{prof};
    const bool profiling = opt->enable_profiling != 0;

    // initial guess (local rows). ref_compat: x = x0 + c, the series still starts from c (neumann.rs:197-211).
    // correct: t0 = D^-1 (b - A x0) needs all of x0 on every rank: the slices are exchanged into the term buffer that
    // stays dead until iteration 1 writes it, then one SpMV over the local rows.
    uint64_t extra_matvec = 0;
    const double *ax0 = nullptr;
    if (x0.p && !compat) {
        const double *x0full = x0.p - p.row0;
        if (p2p) {
            double *dst[kMaxPeers];
            for (int q = 0; q < c->world; q++) dst[q] = c->arena.vec(q, 1);
            SB_TRY(launch_peer_publish(x0.p, p.nloc, p.row0, dst, ws->ctl.p, make_px(c, -1, false), 1, st));
            x0full = T[1];
        } else if (multi) {
            SB_CUDA(cudaMemcpyAsync(T[1] + p.row0, x0.p, p.nloc * 8, cudaMemcpyDeviceToDevice, st));
            SB_NCCL(g_nccl.AllGather(T[1] + (uint64_t)c->rank * p.per, T[1], p.per, ncclDouble, c->comm, st));
            x0full = T[1];
        }
        TileKernelArgs a{};
        fill_tile_args(m, a);
        a.xin = x0full;
        a.xin_own = x0.p;
        a.out = ws->tmp.p;
        SB_TRY(launch_tile_kernel(cfg, EPI_SPMV, a, st));
        ax0 = ws->tmp.p;
        extra_matvec = 1;
    }

    SB_CUDA(cudaEventRecord(ws->ev0, st));
    {
        InitArgs ia{};
        ia.x0 = x0.p;
        ia.ax0 = ax0;
        ia.b = ws->b.p;
        ia.dinv = dinv;
        ia.c_out = ws->c.p;
        ia.t_out = T[0] + p.row0;
        ia.x_out = ws->x.p;
        ia.n = (uint32_t)p.nloc;
        ia.compat = compat;
        ia.ctl = ws->ctl.p;
        ia.partials = ws->partials.p;
        ia.identity_res = identity;
        ia.defer_tail = multi && !p2p;
        const bool resid_due = !identity;
        ia.last_in_iter = !resid_due;
        ia.row_base = (uint32_t)p.row0;
        if (p2p) ia.px = make_px(c, 0, resid_due);
        SB_TRY(launch_init_state(ia, st));
        launches++;
        if (!p2p) {
            SB_TRY(exchange_term(c, ws->ctl.p, T[0], p.per, st));
            if (multi) SB_TRY(launch_dist_tail(ws->ctl.p, 1, 0, !resid_due, identity, 0, nullptr, st));
        }
        if (resid_due) SB_TRY(enqueue_resid(0, T[1], 1, 0, true));
    }
    uint64_t it = 1;
    const uint64_t push_end = std::min(max_it, max_terms);
    const uint64_t kBatch = 8;
    bool alive = true;
    while (alive && it < push_end) {
        const uint64_t end = std::min(push_end, it + kBatch);
        for (; it < end; it++) {
            const bool resid_due = !identity && (it % 5 == 0);
            TileKernelArgs a = base;
            a.xin = T[(it - 1) & 1];
            a.xin_own = a.xin + p.row0;
            a.out = T[it & 1] + p.row0;
            a.sol = ws->x.p;
            a.dinv = dinv;
            a.it = (uint32_t)it;
            a.last_in_iter = !resid_due;
            if (p2p) a.px = make_px(c, (int)(it & 1), resid_due);
            if (profiling) {
                ProfEv q{nullptr, nullptr, it};
                SB_CUDA(cudaEventCreate(&q.e0));
                SB_CUDA(cudaEventCreate(&q.e1));
                SB_CUDA(cudaEventRecord(q.e0, st));
                prof.push_back(q);
            }
            SB_TRY(launch_tile_kernel(cfg, EPI_PUSH, a, st));
            if (profiling) SB_CUDA(cudaEventRecord(prof.back().e1, st));
            launches++;
            if (!p2p) {
                SB_TRY(exchange_term(c, ws->ctl.p, T[it & 1], p.per, st));
                if (multi) SB_TRY(launch_dist_tail(ws->ctl.p, 1, (uint32_t)it, !resid_due, identity, 0, nullptr, st));
            }
            // the previous term buffer is dead once this push has run: reuse it as the x allgather target
            if (resid_due) SB_TRY(enqueue_resid(it, T[(it - 1) & 1], 1, 0, true));
        }
        SB_TRY(read_ctl());
        alive = ws->h_ctl->alive != 0;
    }
    if (it <= 1) {
        SB_TRY(read_ctl());
        alive = ws->h_ctl->alive != 0;
    }
    uint64_t iterations = ws->h_ctl->iterations;
    const uint64_t terms = ws->h_ctl->terms;
    uint64_t resid_in_loop = 0;
    double *scratch = T[terms & 1];  // current term lives in T[(terms-1)&1]
    if (alive && iterations >= max_terms && iterations < max_it) {  // spin phase, see solver.cu
        if (identity) {
            iterations = (ws->h_ctl->res_norm <= opt->tolerance) ? iterations : max_it;
        } else {
            const uint64_t first = (iterations + 4) / 5 * 5;
            if (first < max_it) {
                SB_TRY(enqueue_resid(first, scratch, 0, 1, false));
                SB_TRY(read_ctl());
                resid_in_loop++;
                const double r = ws->h_ctl->res_norm;
                if (!std::isfinite(r)) { ws->h_ctl->nonfinite = 1; iterations = first + 1; }
                else if (r <= opt->tolerance) iterations = first + 1;
                else { for (uint64_t k = first + 5; k < max_it; k += 5) resid_in_loop++; iterations = max_it; }
            } else {
                iterations = max_it;
            }
        }
    }
    const bool nonfinite_spin = ws->h_ctl->nonfinite != 0;
    SB_TRY(enqueue_resid(iterations, scratch, 0, 1, false));  // final residual (:516)
    SB_CUDA(cudaEventRecord(ws->ev1, st));
    SB_TRY(read_ctl());
    float dev_ms = 0.f;
    SB_CUDA(cudaEventElapsedTime(&dev_ms, ws->ev0, ws->ev1));
    SB_TRY(copy_d2h(x_local, ws->x.p, p.nloc * 8, st));
    SB_CUDA(cudaStreamSynchronize(st));
    if (ws->h_ctl->peer_timeout)
        return fail(SB200_ERR_ALGORITHM, "a peer rank never signalled its exchange (rank died or diverged)");

    const LoopCtl &cc = *ws->h_ctl;
    SolveStats stt{};
    stt.iterations = iterations;
    stt.terms = cc.terms;
    stt.series_converged = cc.sconv != 0;
    stt.residual_norm = cc.res_norm;
    stt.last_term_norm = std::sqrt(cc.term_norm2);
    stt.rhs_norm = std::sqrt(cc.rhs_norm2);
    stt.nonfinite = cc.nonfinite != 0 || nonfinite_spin;
    stt.device_ms = dev_ms;
    stt.launches = launches;
    uint64_t loop_resids = 0;
    if (!identity) {
        const uint64_t counted = std::min<uint64_t>(cc.iterations, iterations);
        loop_resids = counted ? (counted - 1) / 5 + 1 : 0;
    }
    stt.matvec = (cc.terms > 0 ? cc.terms - 1 : 0) + loop_resids + resid_in_loop + 1 + extra_matvec;
    stt.converged = (stt.residual_norm <= opt->tolerance) || (stt.series_converged && cc.terms < max_terms);

    out->residual_norm = stt.residual_norm;
    out->iterations = stt.iterations;
    out->converged = stt.converged;
    out->terms_computed = stt.terms;
    out->series_converged = stt.series_converged;
    out->last_term_norm = stt.last_term_norm;
    out->device_time_ms = stt.device_ms;
    out->kernel_launches = stt.launches;
    out->memory_bytes = ws->bytes();
    out->matvec_count = stt.matvec;
    out->has_stats = opt->collect_stats != 0;
    out->h2d_bytes = p.nloc * 8 * (x0.p ? 2 : 1);
    out->d2h_bytes = p.nloc * 8;
    for (auto &q : prof) {
        if (q.it >= cc.terms) continue;  // launches past the end of the loop were no-ops
        float ms = 0.f;
        SB_CUDA(cudaEventElapsedTime(&ms, q.e0, q.e1));
        out->push_kernel_ms += ms;
        out->push_kernel_count++;
    }
    if (stt.nonfinite)
        return fail(SB200_ERR_NUMERICAL_INSTABILITY, "Non-finite residual norm at iteration %llu", (unsigned long long)iterations);
    if (!stt.converged && iterations >= max_it)
        return fail(SB200_ERR_CONVERGENCE_FAILURE, "neumann failed to converge: %llu iterations, residual %.6e, tolerance %.6e",
                    (unsigned long long)iterations, stt.residual_norm, opt->tolerance);
    return SB200_OK;
}

}  // extern "C"
