// kernels.cu — the sm_100a kernels of the push-iteration path.
//
// One kernel body does every pass over the matrix; what differs is the per-row epilogue (row_epilogue below):
//   EPI_SPMV  : y = A x                      Matrix::multiply_vector        (ref src/matrix/sparse.rs:187-203)
//   EPI_PUSH  : t' = t - D^-1 (A t); x += t'; ||t'||^2     apply_iteration_matrix + accumulate + l2_norm
//                                            (ref src/solver/neumann.rs:280-299, 264-266, 271)
//   EPI_RESID : ||A x - rhs||^2               update_residual               (ref src/solver/neumann.rs:302-318)
//   EPI_CG    : ap = A p, p.ap                the SpMV + dot of OptimizedConjugateGradientSolver::solve
//                                            (ref src/optimized_solver.rs:224-232)
//
// B200 mapping (HBM / L2-bound sparse gather-reduce; tensor cores are irrelevant here). Three bodies, all of which add
// the products of a row LEFT TO RIGHT — the reference's order — with FMA contraction off, so results are bit-identical
// to CSRStorage::multiply_vector and to each other:
//   * sell_kernel : SELL-32 copy of the matrix, lane r owns row r of a 32-row block, no shared memory at all
//                   (the whole 256 KB array serves as L1 for the in-flight gathers); default for vectors <= 48 MB and
//                   band-local matrices;
//   * warp_kernel : CSR slices, coalesced 128/256-bit stream loads by the warp, products transposed through a 1 KB
//                   warp-private shared-memory chunk; used for ragged / power-law rows and as the body of the
//   * column-slab passes (launch_tile_kernel): when the gathered vector does not fit the L2 partition of a die, one
//                   warp_kernel pass per <= 28 MB slab of the vector, the running row sums handed from pass to pass
//                   (2.4 -> 1.07 L2 sector operations per gather);
//   * tile_kernel : TMA-staged tile pipeline (cp.async.bulk + mbarrier, LDGSTS gathers), kept selectable
//                   ($SUBLINEAR_B200_TILE_CFG) as the record of what was measured;
//   * long_rows_* : grid-wide pre-pass for hub rows (> 1024 entries) of power-law graphs.
// The stream (values, column indices) is read with L1::no_allocate + L2 evict_first, the gathered vector with L2
// evict_last (a persisting-L2 set-aside is reserved once per device). The epilogue (diagonal scale, term / solution
// update, squared norm) is fused; norms are reduced deterministically (fixed-shape tree per CTA or warp, fixed-order
// sum of the partials by the last CTA) and the last CTA also advances the device-resident loop state (LoopCtl), so the
// host never has to synchronise per term.
#include "common.hpp"

namespace sb200 {

const TileCfg kTileCfgs[kNumTileCfgs] = {{256, 128, 1536}, {512, 256, 2816}, {256, 128, 1376}, {384, 192, 2112}, {512, 256, 3072}, {256, 128, 1408}};

int default_tile_cfg() {
    static int cfg = [] {
        const char *e = getenv("SUBLINEAR_B200_TILE_CFG");
        int v = e ? atoi(e) : -1;  // -1 = warp-stream kernel (default); 0.. = TMA-staged tile pipeline variants
        return (v >= -1 && v < kNumTileCfgs) ? v : -1;
    }();
    return cfg;
}

// ---------------------------------------------------------------------------------------------------------
// PTX helpers: mbarrier + 1-D TMA bulk copy + cache policies
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!done);
}
// global -> shared bulk copy executed by the TMA unit; completion is signalled on `bar` in bytes.
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes, uint64_t *bar,
                                         uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::
            "r"(smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
        : "memory");
}
// order generic-proxy accesses to shared memory before later async-proxy (TMA) writes to the same bytes
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ uint64_t policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
// the random gather: read-only path, keep the line in L2 (it is the only reused data of the iteration)
__device__ __forceinline__ double ld_gather(const double *p, uint64_t policy) {
    double v;
    asm volatile("ld.global.nc.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(policy));
    return v;
}
// streaming loads for the long-row path
__device__ __forceinline__ double ld_stream_f64(const double *p) {
    double v;
    asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ uint32_t ld_stream_u32(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}

// streaming (use-once) accesses with an explicit L2 policy
__device__ __forceinline__ uint32_t ld_stream_u32_hint(const uint32_t *p, uint64_t policy) {
    uint32_t v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(policy));
    return v;
}
__device__ __forceinline__ double ld_stream_f64_hint(const double *p, uint64_t policy) {
    double v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(policy));
    return v;
}

// the same for memory that this kernel also writes (the carried row sums): no .nc
__device__ __forceinline__ double ld_once_f64_hint(const double *p, uint64_t policy) {
    double v;
    asm volatile("ld.global.L1::no_allocate.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(policy) : "memory");
    return v;
}
__device__ __forceinline__ void st_stream_f64_hint(double *p, double v, uint64_t policy) {
    asm volatile("st.global.L2::cache_hint.f64 [%0], %1, %2;" ::"l"(p), "d"(v), "l"(policy) : "memory");
}

// ---------------------------------------------------------------------------------------------------------
// deterministic reductions + device-side loop control
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// sum over the CTA; result valid in thread 0. s_red: NT/32 doubles.
template <int NT>
__device__ __forceinline__ double block_sum(double v, double *s_red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    v = warp_sum(v);
    __syncthreads();  // s_red may still be read from a previous call
    if (lane == 0) s_red[warp] = v;
    __syncthreads();
    double r = 0.0;
    if (warp == 0) {
        r = (lane < NT / 32) ? s_red[lane] : 0.0;
        r = warp_sum(r);
    }
    return r;
}

// End of one iteration of the `while` loop in NeumannSolver::solve (ref src/solver/neumann.rs:498-512 and the
// loop condition :481 evaluated for the next iteration).
__device__ __forceinline__ void end_of_iteration(LoopCtl *c, uint32_t it) {
    c->iterations = it + 1;
    if (!isfinite(c->res_norm)) {  // :501-507 NumericalInstability
        c->nonfinite = 1;
        c->alive = 0;
        return;
    }
    if (c->sconv) {  // :510-512
        c->alive = 0;
        return;
    }
    // :481 `!is_converged && iterations < max_iterations`; series_converged is false here, so
    // is_converged (:422-430) reduces to residual_norm <= tolerance.
    if (c->res_norm <= c->tolerance || it + 1 >= c->max_iterations) c->alive = 0;
}

enum TailKind { TAIL_NONE = 0, TAIL_TERM = 1, TAIL_RESID = 2, TAIL_CG_INIT = 3, TAIL_CG_PAP = 4, TAIL_CG_RS = 5 };

__device__ __forceinline__ void tail_logic(LoopCtl *c, int kind, double sum, double aux, uint32_t it, int last_in_iter,
                                           int identity_res, int defer, double *norm_log) {
    if (defer) {  // row-partitioned: publish this rank's sums; dist_tail_kernel finishes after the allreduce
        c->red[0] = sum;
        c->red[1] = aux;
        return;
    }
    if (kind == TAIL_TERM) {
        c->term_norm2 = sum;
        if (norm_log) norm_log[it] = sum;
        if (identity_res) c->aux_norm2 = aux;
        c->terms = it + 1;  // ref :268
        if (it == 0) c->rhs_norm2 = sum;
        if (sqrt(sum) < c->series_tolerance) c->sconv = 1;  // ref :271-274
        if (identity_res) {
            c->res_norm2 = aux;
            c->res_norm = sqrt(aux);
        }
    } else if (kind == TAIL_RESID) {
        c->res_norm2 = sum;
        c->res_norm = sqrt(sum);  // ref :316
    }
    else if (kind == TAIL_CG_INIT) {  // rsold = r.r with r = b (optimized_solver.rs:211-215), loop test of iteration 0
        c->cg_rsold = sum;
        c->iterations = 0;
        if (c->max_iterations == 0) c->alive = 0;
        else if (sum <= c->cg_tol_sq) { c->cg_converged = 1; c->alive = 0; }
    } else if (kind == TAIL_CG_PAP) {  // :228-238
        c->cg_pap = sum;
        c->cg_matvecs += 1;
        if (fabs(sum) < 1e-16) { c->cg_breakdown = 1; c->alive = 0; }  // `break` before x is touched
        else c->cg_alpha = c->cg_rsold / sum;
    } else if (kind == TAIL_CG_RS) {  // :250-264, then the `while` / `if rsold <= tolerance_sq` of the next pass (:217-221)
        c->cg_beta = sum / c->cg_rsold;
        c->cg_rsold = sum;
        c->iterations += 1;
        if (c->iterations >= c->max_iterations) c->alive = 0;
        else if (sum <= c->cg_tol_sq) { c->cg_converged = 1; c->alive = 0; }
    }
    if (last_in_iter) end_of_iteration(c, it);
}

// system-scope flag/slot accessors for the peer exchange
__device__ __forceinline__ void st_release_sys_u64(unsigned long long *p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys_u64(const unsigned long long *p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed_sys_f64(double *p, double v) {
    asm volatile("st.relaxed.sys.global.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}
__device__ __forceinline__ double ld_relaxed_sys_f64(const double *p) {
    double v;
    asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}

// P2P signal: publish this rank's (sum, aux) of the current exchange into every rank's slots, then raise its flag
// everywhere. Called by one thread after the whole grid's stores are ordered before it (ticket + system fences).
__device__ __forceinline__ void peer_signal(const LoopCtl *ctl, const PeerExchange &px, double sum, double aux) {
    const unsigned long long e = px.epoch_base + ctl->xchg + 1ull;
    const unsigned par = (unsigned)(e & 1ull);
    for (int p = 0; p < px.world; p++) {
        double *s = px.slots[p] + ((size_t)par * px.world + px.rank) * 2;
        st_relaxed_sys_f64(s, sum);
        st_relaxed_sys_f64(s + 1, aux);
    }
    __threadfence_system();
    for (int p = 0; p < px.world; p++) st_release_sys_u64(px.flags[p] + px.rank, e);
}

// CTA partial -> global partial array -> the last CTA to arrive sums all partials in index order.
template <int NT>
__device__ __forceinline__ void grid_reduce_and_tail(double sq, double aux, LoopCtl *ctl, double *partials, int kind,
                                                     uint32_t it, int last_in_iter, int identity_res, int defer,
                                                     double *norm_log, double *s_red, int *s_flag,
                                                     const PeerExchange *px = nullptr) {
    const bool p2p = px != nullptr && px->world > 1;
    if (p2p) __threadfence_system();  // this thread's stores into peer memory, before the CTA reports in
    double bs = block_sum<NT>(sq, s_red);
    double ba = identity_res ? block_sum<NT>(aux, s_red) : 0.0;
    if (threadIdx.x == 0) {
        partials[blockIdx.x] = bs;
        if (identity_res) partials[gridDim.x + blockIdx.x] = ba;
        if (p2p) __threadfence_system(); else __threadfence();
        unsigned t = atomicAdd(&ctl->ticket, 1u);
        *s_flag = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (*s_flag) {
        __threadfence();
        double s = 0.0, a = 0.0;
        for (unsigned i = threadIdx.x; i < gridDim.x; i += NT) s += __ldcg(partials + i);
        if (identity_res)
            for (unsigned i = threadIdx.x; i < gridDim.x; i += NT) a += __ldcg(partials + gridDim.x + i);
        s = block_sum<NT>(s, s_red);
        if (identity_res) a = block_sum<NT>(a, s_red);
        if (threadIdx.x == 0) {
            ctl->ticket = 0;
            if (p2p) {
                __threadfence_system();
                peer_signal(ctl, *px, s, a);  // the wait kernel that follows runs tail_logic on the global sums
            } else {
                tail_logic(ctl, kind, s, a, it, last_in_iter, identity_res, defer, norm_log);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// the tile kernel
// ---------------------------------------------------------------------------------------------------------
// Shared-memory plan per CTA (CAPE = CAP + 8 element slots, D = gather depth in tiles):
//   cols : D+2 stages x CAPE u32   TMA, D+1 tiles ahead (the gathers of tile k+D need its column indices)
//   xg   : D+1 stages x CAPE x 16B cp.async.cg gather targets, D tiles ahead: the aligned pair holding x[col].
//          .cg bypasses L1, so in-flight gathers are not capped by the L1 lines the carve-out leaves over
//          (bench/gather_probe.cu: 266 G gathers/s at any carve-out vs 122 G/s for 8-byte .ca gathers)
//   vals : 2 stages x CAPE f64     TMA, one tile ahead (only the product phase of the current tile reads them)
//   win  : 2 stages x (R+2) f64    TMA, one tile ahead: the tile's OWN slice x[row0..row1) of the gather source.
//          Columns that fall into it (always the diagonal; nearly everything for banded matrices) are served
//          from shared memory and never gathered; it also supplies t_i for the push epilogue.
// Software pipeline, iteration k of a CTA:
//   TMA cols(k+D+1), TMA vals+win(k+1) -> cp.async gathers(k+D) -> wait gathers(k), vals(k)
//   -> product phase(k) -> ordered row sums + epilogue(k)
// The number of gathers in flight per SM is (tiles in flight) x (nnz per tile); it is bounded by shared memory,
// not registers, and the streamed arrays never touch the LSU/L1 path that the gathers are bound by.
template <int R, int CAP, int D>
struct TileSmem {
    static constexpr int kElems = CAP + 8;  // shift (<=3) + round-up (<=3) slack
    static constexpr int kWin = R + 2;
    static constexpr size_t kXgOff = 0;                                             // 16-byte aligned slots first
    static constexpr size_t kValsOff = kXgOff + (size_t)(D + 1) * kElems * 16;
    static constexpr size_t kWinOff = kValsOff + (size_t)2 * kElems * 8;
    static constexpr size_t kColsOff = kWinOff + (size_t)2 * kWin * 8;
    static constexpr size_t kBytes = kColsOff + (size_t)(D + 2) * kElems * 4;
};

struct TileRange {
    uint32_t row0, nnz0, row1, nnz1;
};

// 16-byte global -> shared asynchronous copy that bypasses L1 (LDGSTS.BYPASS): the gather result never occupies a
// register or an L1 line. src must be 16-byte aligned: the caller fetches the aligned pair around x[col].
__device__ __forceinline__ void cp_async_gather16(void *dst_smem, const void *src_gmem, uint64_t policy) {
    asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem),
                 "l"(policy)
                 : "memory");
}
__device__ __forceinline__ void cp_async_copy8(void *dst_smem, const void *src_gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(dst_smem)), "l"(src_gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// NT threads per CTA, at most R <= NT rows and CAP non-zeros per tile: all NT threads issue gathers and form
// products; the first R threads own one row each for the ordered sum and the epilogue.
template <int EPI, int NT, int R, int CAP, int D>
__global__ void __launch_bounds__(NT) tile_kernel(const TileKernelArgs a) {
    static_assert(R <= NT && NT % 32 == 0 && R % 2 == 0 && D >= 1 && D <= 3, "tile geometry");
    using SM = TileSmem<R, CAP, D>;
    constexpr int NC = D + 2, NX = D + 1;       // cols / xg stages
    constexpr uint32_t kFill = 2 * D + 3;        // tile ranges are fetched this many tiles ahead
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t s_bar_c[NC];
    __shared__ __align__(8) uint64_t s_bar_v[2];
    __shared__ double s_red[NT / 32];
    __shared__ int s_flag;
    // tile ranges of this CTA's tile sequence, fetched ahead with cp.async: a plain load would sit on every warp's
    // critical path once per tile (the compiler keeps the range in uniform registers)
    __shared__ __align__(16) TileRange s_ring[16];

    if (EPI != EPI_SPMV) {
        if (!a.force && a.ctl->alive == 0) return;  // loop already finished: no-op launch
    }

    const int tid = threadIdx.x;
    auto s_xg = [&](int s) { return reinterpret_cast<double2 *>(smem_raw + SM::kXgOff) + (size_t)s * SM::kElems; };
    auto s_val = [&](int s) { return reinterpret_cast<double *>(smem_raw + SM::kValsOff) + (size_t)s * SM::kElems; };
    auto s_win = [&](int s) { return reinterpret_cast<double *>(smem_raw + SM::kWinOff) + (size_t)s * SM::kWin; };
    auto s_col = [&](int s) { return reinterpret_cast<uint32_t *>(smem_raw + SM::kColsOff) + (size_t)s * SM::kElems; };
    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < NC; i++) mbar_init(&s_bar_c[i], 1);
#pragma unroll
        for (int i = 0; i < 2; i++) mbar_init(&s_bar_v[i], 1);
        mbar_fence_init();
    }

    const uint64_t pol_stream = policy_evict_first();
    const uint64_t pol_gather = policy_evict_last();
    const TileDesc *__restrict__ tiles = a.tiles;
    const uint32_t ntiles = a.ntiles;
    const uint32_t stride = gridDim.x;

    // s_ring[seq & 15] <- range of the seq-th tile of this CTA; past the end: an empty range (nothing is issued for
    // it). Threads 0/1 copy the two adjacent descriptors {row0,nnz0},{row1,nnz1}; joins the caller's cp.async group.
    auto ring_fill = [&](uint32_t seq) {
        if (tid < 2) {
            const uint64_t tile_idx = (uint64_t)blockIdx.x + (uint64_t)seq * stride;
            TileDesc *dst = reinterpret_cast<TileDesc *>(&s_ring[seq & 15u]) + tid;
            if (tile_idx < ntiles) cp_async_copy8(dst, tiles + tile_idx + tid);
            else *dst = TileDesc{0u, 0u};
        }
    };
    // staged = streamed through shared memory; empty tiles and long-row tiles (cnt > CAP) are not
    auto staged = [&](const TileRange &r) { const uint32_t c = r.nnz1 - r.nnz0; return c != 0u && c <= (uint32_t)CAP; };
    // the tile's own slice of the gather source, in global column coordinates, widened to 16-byte boundaries:
    // [wa, wa + wlen). wlen = 0 when the widened slice would leave the vector (odd tail) or the tile is not staged.
    struct Window {
        uint32_t wa, wlen;
    };
    auto window_of = [&](const TileRange &r) -> Window {
        Window w{0u, 0u};
        if (!staged(r)) return w;
        const uint64_t g0 = (uint64_t)a.row_base + r.row0, g1 = (uint64_t)a.row_base + r.row1;
        const uint64_t wa = g0 & ~1ull, wb = (g1 + 1ull) & ~1ull;
        if (wb <= a.xin_len && g1 > g0) {
            w.wa = (uint32_t)wa;
            w.wlen = (uint32_t)(wb - wa);
        }
        return w;
    };
    auto issue_cols = [&](const TileRange &r, int slot) {  // producer thread only
        if (!staged(r)) return;
        const uint32_t a0 = r.nnz0 & ~3u;
        const uint32_t nel = ((r.nnz0 - a0) + (r.nnz1 - r.nnz0) + 3u) & ~3u;
        mbar_arrive_expect_tx(&s_bar_c[slot], nel * 4u);
        bulk_g2s(s_col(slot), a.cols + a0, nel * 4u, &s_bar_c[slot], pol_stream);
    };
    auto issue_vals = [&](const TileRange &r, int slot) {  // producer thread only: values + own window
        if (!staged(r)) return;
        const uint32_t a0 = r.nnz0 & ~3u;
        const uint32_t nel = ((r.nnz0 - a0) + (r.nnz1 - r.nnz0) + 3u) & ~3u;
        const Window w = window_of(r);
        mbar_arrive_expect_tx(&s_bar_v[slot], nel * 8u + w.wlen * 8u);
        bulk_g2s(s_val(slot), a.vals + a0, nel * 8u, &s_bar_v[slot], pol_stream);
        if (w.wlen) bulk_g2s(s_win(slot), a.xin + w.wa, w.wlen * 8u, &s_bar_v[slot], pol_gather);
    };
    uint32_t phase_c = 0, phase_v = 0;  // bit s = parity to wait for on slot s
    // every thread: wait for the tile's column indices, then launch its share of the x[col] gathers
    auto issue_gathers = [&](const TileRange &r, int cslot, int xslot) {
        if (staged(r)) {
            mbar_wait(&s_bar_c[cslot], (phase_c >> cslot) & 1u);
            phase_c ^= (1u << cslot);
            const uint32_t shift = r.nnz0 & 3u, cnt = r.nnz1 - r.nnz0;
            const Window w = window_of(r);
            const uint32_t *__restrict__ sc = s_col(cslot) + shift;
            double2 *xg = s_xg(xslot) + shift;
#pragma unroll 4
            for (uint32_t j = tid; j < cnt; j += NT) {
                const uint32_t c = sc[j];
                if (c - w.wa >= w.wlen) cp_async_gather16(xg + j, a.xin + (c & ~1u), pol_gather);
            }
        }
        cp_async_commit();  // always: keeps the group count in step with the tile count
    };
    // per-row operands of the epilogue, fetched one tile ahead (t_i itself comes from the window)
    struct RowOps {
        uint32_t rs, re;
        double dv, xs, rh;
    };
    auto load_rows = [&](const TileRange &r) -> RowOps {
        RowOps o{0u, 0u, 0.0, 0.0, 0.0};
        if ((uint32_t)tid < r.row1 - r.row0) {
            const uint32_t row = r.row0 + tid;
            o.rs = a.row_ptr[row];
            o.re = a.row_ptr[row + 1];
            if (EPI == EPI_PUSH) {
                o.dv = a.dinv[row];
                o.xs = a.sol[row];
            } else if (EPI == EPI_RESID) {
                o.rh = a.rhs[row];
            } else if (a.accumulate) {
                o.xs = a.out[row];
            }
        }
        return o;
    };

    // ---- prologue: ranges of the first kFill tiles, cols of tiles 0..D, vals of tile 0, gathers of tiles 0..D-1 ----
    for (uint32_t s = 0; s < kFill; s++) ring_fill(s);
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();  // also publishes the mbarrier inits
    TileRange rr[D + 2];  // ranges of tiles k .. k+D+1
#pragma unroll
    for (int i = 0; i < D + 2; i++) rr[i] = s_ring[i];
    if (tid == 0) {
#pragma unroll
        for (int i = 0; i <= D; i++) issue_cols(rr[i], i);
        issue_vals(rr[0], 0);
    }
#pragma unroll
    for (int i = 0; i < D; i++) issue_gathers(rr[i], i, i);
    RowOps cur = load_rows(rr[0]);

    double sq = 0.0, aux = 0.0;
    uint32_t seq = 0;
    int c0 = 0, x0 = 0, v0 = 0;  // cols / xg / vals slot of tile k
    // optional phase timing (debug aid, off unless phase_log is set): cycles seen by one thread per CTA
    constexpr int kPhases = 12;
    unsigned long long pacc[kPhases];
#pragma unroll
    for (int i = 0; i < kPhases; i++) pacc[i] = 0ull;
    const bool ptime = a.phase_log != nullptr && tid == (int)(a.phase_log[63] % NT);
    long long pt0 = ptime ? clock64() : 0;
#define SB_PHASE(p)                              \
    if (ptime) {                                 \
        const long long t1_ = clock64();         \
        pacc[p] += (unsigned long long)(t1_ - pt0); \
        pt0 = t1_;                               \
    }
    for (uint32_t tile = blockIdx.x; tile < ntiles; tile += stride, seq++) {
        ring_fill(seq + kFill);                                           // complete + visible well before it is read
        const TileRange rnext = s_ring[(seq + D + 2u) & 15u];              // range of tile k+D+2, used next iteration
        if (tid == 0) {
            issue_cols(rr[D + 1], (c0 + D + 1) % NC);
            issue_vals(rr[1], v0 ^ 1);
        }
        SB_PHASE(0)
        issue_gathers(rr[D], (c0 + D) % NC, (x0 + D) % NX);
        SB_PHASE(1)
        const RowOps nxt = load_rows(rr[1]);
        SB_PHASE(2)

        const TileRange r0 = rr[0];
        const uint32_t nrows = r0.row1 - r0.row0, cnt = r0.nnz1 - r0.nnz0;
        const bool is_long = cnt > (uint32_t)CAP;  // exactly one row, streamed by the whole CTA
        const bool active = (uint32_t)tid < nrows;
        const uint32_t row = r0.row0 + tid;
        double sum = 0.0, own = 0.0;
        if (!is_long) {
            cp_async_wait<D>();  // this thread's gathers for tile k have landed (those of k+1..k+D may be in flight)
            SB_PHASE(3)
            if (cnt > 0) {
                mbar_wait(&s_bar_v[v0], (phase_v >> v0) & 1u);
                phase_v ^= (1u << v0);
            }
            SB_PHASE(4)
            __syncthreads();  // ... and so have everybody else's
            SB_PHASE(5)
            const Window w = window_of(r0);
            const double *__restrict__ sw = s_win(v0);
            {
                // product phase: element j -> thread j (stride-1 shared-memory traffic, no bank conflicts); the
                // product overwrites the staged value. x[col] comes from the own window or from the gathered pair.
                const uint32_t shift = r0.nnz0 & 3u;
                double *__restrict__ sv = s_val(v0) + shift;
                const double2 *__restrict__ sx = s_xg(x0) + shift;
                const uint32_t *__restrict__ sc = s_col(c0) + shift;
#pragma unroll 4
                for (uint32_t j = tid; j < cnt; j += NT) {
                    const uint32_t c = sc[j];
                    double xv;
                    if (c - w.wa < w.wlen) {
                        xv = sw[c - w.wa];
                    } else {
                        const double2 pr = sx[j];
                        xv = (c & 1u) ? pr.y : pr.x;
                    }
                    sv[j] = sv[j] * xv;
                }
            }
            SB_PHASE(6)
            __syncthreads();
            SB_PHASE(7)
            if (active) {
                // left-to-right accumulation, the order of CSRStorage::multiply_vector_add (sparse.rs:193-203)
                const uint32_t a0 = r0.nnz0 & ~3u;
                const double *__restrict__ sp = s_val(v0) - a0;
                double acc = (EPI == EPI_SPMV && a.accumulate) ? cur.xs : 0.0;
                for (uint32_t k = cur.rs; k < cur.re; k++) acc += sp[k];
                sum = acc;
                if (EPI == EPI_PUSH) {
                    const uint32_t g = a.row_base + row;
                    own = (w.wlen != 0u) ? sw[g - w.wa] : a.xin[g];
                }
            }
        } else {
            double acc = 0.0;
            for (uint32_t k = r0.nnz0 + tid; k < r0.nnz1; k += NT)
                acc += ld_stream_f64(a.vals + k) * ld_gather(a.xin + ld_stream_u32(a.cols + k), pol_gather);
            acc = block_sum<NT>(acc, s_red);
            if (tid == 0) {
                sum = (EPI == EPI_SPMV && a.accumulate) ? cur.xs + acc : acc;
                if (EPI == EPI_PUSH) own = a.xin[a.row_base + row];
            }
        }

        if (active) {
            if (EPI == EPI_SPMV) {
                a.out[row] = sum;
            } else if (EPI == EPI_PUSH) {
                const double tmp = sum * cur.dv;   // temp *= d_inv        (neumann.rs:289-291)
                const double tn = own - tmp;       // term -= temp         (neumann.rs:294-296)
                a.out[row] = tn;
                a.sol[row] = cur.xs + tn;          // solution += term     (neumann.rs:264-266)
                sq += tn * tn;                     // l2_norm accumulation (solver/mod.rs:369-371)
                if (a.identity_res) {
                    const double r = tn / cur.dv;  // (D o t')_i = (b - A x)_i, SURVEY F12
                    aux += r * r;
                }
            } else {
                const double r = sum - cur.rh;     // r = A x - rhs        (neumann.rs:308-310)
                sq += r * r;
            }
        }
        SB_PHASE(8)
        // every thread is done with this tile's stages before the TMA / the gathers refill them next iteration
        fence_proxy_async_smem();
        __syncthreads();
        SB_PHASE(9)
#pragma unroll
        for (int i = 0; i < D + 1; i++) rr[i] = rr[i + 1];
        rr[D + 1] = rnext;
        cur = nxt;
        c0 = (c0 + 1) % NC;
        x0 = (x0 + 1) % NX;
        v0 ^= 1;
    }
    cp_async_wait<0>();
    if (ptime) {
#pragma unroll
        for (int i = 0; i < kPhases; i++) atomicAdd(a.phase_log + i, pacc[i]);
        atomicAdd(a.phase_log + 32, 1ull);
    }
#undef SB_PHASE

    if (EPI != EPI_SPMV) {
        grid_reduce_and_tail<NT>(sq, aux, a.ctl, a.partials, EPI == EPI_PUSH ? TAIL_TERM : TAIL_RESID, a.it,
                                 a.last_in_iter, a.identity_res, a.defer_tail, a.norm_log, s_red, &s_flag);
    }
}

template <int EPI, int NT, int R, int CAP, int D>
static int32_t launch_one(const TileKernelArgs &a, cudaStream_t stream, int *max_grid_out) {
    using SM = TileSmem<R, CAP, D>;
    static int max_grid[16] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 16) return fail(SB200_ERR_ALGORITHM, "device index %d out of range", dev);
    if (max_grid[dev] == 0) {
        SB_CUDA(cudaFuncSetAttribute(tile_kernel<EPI, NT, R, CAP, D>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)SM::kBytes));
        int per_sm = 0, sms = 0;
        SB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, tile_kernel<EPI, NT, R, CAP, D>, NT, SM::kBytes));
        SB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        if (per_sm < 1) return fail(SB200_ERR_ALGORITHM, "tile kernel does not fit on an SM");
        max_grid[dev] = per_sm * sms;  // persistent grid: a whole number of CTAs per SM (148 SMs on B200)
    }
    if (max_grid_out) {
        *max_grid_out = max_grid[dev];
        return SB200_OK;
    }
    if (a.ntiles == 0 && EPI == EPI_SPMV) return SB200_OK;
    unsigned grid = a.ntiles < (uint32_t)max_grid[dev] ? a.ntiles : (uint32_t)max_grid[dev];
    if (grid == 0) grid = 1;  // reductions still need their tail
    tile_kernel<EPI, NT, R, CAP, D><<<grid, NT, SM::kBytes, stream>>>(a);
    SB_CUDA(cudaGetLastError());
    return SB200_OK;
}

template <int NT, int R, int CAP, int D>
static int32_t launch_cfg(Epilogue epi, const TileKernelArgs &a, cudaStream_t stream, int *mg) {
    switch (epi) {
        case EPI_SPMV: return launch_one<EPI_SPMV, NT, R, CAP, D>(a, stream, mg);
        case EPI_PUSH: return launch_one<EPI_PUSH, NT, R, CAP, D>(a, stream, mg);
        default: return launch_one<EPI_RESID, NT, R, CAP, D>(a, stream, mg);
    }
}

// id, threads, rows, cap, depth   (keep in sync with kTileCfgs)
#define SB_TILE_CFGS(X) \
    X(0, 256, 128, 1536, 1) X(1, 512, 256, 2816, 2) X(2, 256, 128, 1376, 2) X(3, 384, 192, 2112, 3) \
    X(4, 512, 256, 3072, 1) X(5, 256, 128, 1408, 3)

static int32_t launch_any(int cfg, Epilogue epi, const TileKernelArgs &a, cudaStream_t stream, int *mg) {
    switch (cfg) {
#define X(id, nt, r, cap, d) \
    case id: return launch_cfg<nt, r, cap, d>(epi, a, stream, mg);
        SB_TILE_CFGS(X)
#undef X
        default: return fail(SB200_ERR_INVALID_INPUT, "unknown tile configuration %d", cfg);
    }
}

static int32_t launch_warp_any(Epilogue epi, const TileKernelArgs &a, cudaStream_t stream, int *mg);

static int32_t launch_sell_any(Epilogue epi, const TileKernelArgs &a, cudaStream_t stream, int *mg);

// ---------------------------------------------------------------------------------------------------------
// hub rows: a row with millions of entries must not be left to one warp (measured: 470 ms per SpMV for a PageRank
// system with a 2 M-entry row). Ingest lists the rows above kLongRow entries and cuts them into chunks of kLongChunk;
// before the row-block kernel, one CTA per chunk forms a partial sum (coalesced stream loads, lane-strided order +
// fixed tree) and one thread per row adds the row's partials in order. Tolerance-level parity, as for every long row.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) long_rows_chunk_kernel(const TileKernelArgs a, double *partial, int check_alive) {
    __shared__ double s_red[8];
    if (check_alive && a.ctl->alive == 0) return;  // loop already finished: no-op launch
    const uint64_t pol_gather = policy_evict_last();
    for (uint32_t c = blockIdx.x; c < a.nlong_chunks; c += gridDim.x) {
        const uint2 r = a.long_chunks[c];
        double part = 0.0;
        for (uint32_t k = r.x + threadIdx.x; k < r.y; k += 256u)
            part += ld_stream_f64(a.vals + k) * ld_gather(a.xin + ld_stream_u32(a.cols + k), pol_gather);
        part = block_sum<256>(part, s_red);
        if (threadIdx.x == 0) partial[c] = part;
    }
}

__global__ void long_rows_combine_kernel(const TileKernelArgs a, const double *partial, double *sum, int check_alive) {
    if (check_alive && a.ctl->alive == 0) return;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < a.nlong; i += gridDim.x * blockDim.x) {
        double s = 0.0;
        for (uint32_t c = a.long_first[i]; c < a.long_first[i + 1u]; c++) s += partial[c];
        sum[i] = s;
    }
}

static int32_t launch_with_long_rows(Epilogue epi, const TileKernelArgs &a, cudaStream_t stream) {
    double *scratch = nullptr;  // per launch, stream ordered: concurrent solves on one matrix handle do not share it
    SB_CUDA(cudaMallocAsync((void **)&scratch, ((size_t)a.nlong_chunks + a.nlong) * sizeof(double), stream));
    const int check_alive = epi != EPI_SPMV && !a.force;
    const unsigned grid = a.nlong_chunks < 148u * 8u ? a.nlong_chunks : 148u * 8u;
    long_rows_chunk_kernel<<<grid, 256, 0, stream>>>(a, scratch, check_alive);
    long_rows_combine_kernel<<<(a.nlong + 127u) / 128u, 128, 0, stream>>>(a, scratch, scratch + a.nlong_chunks, check_alive);
    TileKernelArgs p = a;
    p.long_sum = scratch + a.nlong_chunks;
    int32_t rc = cudaGetLastError() == cudaSuccess ? SB200_OK : fail(SB200_ERR_ALGORITHM, "long-row pre-pass launch failed");
    if (rc == SB200_OK) rc = launch_warp_any(epi, p, stream, nullptr);
    cudaFreeAsync(scratch, stream);
    return rc;
}

int32_t launch_tile_kernel(int cfg, Epilogue epi, const TileKernelArgs &a, cudaStream_t stream) {
    if (cfg < 0 && a.nslabs > 1) {
        // column-slab passes: the gather source of one pass is a slab of the vector small enough to stay in the L2
        // partition of each die; the row sums continue from pass to pass in column = CSR order
        double *acc = a.acc ? a.acc : a.out;
        if (!acc) return fail(SB200_ERR_ALGORITHM, "column-slab passes need a buffer for the partial row sums");
        for (int s = 0; s < a.nslabs; s++) {
            TileKernelArgs p = a;
            p.vals = a.slab_vals[s];
            p.cols = a.slab_cols[s];
            p.row_ptr = a.slab_row_ptr[s];
            p.sell_ptr = nullptr;
            p.acc_in = s > 0 ? acc : nullptr;
            p.acc_out = s + 1 < a.nslabs ? acc : nullptr;
            SB_TRY(launch_warp_any(epi, p, stream, nullptr));
        }
        return SB200_OK;
    }
    if (cfg < 0 && a.sell_ptr != nullptr) return launch_sell_any(epi, a, stream, nullptr);
    if (cfg < 0 && a.nlong > 0) return launch_with_long_rows(epi, a, stream);
    if (cfg < 0) return launch_warp_any(epi, a, stream, nullptr);
    if (epi == EPI_CG) return fail(SB200_ERR_INVALID_INPUT, "the CG epilogue exists in the warp-stream kernel only");
    if (reinterpret_cast<uintptr_t>(a.xin) & 15u)  // 16-byte gathers and the TMA window copy
        return fail(SB200_ERR_INVALID_INPUT, "device vectors must be 16-byte aligned");
    return launch_any(cfg, epi, a, stream, nullptr);
}

// upper bound of the number of partial sums a launch of this configuration writes (callers size `partials` from it)
int tile_kernel_max_grid(int cfg, Epilogue epi) {
    int mg = 0;
    TileKernelArgs dummy{};
    if (cfg >= 0) return launch_any(cfg, epi, dummy, nullptr, &mg) == SB200_OK ? mg : 0;
    int ms = 0;
    if (launch_warp_any(epi, dummy, nullptr, &mg) != SB200_OK) return 0;
    if (launch_sell_any(epi, dummy, nullptr, &ms) != SB200_OK) return 0;  // one partial per warp
    return mg > ms ? mg : ms;
}

// ---------------------------------------------------------------------------------------------------------
// per-row epilogue shared by the warp-stream and the SELL kernel: `acc` = (A xin)_row
// ---------------------------------------------------------------------------------------------------------
template <int EPI>
__device__ __forceinline__ void row_epilogue(const TileKernelArgs &a, uint32_t row, double acc, double own, double dv,
                                             double xs, double rh, double &sq, double &aux) {
    if (EPI == EPI_SPMV) {
        a.out[row] = acc;
    } else if (EPI == EPI_PUSH) {
        const double tmp = acc * dv;   // temp *= d_inv        (neumann.rs:289-291)
        const double tn = own - tmp;   // term -= temp         (neumann.rs:294-296)
        a.out[row] = tn;
        a.sol[row] = xs + tn;          // solution += term     (neumann.rs:264-266)
        if (a.px.world > 1) {
            // fused exchange: this rank's slice of the new term (and of x when a residual check follows)
            // goes straight into every peer's buffers over NVLink, 256 contiguous bytes per warp and peer
            const size_t g = (size_t)a.row_base + row;
            for (int p = 0; p < a.px.world; p++) {
                if (p == a.px.rank || !a.px.t_out[p]) continue;
                a.px.t_out[p][g] = tn;
                if (a.px.x_out[p]) a.px.x_out[p][g] = xs + tn;
            }
            if (a.px.x_out[a.px.rank]) a.px.x_out[a.px.rank][g] = xs + tn;
        }
        sq += tn * tn;                 // l2_norm accumulation (solver/mod.rs:369-371)
        if (a.identity_res) {
            const double r = tn / dv;  // (D o t')_i = (b - A x)_i, SURVEY F12
            aux += r * r;
        }
    } else if (EPI == EPI_CG) {
        a.out[row] = acc;              // ap = A p             (optimized_solver.rs:224)
        sq += own * acc;               // p^T ap               (optimized_solver.rs:228-232)
    } else {
        const double r = acc - rh;     // r = A x - rhs        (neumann.rs:308-310)
        sq += r * r;
    }
}

// per-row operands of the epilogue (coalesced: lane r <-> row r)
template <int EPI>
__device__ __forceinline__ void row_operands(const TileKernelArgs &a, uint32_t row, double &own, double &dv, double &xs,
                                             double &rh) {
    if (EPI == EPI_PUSH) {
        own = a.xin[a.row_base + row];
        dv = a.dinv[row];
        xs = a.sol[row];
    } else if (EPI == EPI_RESID) {
        rh = a.rhs[row];
    } else if (EPI == EPI_CG) {
        own = a.xin[a.row_base + row];
    } else if (a.accumulate) {
        xs = a.out[row];
    }
}

// ---------------------------------------------------------------------------------------------------------
// the warp-stream kernel (tile configuration -1)
// ---------------------------------------------------------------------------------------------------------
// Same three epilogues as the tile kernel, no CTA-level synchronisation at all: every warp owns blocks of 32
// consecutive rows.  The block's contiguous slice of col_indices / values is read with coalesced 128-bit loads
// (4 elements per lane) straight into registers, the x[col] gathers go to registers as well, and only the products
// pass through a 1 KB warp-private shared-memory chunk so that lane r can add the products of row r LEFT TO RIGHT
// (the reference's order, bit for bit) — 4-5 LSU operations per non-zero instead of the 7 of the staged pipeline,
// which matters because the load/store pipe, not HBM, is what this access pattern saturates.
// Lane r <-> row r also makes every epilogue load/store fully coalesced.
template <int EPI, int NT, int EPL>
__global__ void __launch_bounds__(NT) warp_kernel(const TileKernelArgs a) {
    static_assert(EPL == 4 || EPL == 8, "elements per lane and chunk");
    constexpr int WARPS = NT / 32;
    constexpr uint32_t CH = 32 * EPL;    // elements per chunk: EPL per lane (4: default, 8: twice the gathers in flight)
    __shared__ __align__(16) double s_prod[WARPS][CH];
    __shared__ double s_red[WARPS];
    __shared__ int s_flag;

    if (EPI != EPI_SPMV) {
        if (!a.force && a.ctl->alive == 0) return;  // loop already finished: no-op launch
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t pol_stream = policy_evict_first();
    const uint64_t pol_gather = policy_evict_last();
    const uint32_t nrows = a.nrows;
    const uint32_t nblocks = (nrows + 31u) >> 5;
    const uint32_t nwarps = gridDim.x * WARPS;
    double *__restrict__ sp = s_prod[warp];

    double sq = 0.0, aux = 0.0;
    for (uint32_t blk = blockIdx.x * WARPS + warp; blk < nblocks; blk += nwarps) {
        const uint32_t row_first = blk << 5;
        const uint32_t row = row_first + lane;
        const bool active = row < nrows;
        const uint32_t rlast = min(nrows, row_first + 32u) - 1u;  // last row of the block
        const uint32_t rs = a.row_ptr[active ? row : rlast + 1u];
        const uint32_t re = active ? a.row_ptr[row + 1u] : rs;
        const uint32_t b0 = __shfl_sync(0xffffffffu, rs, 0);
        const uint32_t b1 = __shfl_sync(0xffffffffu, re, (int)(rlast & 31u));
        double own = 0.0, dv = 0.0, xs = 0.0, rh = 0.0;
        // column-slab passes (launch_tile_kernel): only the last pass runs the epilogue, the others hand the running
        // row sums on through acc_out -> acc_in; the order of the additions is the single-pass order
        if (active && (a.acc_out == nullptr || (EPI == EPI_SPMV && a.acc_in == nullptr))) row_operands<EPI>(a, row, own, dv, xs, rh);
        double acc = (EPI == EPI_SPMV && a.accumulate) ? xs : 0.0;
        if (a.acc_in != nullptr) acc = active ? ld_once_f64_hint(a.acc_in + row, pol_stream) : 0.0;  // used once: evict first
        const uint32_t max_len = __reduce_max_sync(0xffffffffu, re - rs);
        if (max_len <= kLongRow) {
            for (uint32_t c0 = b0 & ~(uint32_t)(EPL - 1); c0 < b1; c0 += CH) {
                const uint32_t e = c0 + EPL * lane;  // this lane's elements e .. e+EPL-1 (32-byte aligned slices of `values`)
                double p[EPL];
#pragma unroll
                for (int q = 0; q < EPL; q++) p[q] = 0.0;
                if (e < b1) {
                    uint32_t cx[EPL];
                    double v[EPL];
                    // 256-bit loads (SASS LDG.E.256): the warp reads its slice of `values` contiguously and every
                    // 32-byte sector is requested from L2 exactly once. (Two 128-bit loads per lane touch each sector
                    // twice — with L1::no_allocate both requests reach L2, and L2 sector lookups, not DRAM bytes, are
                    // what bound this kernel on uniform-random columns: DESIGN.md §4.)
                    if constexpr (EPL == 4) {
                        asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
                                     : "=r"(cx[0]), "=r"(cx[1]), "=r"(cx[2]), "=r"(cx[3])
                                     : "l"(a.cols + e), "l"(pol_stream));
                    } else {
                        asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8], %9;"
                                     : "=r"(cx[0]), "=r"(cx[1]), "=r"(cx[2]), "=r"(cx[3]), "=r"(cx[EPL - 4]), "=r"(cx[EPL - 3]),
                                       "=r"(cx[EPL - 2]), "=r"(cx[EPL - 1])
                                     : "l"(a.cols + e), "l"(pol_stream));
                    }
#pragma unroll
                    for (int q = 0; q < EPL; q += 4)
                        asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f64 {%0,%1,%2,%3}, [%4], %5;"
                                     : "=d"(v[q]), "=d"(v[q + 1]), "=d"(v[q + 2]), "=d"(v[q + 3])
                                     : "l"(a.vals + e + q), "l"(pol_stream));
                    // elements outside [b0,b1) belong to neighbouring rows or to the zero padding: their columns are
                    // valid, their products are never summed
                    double xg[EPL];
#pragma unroll
                    for (int q = 0; q < EPL; q++) xg[q] = ld_gather(a.xin + cx[q], pol_gather);
#pragma unroll
                    for (int q = 0; q < EPL; q++) p[q] = v[q] * xg[q];
                }
#pragma unroll
                for (int q = 0; q < EPL; q += 2)
                    *reinterpret_cast<double2 *>(sp + EPL * lane + q) = make_double2(p[q], p[q + 1]);
                __syncwarp();
                // left-to-right accumulation, the order of CSRStorage::multiply_vector_add (sparse.rs:193-203)
                const uint32_t lo = max(rs, c0), hi = min(re, c0 + CH);
                for (uint32_t k = lo; k < hi; k++) acc += sp[k - c0];
                __syncwarp();
            }
        } else {
            // a block holding a very long row: one row at a time, lane-strided partial sums + shuffle tree
            // (order differs from the reference: tolerance-level parity only)
            const uint32_t nr = rlast - row_first + 1u;
            for (uint32_t i = 0; i < nr; i++) {
                const uint32_t s = __shfl_sync(0xffffffffu, rs, (int)i), t = __shfl_sync(0xffffffffu, re, (int)i);
                double part = 0.0;
                if (a.long_sum != nullptr && t - s > kLongRow) {
                    // a hub row: its sum was computed by the whole grid before this launch (long_rows_* below);
                    // every lane finds the row's slot in the ascending list of long rows
                    const uint32_t target = row_first + i;
                    uint32_t lo = 0, hi = a.nlong;
                    while (lo < hi) {
                        const uint32_t mid = (lo + hi) >> 1;
                        if (a.long_rows[mid] < target) lo = mid + 1; else hi = mid;
                    }
                    part = a.long_sum[lo];
                } else {
                    for (uint32_t k = s + lane; k < t; k += 32u)
                        part += ld_stream_f64(a.vals + k) * ld_gather(a.xin + ld_stream_u32(a.cols + k), pol_gather);
                    part = warp_sum(part);
                }
                if ((uint32_t)lane == i) acc += part;
            }
        }
        if (active) {
            if (a.acc_out != nullptr) st_stream_f64_hint(a.acc_out + row, acc, pol_stream);  // keep the slab of x in L2
            else row_epilogue<EPI>(a, row, acc, own, dv, xs, rh, sq, aux);
        }
    }
    if (EPI != EPI_SPMV && a.acc_out == nullptr) {
        const int kind = EPI == EPI_PUSH ? TAIL_TERM : (EPI == EPI_CG ? TAIL_CG_PAP : TAIL_RESID);
        grid_reduce_and_tail<NT>(sq, aux, a.ctl, a.partials, kind, a.it, a.last_in_iter, a.identity_res, a.defer_tail,
                                 a.norm_log, s_red, &s_flag, &a.px);
    }
}

template <int EPI, int NT, int EPL>
static int32_t launch_warp_one(const TileKernelArgs &a, cudaStream_t stream, int *max_grid_out) {
    static int max_grid[16] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 16) return fail(SB200_ERR_ALGORITHM, "device index %d out of range", dev);
    if (max_grid[dev] == 0) {
        int per_sm = 0, sms = 0;
        SB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, warp_kernel<EPI, NT, EPL>, NT, 0));
        SB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        if (per_sm < 1) return fail(SB200_ERR_ALGORITHM, "warp kernel does not fit on an SM");
        static int cap = [] { const char *e = getenv("SUBLINEAR_B200_WARP_CTAS"); return e ? atoi(e) : 0; }();
        if (cap > 0 && per_sm > cap) per_sm = cap;
        max_grid[dev] = per_sm * sms;  // persistent grid: a whole number of CTAs per SM (148 SMs on B200)
    }
    if (max_grid_out) {
        *max_grid_out = max_grid[dev];
        return SB200_OK;
    }
    if (a.nrows == 0 && EPI == EPI_SPMV) return SB200_OK;
    const uint32_t nblocks = (a.nrows + 31u) / 32u;
    unsigned need = (nblocks + NT / 32 - 1) / (NT / 32);
    unsigned grid = need < (unsigned)max_grid[dev] ? need : (unsigned)max_grid[dev];
    if (grid == 0) grid = 1;
    warp_kernel<EPI, NT, EPL><<<grid, NT, 0, stream>>>(a);
    SB_CUDA(cudaGetLastError());
    return SB200_OK;
}

static int32_t launch_warp_any(Epilogue epi, const TileKernelArgs &a, cudaStream_t stream, int *mg) {
    // elements per lane and chunk: 4 (default) or 8 ($SUBLINEAR_B200_WARP_EPL; measurement aid, same results bit for bit)
    static int epl = [] { const char *e = getenv("SUBLINEAR_B200_WARP_EPL"); return (e && atoi(e) == 8) ? 8 : 4; }();
    if (epl == 8) {
        switch (epi) {
            case EPI_SPMV: return launch_warp_one<EPI_SPMV, 256, 8>(a, stream, mg);
            case EPI_PUSH: return launch_warp_one<EPI_PUSH, 256, 8>(a, stream, mg);
            case EPI_CG: return launch_warp_one<EPI_CG, 256, 8>(a, stream, mg);
            default: return launch_warp_one<EPI_RESID, 256, 8>(a, stream, mg);
        }
    }
    switch (epi) {
        case EPI_SPMV: return launch_warp_one<EPI_SPMV, 256, 4>(a, stream, mg);
        case EPI_PUSH: return launch_warp_one<EPI_PUSH, 256, 4>(a, stream, mg);
        case EPI_CG: return launch_warp_one<EPI_CG, 256, 4>(a, stream, mg);
        default: return launch_warp_one<EPI_RESID, 256, 4>(a, stream, mg);
    }
}

// ---------------------------------------------------------------------------------------------------------
// the SELL-32 kernel (vectors <= 48 MB and band-local matrices, when the padded layout costs <= 25 % extra slots)
// ---------------------------------------------------------------------------------------------------------
// Measured on the warp-stream kernel (profiles/r1_warp_probe_phases.log): on uniform-random columns the x[col] gathers
// alone cost 687 us of the 808 us launch although the same gathers run at 265-287 G/s (350-377 us) in isolation. The
// difference is the L1: every in-flight gather holds a 128-byte L1 line, the line count is what shared memory leaves
// of the 256 KB array, and the driver sizes the carve-out for the maximum number of resident CTAs — 64 KB for the
// warp-stream kernel's 9 KB of product staging per CTA (bench/gather_probe.cu: 267 / 223 / 176-212 / 120 / 56 G/s at
// 256 / 192 / 124 / 60 / 28 KB of L1). This kernel therefore uses NO shared memory at all:
//   * ingest re-lays the CSR slices out in blocks of 32 consecutive rows, element k of row r at
//     sell_ptr[blk]*32 + k*32 + r (sliced ELLPACK, slice height 32 = one warp; zero-padded to the longest row of the
//     block), so lane r reads ITS OWN row's k-th column index / value with fully coalesced 128 / 256-byte warp loads —
//     no transposition through shared memory, no shuffles;
//   * lane r accumulates row r left to right in registers: the reference's order (sparse.rs:193-203), bit for bit;
//   * U stream loads, then U gathers are in flight per lane before the first add;
//   * the norm reduction goes warp shuffle -> one partial per warp in global memory -> fixed-order sum by one warp of
//     the last CTA (elected with __syncthreads_or, no shared flag).
// Padding slots hold value 0 / column 0 and are never gathered or added (k < row length), so non-finite x entries
// cannot leak into other rows.
// warp partial -> global partial array -> warp 0 of the last CTA sums all partials in index order. No shared memory.
template <int NT>
__device__ __forceinline__ void grid_reduce_and_tail_regs(double sq, double aux, LoopCtl *ctl, double *partials, int kind,
                                                          uint32_t it, int last_in_iter, int identity_res, int defer,
                                                          double *norm_log, const PeerExchange *px) {
    constexpr int WARPS = NT / 32;
    const bool p2p = px != nullptr && px->world > 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned nparts = gridDim.x * WARPS;
    if (p2p) __threadfence_system();  // this thread's stores into peer memory, before the CTA reports in
    sq = warp_sum(sq);
    if (identity_res) aux = warp_sum(aux);
    if (lane == 0) {
        partials[blockIdx.x * WARPS + warp] = sq;
        if (identity_res) partials[nparts + blockIdx.x * WARPS + warp] = aux;
        if (p2p) __threadfence_system(); else __threadfence();
    }
    __syncthreads();
    int last = 0;
    if (threadIdx.x == 0) last = atomicAdd(&ctl->ticket, 1u) == gridDim.x - 1;
    last = __syncthreads_or(last);
    if (last && warp == 0) {
        __threadfence();
        double s = 0.0, a2 = 0.0;
        for (unsigned i = lane; i < nparts; i += 32) s += __ldcg(partials + i);
        if (identity_res)
            for (unsigned i = lane; i < nparts; i += 32) a2 += __ldcg(partials + nparts + i);
        s = warp_sum(s);
        if (identity_res) a2 = warp_sum(a2);
        if (lane == 0) {
            ctl->ticket = 0;
            if (p2p) {
                __threadfence_system();
                peer_signal(ctl, *px, s, a2);  // the wait kernel that follows runs tail_logic on the global sums
            } else {
                tail_logic(ctl, kind, s, a2, it, last_in_iter, identity_res, defer, norm_log);
            }
        }
    }
}

template <int EPI, int NT, int U>
__global__ void __launch_bounds__(NT) sell_kernel(const TileKernelArgs a) {
    constexpr int WARPS = NT / 32;
    if (EPI != EPI_SPMV) {
        if (!a.force && a.ctl->alive == 0) return;  // loop already finished: no-op launch
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint64_t pol_stream = policy_evict_first();
    const uint64_t pol_gather = policy_evict_last();
    const uint32_t nrows = a.nrows;
    const uint32_t nblocks = (nrows + 31u) >> 5;
    const uint32_t nwarps = gridDim.x * WARPS;

    double sq = 0.0, aux = 0.0;
    for (uint32_t blk = blockIdx.x * WARPS + warp; blk < nblocks; blk += nwarps) {
        const uint32_t row = (blk << 5) + lane;
        const bool active = row < nrows;
        uint32_t len = 0;
        if (active) len = a.row_ptr[row + 1u] - a.row_ptr[row];
        const uint32_t off = a.sell_ptr[blk];           // warp-uniform: first 32-slot slab of the block
        const uint32_t width = a.sell_ptr[blk + 1u] - off;  // slabs = longest row of the block
        double own = 0.0, dv = 0.0, xs = 0.0, rh = 0.0;
        if (active) row_operands<EPI>(a, row, own, dv, xs, rh);
        double acc = (EPI == EPI_SPMV && a.accumulate) ? xs : 0.0;
        const size_t base = (size_t)off * 32u + (size_t)lane;
        const uint32_t *__restrict__ cp = a.sell_cols + base;
        const double *__restrict__ vp = a.sell_vals + base;
        for (uint32_t k0 = 0; k0 < width; k0 += U) {
            uint32_t c[U];
            double v[U], x[U];
#pragma unroll
            for (int u = 0; u < U; u++) {
                c[u] = 0u;
                v[u] = 0.0;
                if (k0 + u < width) {  // warp-uniform
                    c[u] = ld_stream_u32_hint(cp + (size_t)(k0 + u) * 32u, pol_stream);
                    v[u] = ld_stream_f64_hint(vp + (size_t)(k0 + u) * 32u, pol_stream);
                }
            }
#pragma unroll
            for (int u = 0; u < U; u++) {
                x[u] = 0.0;
                if (k0 + u < len) x[u] = ld_gather(a.xin + c[u], pol_gather);
            }
            // left-to-right accumulation, the order of CSRStorage::multiply_vector_add (sparse.rs:193-203)
#pragma unroll
            for (int u = 0; u < U; u++)
                if (k0 + u < len) acc += v[u] * x[u];
        }
        if (active) row_epilogue<EPI>(a, row, acc, own, dv, xs, rh, sq, aux);
    }
    if (EPI != EPI_SPMV) {
        const int kind = EPI == EPI_PUSH ? TAIL_TERM : (EPI == EPI_CG ? TAIL_CG_PAP : TAIL_RESID);
        grid_reduce_and_tail_regs<NT>(sq, aux, a.ctl, a.partials, kind, a.it, a.last_in_iter, a.identity_res, a.defer_tail,
                                      a.norm_log, &a.px);
    }
}

constexpr int kSellThreads = 256;

template <int EPI, int U>
static int32_t launch_sell_one(const TileKernelArgs &a, cudaStream_t stream, int *max_grid_out) {
    static int max_grid[16] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 16) return fail(SB200_ERR_ALGORITHM, "device index %d out of range", dev);
    if (max_grid[dev] == 0) {
        // no shared memory: ask for the smallest carve-out so the whole 256 KB array serves as L1
        SB_CUDA(cudaFuncSetAttribute(sell_kernel<EPI, kSellThreads, U>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                     cudaSharedmemCarveoutMaxL1));
        int per_sm = 0, sms = 0;
        SB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, sell_kernel<EPI, kSellThreads, U>, kSellThreads, 0));
        SB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        if (per_sm < 1) return fail(SB200_ERR_ALGORITHM, "SELL kernel does not fit on an SM");
        static int cap = [] { const char *e = getenv("SUBLINEAR_B200_SELL_CTAS"); return e ? atoi(e) : 0; }();
        if (cap > 0 && per_sm > cap) per_sm = cap;
        max_grid[dev] = per_sm * sms;  // persistent grid: a whole number of CTAs per SM (148 SMs on B200)
    }
    if (max_grid_out) {
        *max_grid_out = max_grid[dev] * (kSellThreads / 32);  // one partial per WARP (callers size `partials` from this)
        return SB200_OK;
    }
    if (a.nrows == 0 && EPI == EPI_SPMV) return SB200_OK;
    const uint32_t nblocks = (a.nrows + 31u) / 32u;
    unsigned need = (nblocks + kSellThreads / 32 - 1) / (kSellThreads / 32);
    unsigned grid = need < (unsigned)max_grid[dev] ? need : (unsigned)max_grid[dev];
    if (grid == 0) grid = 1;
    sell_kernel<EPI, kSellThreads, U><<<grid, kSellThreads, 0, stream>>>(a);
    SB_CUDA(cudaGetLastError());
    return SB200_OK;
}

template <int U>
static int32_t launch_sell_u(Epilogue epi, const TileKernelArgs &a, cudaStream_t stream, int *mg) {
    switch (epi) {
        case EPI_SPMV: return launch_sell_one<EPI_SPMV, U>(a, stream, mg);
        case EPI_PUSH: return launch_sell_one<EPI_PUSH, U>(a, stream, mg);
        case EPI_CG: return launch_sell_one<EPI_CG, U>(a, stream, mg);
        default: return launch_sell_one<EPI_RESID, U>(a, stream, mg);
    }
}

static int32_t launch_sell_any(Epilogue epi, const TileKernelArgs &a, cudaStream_t stream, int *mg) {
    // stream loads / gathers in flight per lane ($SUBLINEAR_B200_SELL_U; measurement aid, same results bit for bit)
    static int u = [] { const char *e = getenv("SUBLINEAR_B200_SELL_U"); return e ? atoi(e) : 0; }();
    switch (u) {
        case 4: return launch_sell_u<4>(epi, a, stream, mg);
        case 5: return launch_sell_u<5>(epi, a, stream, mg);
        case 10: return launch_sell_u<10>(epi, a, stream, mg);
        default: return launch_sell_u<8>(epi, a, stream, mg);
    }
}

// one-off layout pass at ingest: CSR slices -> SELL-32 slabs (warp per block of 32 rows)
__global__ void csr_to_sell_kernel(const double *__restrict__ vals, const uint32_t *__restrict__ cols,
                                   const uint32_t *__restrict__ row_ptr, uint32_t nrows,
                                   const uint32_t *__restrict__ sell_ptr, uint32_t *__restrict__ sc,
                                   double *__restrict__ sv) {
    const int lane = threadIdx.x & 31;
    const uint32_t nblocks = (nrows + 31u) >> 5;
    const uint32_t nwarps = gridDim.x * (blockDim.x >> 5);
    for (uint32_t blk = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); blk < nblocks; blk += nwarps) {
        const uint32_t row = (blk << 5) + lane;
        uint32_t rs = 0, len = 0;
        if (row < nrows) {
            rs = row_ptr[row];
            len = row_ptr[row + 1u] - rs;
        }
        const uint32_t off = sell_ptr[blk], width = sell_ptr[blk + 1u] - off;
        for (uint32_t k = 0; k < width; k++) {
            const size_t idx = ((size_t)off + k) * 32u + (size_t)lane;
            sc[idx] = k < len ? cols[rs + k] : 0u;
            sv[idx] = k < len ? vals[rs + k] : 0.0;
        }
    }
}

int32_t launch_csr_to_sell(const double *vals, const uint32_t *cols, const uint32_t *row_ptr, uint32_t nrows,
                           const uint32_t *sell_ptr, uint32_t *sell_cols, double *sell_vals, cudaStream_t stream) {
    if (nrows == 0) return SB200_OK;
    const uint32_t nblocks = (nrows + 31u) / 32u;
    unsigned grid = (nblocks + 7) / 8;
    if (grid > 148 * 16) grid = 148 * 16;
    csr_to_sell_kernel<<<grid, 256, 0, stream>>>(vals, cols, row_ptr, nrows, sell_ptr, sell_cols, sell_vals);
    SB_CUDA(cudaGetLastError());
    return SB200_OK;
}

// ---------------------------------------------------------------------------------------------------------
// column-slab split at ingest (one-off; matrix.cu build_slabs)
// ---------------------------------------------------------------------------------------------------------
struct SlabPtrs {
    uint32_t *counts[kMaxSlabs];
    const uint32_t *row_ptr[kMaxSlabs];
    uint32_t *cols[kMaxSlabs];
    double *vals[kMaxSlabs];
};

__global__ void slab_count_kernel(const uint32_t *__restrict__ cols, const uint32_t *__restrict__ row_ptr, uint32_t nrows,
                                  uint32_t slab_width, int nslabs, SlabPtrs p, int *unsorted) {
    for (uint32_t row = blockIdx.x * blockDim.x + threadIdx.x; row < nrows; row += gridDim.x * blockDim.x) {
        uint32_t cnt[kMaxSlabs] = {0u, 0u, 0u, 0u};
        uint32_t prev = 0;
        bool bad = false;
        for (uint32_t k = row_ptr[row]; k < row_ptr[row + 1]; k++) {
            const uint32_t c = cols[k];
            bad |= c < prev;
            prev = c;
            const uint32_t s = min(c / slab_width, (uint32_t)(nslabs - 1));
            cnt[s]++;
        }
        int used = 0;
        for (int s = 0; s < nslabs; s++) {
            p.counts[s][row] = cnt[s];
            used += cnt[s] != 0u;
        }
        if (bad) unsorted[0] = 1;
        if (used > 1) atomicAdd(unsorted + 1, 1);  // rows whose gathers spread over several slabs
    }
}

__global__ void slab_fill_kernel(const double *__restrict__ vals, const uint32_t *__restrict__ cols,
                                 const uint32_t *__restrict__ row_ptr, uint32_t nrows, uint32_t slab_width, int nslabs,
                                 SlabPtrs p) {
    for (uint32_t row = blockIdx.x * blockDim.x + threadIdx.x; row < nrows; row += gridDim.x * blockDim.x) {
        uint32_t pos[kMaxSlabs];
        for (int s = 0; s < nslabs; s++) pos[s] = p.row_ptr[s][row];
        for (uint32_t k = row_ptr[row]; k < row_ptr[row + 1]; k++) {  // in CSR order: the order inside a slab row is kept
            const uint32_t c = cols[k];
            const uint32_t s = min(c / slab_width, (uint32_t)(nslabs - 1));
            p.cols[s][pos[s]] = c;
            p.vals[s][pos[s]] = vals[k];
            pos[s]++;
        }
    }
}

int32_t launch_slab_count(const uint32_t *cols, const uint32_t *row_ptr, uint32_t nrows, uint32_t slab_width, int nslabs,
                          uint32_t *const *counts, int *unsorted, cudaStream_t stream) {
    SlabPtrs p{};
    for (int s = 0; s < nslabs; s++) p.counts[s] = counts[s];
    unsigned grid = (nrows + 255) / 256;
    if (grid > 148 * 16) grid = 148 * 16;
    if (grid == 0) grid = 1;
    slab_count_kernel<<<grid, 256, 0, stream>>>(cols, row_ptr, nrows, slab_width, nslabs, p, unsorted);
    SB_CUDA(cudaGetLastError());
    return SB200_OK;
}

int32_t launch_slab_fill(const double *vals, const uint32_t *cols, const uint32_t *row_ptr, uint32_t nrows,
                         uint32_t slab_width, int nslabs, const uint32_t *const *slab_row_ptr, uint32_t *const *slab_cols,
                         double *const *slab_vals, cudaStream_t stream) {
    SlabPtrs p{};
    for (int s = 0; s < nslabs; s++) {
        p.row_ptr[s] = slab_row_ptr[s];
        p.cols[s] = slab_cols[s];
        p.vals[s] = slab_vals[s];
    }
    unsigned grid = (nrows + 255) / 256;
    if (grid > 148 * 16) grid = 148 * 16;
    if (grid == 0) grid = 1;
    slab_fill_kernel<<<grid, 256, 0, stream>>>(vals, cols, row_ptr, nrows, slab_width, nslabs, p);
    SB_CUDA(cudaGetLastError());
    return SB200_OK;
}

// exclusive prefix sum of n u32 counts in place (data has n + 1 entries: data[n] receives the total), single CTA:
// an ingest-time helper, not a hot path (10 M rows: ~1 ms)
__global__ void __launch_bounds__(1024) exclusive_scan_u32_kernel(uint32_t *data, uint64_t n, unsigned long long *total) {
    __shared__ unsigned long long s_warp[32];
    __shared__ unsigned long long s_carry;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) s_carry = 0ull;
    __syncthreads();
    for (uint64_t base = 0; base < n; base += 1024) {
        const uint64_t i = base + threadIdx.x;
        const unsigned long long v = i < n ? data[i] : 0ull;
        unsigned long long x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) s_warp[warp] = x;
        __syncthreads();
        if (warp == 0) {
            unsigned long long w = s_warp[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned long long y = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += y;
            }
            s_warp[lane] = w;  // inclusive over warps
        }
        __syncthreads();
        const unsigned long long before = s_carry + (warp ? s_warp[warp - 1] : 0ull) + (x - v);
        if (i < n) data[i] = (uint32_t)before;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = before + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        data[n] = (uint32_t)s_carry;
        *total = s_carry;
    }
}

int32_t device_exclusive_scan_u32(uint32_t *data, uint64_t n, uint64_t *total, cudaStream_t stream) {
    DevBuf<unsigned long long> t;
    SB_TRY(t.alloc(1));
    exclusive_scan_u32_kernel<<<1, 1024, 0, stream>>>(data, n, t.p);
    SB_CUDA(cudaGetLastError());
    unsigned long long h = 0;
    SB_CUDA(cudaMemcpyAsync(&h, t.p, 8, cudaMemcpyDeviceToHost, stream));
    SB_CUDA(cudaStreamSynchronize(stream));
    if (total) *total = h;
    return SB200_OK;
}

// ---------------------------------------------------------------------------------------------------------
// K4: setup pass — dominance, diagonal, D^-1   (ref src/solver/neumann.rs:162-188, src/matrix/mod.rs:467-485)
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void atomic_min_u64(unsigned long long *addr, unsigned long long v) { atomicMin(addr, v); }

__global__ void setup_rows_kernel(const double *__restrict__ vals, const uint32_t *__restrict__ cols,
                                  const uint32_t *__restrict__ row_ptr, uint32_t nrows, uint32_t row_base, int compat_diag,
                                  SetupOut o) {
    for (uint32_t lrow = blockIdx.x * blockDim.x + threadIdx.x; lrow < nrows; lrow += gridDim.x * blockDim.x) {
        const uint32_t rs = row_ptr[lrow], re = row_ptr[lrow + 1];
        const uint32_t row = row_base + lrow;  // global index of this row = column of its diagonal entry
        double diag_last = 0.0, diag_sum = 0.0, off = 0.0;
        bool has = false;
        for (uint32_t k = rs; k < re; k++) {
            const uint32_t c = cols[k];
            const double v = vals[k];
            if (c == row) {
                diag_last = fabs(v);  // `diagonal = value.abs()` is overwritten per entry (mod.rs:474-476)
                diag_sum += v;
                has = true;
            } else {
                off += fabs(v);
                if (o.col_off) atomicAdd(o.col_off + c, fabs(v));
            }
        }
        if (o.col_diag && has) o.col_diag[row] = diag_last;
        if (diag_last < off) atomic_min_u64(o.first_bad_dd, lrow);  // mod.rs:480
        double d = diag_sum;
        if (compat_diag && has) {
            // CSRStorage::get (sparse.rs:142-155): bisection over the (column-sorted) row; with duplicated
            // diagonal entries it returns whichever one the probe sequence meets first.
            uint32_t lo = rs, hi = re;
            bool found = false;
            while (lo < hi) {
                const uint32_t mid = lo + (hi - lo) / 2;
                const uint32_t c = cols[mid];
                if (c == row) {
                    d = vals[mid];
                    found = true;
                    break;
                }
                if (c < row) lo = mid + 1; else hi = mid;
            }
            if (!found) has = false;  // unsorted row: the reference's binary search would miss it too
        }
        if (!has || fabs(d) < 1e-14) {  // neumann.rs:174-187
            atomic_min_u64(o.first_bad_diag, lrow);
            o.dinv[lrow] = 0.0;
        } else {
            o.dinv[lrow] = 1.0 / d;
        }
        if (o.min_factor_bits && off > 0.0) {
            // positive doubles order like their bit patterns
            atomicMin(reinterpret_cast<unsigned long long *>(o.min_factor_bits),
                      (unsigned long long)__double_as_longlong(diag_last / off));
        }
    }
}

int32_t launch_setup_rows(const double *vals, const uint32_t *cols, const uint32_t *row_ptr, uint32_t nrows,
                          uint32_t row_base, int compat_diag, SetupOut out, cudaStream_t stream) {
    if (nrows == 0) return SB200_OK;
    unsigned grid = (nrows + 255) / 256;
    if (grid > 148 * 16) grid = 148 * 16;
    setup_rows_kernel<<<grid, 256, 0, stream>>>(vals, cols, row_ptr, nrows, row_base, compat_diag, out);
    SB_CUDA(cudaGetLastError());
    return SB200_OK;
}

__global__ void col_dominance_kernel(const double *__restrict__ col_diag, const double *__restrict__ col_off, uint32_t n,
                                     unsigned long long *first_bad) {
    for (uint32_t c = blockIdx.x * blockDim.x + threadIdx.x; c < n; c += gridDim.x * blockDim.x)
        if (col_diag[c] < col_off[c]) atomic_min_u64(first_bad, c);
}

int32_t launch_col_dominance(const double *col_diag, const double *col_off, uint32_t n, unsigned long long *first_bad,
                             cudaStream_t stream) {
    if (n == 0) return SB200_OK;
    unsigned grid = (n + 255) / 256;
    if (grid > 148 * 16) grid = 148 * 16;
    col_dominance_kernel<<<grid, 256, 0, stream>>>(col_diag, col_off, n, first_bad);
    SB_CUDA(cudaGetLastError());
    return SB200_OK;
}

// ---------------------------------------------------------------------------------------------------------
// iteration 0: scaled rhs, first term, first accumulation (ref neumann.rs:191-211 and compute_next_term k=0)
// ---------------------------------------------------------------------------------------------------------
constexpr int kInitThreads = 256;

__global__ void __launch_bounds__(kInitThreads) init_state_kernel(const InitArgs a) {
    __shared__ double s_red[kInitThreads / 32];
    __shared__ int s_flag;
    double sq = 0.0, aux = 0.0;
    for (uint32_t i = blockIdx.x * kInitThreads + threadIdx.x; i < a.n; i += gridDim.x * kInitThreads) {
        const double dv = a.dinv[i];
        const double bi = a.b[i];
        const double c = bi * dv;  // rhs = b o D^-1 (neumann.rs:191-194)
        if (a.c_out) a.c_out[i] = c;
        double t0, base;
        if (a.compat) {
            t0 = c;                          // current_term = rhs.clone()      (neumann.rs:211)
            base = a.x0 ? a.x0[i] : c;       // solution = initial_guess or rhs (neumann.rs:197-208)
        } else {
            t0 = a.ax0 ? (bi - a.ax0[i]) * dv : c;  // t0 = D^-1 (b - A x0)
            base = a.x0 ? a.x0[i] : 0.0;
        }
        a.t_out[i] = t0;
        const double x_new = a.skip_term0 ? base : base + t0;  // k = 0: solution += term (neumann.rs:264-266)
        a.x_out[i] = x_new;
        if (a.px.world > 1) {
            const size_t g = (size_t)a.row_base + i;
            for (int p = 0; p < a.px.world; p++) {
                if (p != a.px.rank) a.px.t_out[p][g] = t0;
                if (a.px.x_out[p]) a.px.x_out[p][g] = x_new;
            }
        }
        sq += t0 * t0;
        if (a.identity_res) {
            const double r = t0 / dv;
            aux += r * r;
        }
    }
    grid_reduce_and_tail<kInitThreads>(sq, aux, a.ctl, a.partials, TAIL_TERM, 0u, a.last_in_iter, a.identity_res,
                                       a.defer_tail, a.norm_log, s_red, &s_flag, &a.px);
}

int init_state_grid() { return 148 * 4; }

int32_t launch_init_state(const InitArgs &a, cudaStream_t stream) {
    unsigned grid = (a.n + kInitThreads - 1) / kInitThreads;
    if (grid > (unsigned)init_state_grid()) grid = init_state_grid();
    if (grid == 0) grid = 1;
    init_state_kernel<<<grid, kInitThreads, 0, stream>>>(a);
    SB_CUDA(cudaGetLastError());
    return SB200_OK;
}

// ---------------------------------------------------------------------------------------------------------
// small vector passes of the SolverAlgorithm state interface (csrc/state.cu)
// ---------------------------------------------------------------------------------------------------------
// op 0: term 0 of compute_next_term (neumann.rs:264-271 with terms_computed == 0): x += t, ||t||^2 -> TAIL_TERM(it = 0)
// op 1: ||v||^2 -> ctl->red[0] (utils::l2_norm, solver/mod.rs:369-371)
__global__ void __launch_bounds__(kInitThreads) state_vec_kernel(int op, const double *__restrict__ t, double *x, uint64_t n,
                                                                 LoopCtl *ctl, double *partials) {
    __shared__ double s_red[kInitThreads / 32];
    __shared__ int s_flag;
    double sq = 0.0;
    for (uint64_t i = blockIdx.x * (uint64_t)kInitThreads + threadIdx.x; i < n; i += (uint64_t)gridDim.x * kInitThreads) {
        const double ti = t[i];
        if (op == 0) x[i] = x[i] + ti;
        sq += ti * ti;
    }
    grid_reduce_and_tail<kInitThreads>(sq, 0.0, ctl, partials, op == 0 ? TAIL_TERM : TAIL_NONE, 0u, 0, 0, op == 1, nullptr,
                                       s_red, &s_flag);
}

int32_t launch_state_vec(int op, const double *t, double *x, uint64_t n, LoopCtl *ctl, double *partials,
                         cudaStream_t stream) {
    uint64_t g = (n + kInitThreads - 1) / kInitThreads;
    unsigned grid = g > (uint64_t)init_state_grid() ? (unsigned)init_state_grid() : (unsigned)(g ? g : 1);
    state_vec_kernel<<<grid, kInitThreads, 0, stream>>>(op, t, x, n, ctl, partials);
    SB_CUDA(cudaGetLastError());
    return SB200_OK;
}

// update_rhs (neumann.rs:436-462): the (index, delta) pairs are applied IN ORDER by one thread — the reference's loop
// is sequential and an index may repeat; the lists are small by nature (an incremental update). b += delta,
// rhs += delta * dinv; `also` (the solution in ref_compat, the restarted term in correct mode) takes the scaled delta too.
__global__ void update_rhs_kernel(const uint64_t *__restrict__ idx, const double *__restrict__ delta, uint64_t count,
                                  const double *__restrict__ dinv, double *b, double *rhs, double *also) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    for (uint64_t k = 0; k < count; k++) {
        const uint64_t i = idx[k];
        const double scaled = delta[k] * dinv[i];  // :448
        rhs[i] += scaled;                          // :449
        b[i] += delta[k];
        also[i] += scaled;                         // :453 (solution) / restarted term
    }
}

int32_t launch_update_rhs(const uint64_t *idx, const double *delta, uint64_t count, const double *dinv, double *b,
                          double *rhs, double *also, cudaStream_t stream) {
    if (count == 0) return SB200_OK;
    update_rhs_kernel<<<1, 32, 0, stream>>>(idx, delta, count, dinv, b, rhs, also);
    SB_CUDA(cudaGetLastError());
    return SB200_OK;
}

__global__ void scale_kernel(double *v, uint64_t n, double f) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        v[i] *= f;
}

// Row-partitioned runs: the tile kernel only published this rank's partial sums (defer_tail); after the
// allreduce every rank holds the global sums and takes the same decision here.
__global__ void dist_tail_kernel(LoopCtl *c, int kind, uint32_t it, int last_in_iter, int identity_res, int force,
                                 double *norm_log) {
    if (c->alive == 0 && !force) return;  // dead loop: red[] only holds re-reduced garbage
    tail_logic(c, kind, c->red[0], c->red[1], it, last_in_iter, identity_res, 0, norm_log);
}

int32_t launch_dist_tail(LoopCtl *ctl, int kind, uint32_t it, int last_in_iter, int identity_res, int force,
                         double *norm_log, cudaStream_t stream) {
    dist_tail_kernel<<<1, 1, 0, stream>>>(ctl, kind, it, last_in_iter, identity_res, force, norm_log);
    SB_CUDA(cudaGetLastError());
    return SB200_OK;
}

// P2P exchange, consumer side: one warp. Lane r waits for rank r's flag of the current exchange; lane 0 then adds the
// ranks' partial sums in rank order (every rank computes the same bits) and runs the loop logic on them.
__global__ void peer_wait_kernel(LoopCtl *c, const unsigned long long *flags, const double *slots, int world,
                                 unsigned long long epoch_base, int kind, uint32_t it, int last_in_iter, int identity_res,
                                 int force, double *norm_log) {
    if (c->alive == 0 && !force) return;  // dead loop: nobody signalled, nothing to wait for
    const unsigned long long e = epoch_base + c->xchg + 1ull;
    const int lane = threadIdx.x;
    bool ok = true;
    if (lane < world) {
        const long long t0 = clock64();
        while (ld_acquire_sys_u64(flags + lane) < e) {
            if (clock64() - t0 > 40000000000ll) {  // ~20 s: a peer died; do not hang the GPU
                ok = false;
                break;
            }
            __nanosleep(200);
        }
    }
    ok = __all_sync(0xffffffffu, ok);
    if (lane == 0) {
        if (!ok) {
            c->peer_timeout = 1;
            c->alive = 0;
            return;
        }
        const unsigned par = (unsigned)(e & 1ull);
        double s = 0.0, a = 0.0;
        for (int r = 0; r < world; r++) {
            s += ld_relaxed_sys_f64(slots + ((size_t)par * world + r) * 2);
            a += ld_relaxed_sys_f64(slots + ((size_t)par * world + r) * 2 + 1);
        }
        c->xchg += 1;
        if (kind != TAIL_NONE) tail_logic(c, kind, s, a, it, last_in_iter, identity_res, 0, norm_log);
    }
}

int32_t launch_peer_wait(LoopCtl *ctl, const unsigned long long *flags_local, const double *slots_local, int world,
                         unsigned long long epoch_base, int kind, uint32_t it, int last_in_iter, int identity_res,
                         int force, double *norm_log, cudaStream_t stream) {
    peer_wait_kernel<<<1, 32, 0, stream>>>(ctl, flags_local, slots_local, world, epoch_base, kind, it, last_in_iter,
                                          identity_res, force, norm_log);
    SB_CUDA(cudaGetLastError());
    return SB200_OK;
}

// P2P exchange, plain publish: src[0..n) -> dst_p[offset .. offset+n) on every rank, then signal (no sums).
struct PublishDst {
    double *p[kMaxPeers];
};
__global__ void __launch_bounds__(256) peer_publish_kernel(const double *__restrict__ src, uint64_t n, uint64_t offset,
                                                           PublishDst dst, LoopCtl *ctl, PeerExchange px, int force) {
    __shared__ int s_flag;
    if (ctl->alive == 0 && !force) return;
    for (uint64_t i = blockIdx.x * 256ull + threadIdx.x; i < n; i += (uint64_t)gridDim.x * 256ull) {
        const double v = src[i];
        for (int p = 0; p < px.world; p++) dst.p[p][offset + i] = v;
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        unsigned t = atomicAdd(&ctl->ticket, 1u);
        s_flag = (t == gridDim.x - 1);
        if (s_flag) {
            ctl->ticket = 0;
            __threadfence_system();
            peer_signal(ctl, px, 0.0, 0.0);
        }
    }
}

int32_t launch_peer_publish(const double *src, uint64_t n, uint64_t offset, double *const *dst, LoopCtl *ctl,
                            const PeerExchange &px, int force, cudaStream_t stream) {
    PublishDst d{};
    for (int p = 0; p < px.world; p++) d.p[p] = dst[p];
    uint64_t g = (n + 255) / 256;
    unsigned grid = g > 148ull * 4 ? 148u * 4 : (unsigned)(g ? g : 1);
    peer_publish_kernel<<<grid, 256, 0, stream>>>(src, n, offset, d, ctl, px, force);
    SB_CUDA(cudaGetLastError());
    return SB200_OK;
}

// ---------------------------------------------------------------------------------------------------------
// conjugate gradient vector passes (ref src/optimized_solver.rs:202-215, 240-260). HBM-bound streaming kernels:
// phase 1 moves 48 B/row, phase 2 24 B/row; products and sums stay separate IEEE operations (-fmad=false).
// ---------------------------------------------------------------------------------------------------------
constexpr int kCgThreads = 256;

__global__ void __launch_bounds__(kCgThreads) cg_vec_kernel(const CgVecArgs a) {
    __shared__ double s_red[kCgThreads / 32];
    __shared__ int s_flag;
    if (a.phase != 0 && a.ctl->alive == 0) return;  // loop already finished: no-op launch
    const uint64_t stride = (uint64_t)gridDim.x * kCgThreads;
    double sq = 0.0;
    if (a.phase == 0) {
        for (uint64_t i = blockIdx.x * (uint64_t)kCgThreads + threadIdx.x; i < a.n; i += stride) {
            const double bi = a.b[i];
            a.x[i] = 0.0;
            a.r[i] = bi;
            a.p[i] = bi;
            sq += bi * bi;
        }
        grid_reduce_and_tail<kCgThreads>(sq, 0.0, a.ctl, a.partials, TAIL_CG_INIT, 0u, 0, 0, 0, nullptr, s_red, &s_flag);
    } else if (a.phase == 1) {
        const double alpha = a.ctl->cg_alpha;
        for (uint64_t i = blockIdx.x * (uint64_t)kCgThreads + threadIdx.x; i < a.n; i += stride) {
            a.x[i] = a.x[i] + alpha * a.p[i];        // x += alpha p   (:241-243)
            const double ri = a.r[i] - alpha * a.ap[i];  // r -= alpha ap  (:246-248)
            a.r[i] = ri;
            sq += ri * ri;                           // rsnew          (:250-253)
        }
        grid_reduce_and_tail<kCgThreads>(sq, 0.0, a.ctl, a.partials, TAIL_CG_RS, 0u, 0, 0, 0, nullptr, s_red, &s_flag);
    } else {
        const double beta = a.ctl->cg_beta;
        for (uint64_t i = blockIdx.x * (uint64_t)kCgThreads + threadIdx.x; i < a.n; i += stride)
            a.p[i] = a.r[i] + beta * a.p[i];         // p = r + beta p (:258-260)
    }
}

int cg_vec_grid() { return 148 * 8; }

int32_t launch_cg_vec(const CgVecArgs &a, cudaStream_t stream) {
    uint64_t g = (a.n + kCgThreads - 1) / kCgThreads;
    unsigned grid = g > (uint64_t)cg_vec_grid() ? (unsigned)cg_vec_grid() : (unsigned)(g ? g : 1);
    cg_vec_kernel<<<grid, kCgThreads, 0, stream>>>(a);
    SB_CUDA(cudaGetLastError());
    return SB200_OK;
}

int32_t launch_scale(double *v, uint64_t n, double factor, cudaStream_t stream) {
    if (n == 0) return SB200_OK;
    uint64_t g = (n + 255) / 256;
    unsigned grid = g > 148ull * 16 ? 148u * 16 : (unsigned)g;
    scale_kernel<<<grid, 256, 0, stream>>>(v, n, factor);
    SB_CUDA(cudaGetLastError());
    return SB200_OK;
}

}  // namespace sb200
