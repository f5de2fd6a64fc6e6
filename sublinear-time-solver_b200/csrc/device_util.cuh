// device_util.cuh — PTX helpers, deterministic reductions and the device-side loop control shared by the kernel
// translation units (kernels.cu, kernels_tile.cu, kernels_vec.cu, kernels_ingest.cu). Internal; nothing here is part of the ABI.
#pragma once

#include "common.hpp"

namespace sb200 {

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
__device__ __forceinline__ void st_relaxed_sys_u64(unsigned long long *p, unsigned long long v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
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
    // release = ONE system-scope fence, then relaxed flag stores (a st.release per peer pays a fence + NVLink round trip
    // each, serialised on the critical path of every exchange)
    __threadfence_system();
    for (int p = 0; p < px.world; p++) st_relaxed_sys_u64(px.flags[p] + px.rank, e);
}

// P2P consume, run by the thread that just signalled (the last CTA of a kernel): wait until every rank has raised its flag
// for the current exchange, add the ranks' partial sums in rank order (every rank computes the same bits) and run the loop
// logic on the global sums. Folding the wait into the producing kernel saves the separate one-warp wait launch per
// exchange (round 1: ~5 launches per term at G = 8). No circular wait: a rank signals exchange k before it waits for it,
// and its kernel k only started after every rank had signalled k - 1.
__device__ __forceinline__ void peer_consume(LoopCtl *c, const PeerExchange &px, int kind, uint32_t it, int last_in_iter,
                                             int identity_res, double *norm_log) {
    const unsigned long long e = px.epoch_base + c->xchg + 1ull;
    const unsigned long long *flags = px.flags[px.rank];
    const long long t0 = clock64();
    for (int r = 0; r < px.world; r++) {
        while (ld_acquire_sys_u64(flags + r) < e) {
            if (clock64() - t0 > 40000000000ll) {  // ~20 s: a peer died; do not hang the GPU
                c->peer_timeout = 1;
                c->alive = 0;
                return;
            }
            __nanosleep(100);
        }
    }
    const unsigned par = (unsigned)(e & 1ull);
    const double *slots = px.slots[px.rank];
    double s = 0.0, a = 0.0;
    for (int r = 0; r < px.world; r++) {
        s += ld_relaxed_sys_f64(slots + ((size_t)par * px.world + r) * 2);
        a += ld_relaxed_sys_f64(slots + ((size_t)par * px.world + r) * 2 + 1);
    }
    c->xchg += 1;
    if (kind != TAIL_NONE) tail_logic(c, kind, s, a, it, last_in_iter, identity_res, 0, norm_log);
}

// CTA partial -> global partial array -> the last CTA to arrive sums all partials in index order.
template <int NT>
__device__ __forceinline__ void grid_reduce_and_tail(double sq, double aux, LoopCtl *ctl, double *partials, int kind,
                                                     uint32_t it, int last_in_iter, int identity_res, int defer,
                                                     double *norm_log, double *s_red, int *s_flag,
                                                     const PeerExchange *px = nullptr) {
    const bool p2p = px != nullptr && px->world > 1;
    // stores into peer memory: ordered by the CTA barriers inside block_sum + ONE cumulative system-scope fence by thread 0
    // before the CTA reports in (a fence per thread — 9 000 MEMBAR.SYS per launch — is what round 1 paid here)
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
                peer_signal(ctl, *px, s, a);
                peer_consume(ctl, *px, kind, it, last_in_iter, identity_res, norm_log);  // global sums -> loop logic
            } else {
                tail_logic(ctl, kind, s, a, it, last_in_iter, identity_res, defer, norm_log);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// per-row epilogue shared by the warp-stream, the SELL and the fused column-slab kernel: `acc` = (A xin)_row
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

}  // namespace sb200
