// kernels_vec.cu — the vector passes around the matrix kernels: iteration 0 of the Neumann loop, the state-interface
// passes, the CG vector updates, scaling, and the consumer side of the multi-GPU exchange.
#include "device_util.cuh"

namespace sb200 {

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
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();  // cumulative: after the barrier it orders the whole CTA's peer stores before the ticket
        unsigned t = atomicAdd(&ctl->ticket, 1u);
        s_flag = (t == gridDim.x - 1);
        if (s_flag) {
            ctl->ticket = 0;
            __threadfence_system();
            peer_signal(ctl, px, 0.0, 0.0);
            peer_consume(ctl, px, TAIL_NONE, 0u, 0, 0, nullptr);
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
