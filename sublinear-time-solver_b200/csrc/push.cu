// push.cu — forward / backward push as a sparse-frontier device algorithm, and the TS solver's
// forward push for A x = b.
//
// Reference: ForwardPushSolver::{solve_single_source, solve_multi_source, solve_with_target} (src/solver/forward_push.rs:
// 66-290), BackwardPushSolver::{solve_single_target, solve_multi_target, solve_with_source, combine_with_forward}
// (src/solver/backward_push.rs:66-336), BidirectionalPushSolver (backward_push.rs:338-410) over PushGraph::from_matrix
// (src/graph/adjacency.rs:199-277), and SublinearSolver.solveForwardPush (src/core/solver.ts:437-522).
//
// The reference pops one node at a time from a priority queue and touches only that node's edges: work proportional to
// the frontier, not to the graph. A push only reads the node's own residual, so any set of queued nodes can be pushed
// together; the device formulation keeps the work local and is deterministic without floating-point atomics:
//   candidates Q  : the nodes whose residual changed in the previous round (ascending, unique) — a superset of what the
//                   reference's queue would hold
//   flag + scan   : q in Q is pushed when r >= eps * max(deg, 1) (and the queue's admission threshold); exclusive scans
//                   give every pushed node its rank (max_pushes cuts the rank, not the round) and its slot range
//   apply + expand: est += alpha r; remaining = (1 - alpha) r; r = 0; one (neighbour, (remaining * w) / deg) pair per edge
//                   in the push direction — only the frontier's edges are read (a node without edges keeps its mass)
//   sort + reduce : stable radix sort of the pairs by neighbour (cub::DeviceRadixSort), then ONE thread per distinct
//                   neighbour adds its contributions onto the residual in (pushing node, edge) order; the distinct
//                   neighbours are the next round's Q
// Rounds whose frontier covers a large part of the graph (pairs > nnz / 4) run as the dense select + SpMV round on the
// library's multiply_vector_add kernel instead (round 1's formulation), and the walk drops back to the sparse form
// when the frontier shrinks.
// Invariants kept from the sequential algorithm: estimates and residuals stay >= 0, forward mass sum(est) + sum(res)
// is conserved, at exit every node has r < eps * max(deg, 1) — which bounds |estimate - PPR| exactly as it does for
// the reference's order. push_count never exceeds max_pushes. The adaptive queue threshold (x1.1 / x0.9 every 1 000
// pushes by queue length, src/graph/mod.rs:204-212) is applied per 1 000 pushes with |Q| as the queue length.
#include <cub/device/device_radix_sort.cuh>

#include <algorithm>
#include <cmath>
#include <cstring>

#include "matrix.hpp"

using namespace sb200;

struct sb200_push_graph {
    int device = 0;
    uint64_t n = 0, nnz = 0;
    DevBuf<uint32_t> d_ptr, d_col;    // adjacency: out-edges of u in [ptr[u], ptr[u+1])
    DevBuf<double> d_w;
    DevBuf<uint32_t> d_tptr, d_tcol;  // transpose: in-edges (predecessors) of v, ascending source order (mod.rs:93-127)
    DevBuf<double> d_tw;
    DevBuf<double> d_deg, d_rdeg;     // out-degree (row sums) / in-degree (column sums), adjacency.rs:214-215
    std::vector<double> h_deg, h_rdeg;
    // host copies for the lazily built dense-round matrices
    std::vector<uint64_t> h_ptr, h_tptr;
    std::vector<uint32_t> h_col, h_tcol;
    std::vector<double> h_w, h_tw;
    std::mutex mu;
    sb200_matrix *fwd = nullptr;      // M[v][u] = w_uv / deg_out(u)            (dense forward round)
    sb200_matrix *bwd = nullptr;      // M[p][v] = w_pv / max(deg_out(p), 1)    (dense backward round)
    cudaStream_t stream = nullptr;
    ~sb200_push_graph() {
        matrix_release(fwd);
        matrix_release(bwd);
        if (stream) cudaStreamDestroy(stream);
    }
};

namespace {

constexpr int kThreads = 256;

unsigned grid_for(uint64_t items, unsigned per_cta = kThreads) {
    uint64_t g = (items + per_cta - 1) / per_cta;
    return (unsigned)std::min<uint64_t>(std::max<uint64_t>(g, 1), 148ull * 16);
}

// what a push of node v does in one direction (forward: along out-edges; backward: to predecessors)
struct PushView {
    const uint32_t *ptr, *col;  // expansion CSR
    const double *w;
    const double *tdeg;         // degree the thresholds use (out-degree forward, in-degree backward)
    const double *deg;          // out-degrees (the backward transition divides by the predecessor's)
    int backward;
};

// ---- sparse round -----------------------------------------------------------------------------------------------
// flag[q] = node Q[q] is pushed this round; cnt[q] = pairs it emits (its edges, or 1 to stay a candidate without edges)
__global__ void __launch_bounds__(kThreads) push_flag_kernel(const uint32_t *__restrict__ Q, uint32_t nq,
                                                             const double *__restrict__ r, PushView g, double eps, double thr,
                                                             uint32_t *__restrict__ flag, uint32_t *__restrict__ cnt) {
    for (uint32_t q = blockIdx.x * kThreads + threadIdx.x; q < nq; q += gridDim.x * kThreads) {
        const uint32_t v = Q[q];
        const double rv = r[v], d = g.tdeg[v], dm = fmax(d, 1.0);
        // pop-time test (forward_push.rs:97-99), push_node's guard (:185), the queue's admission test (mod.rs:171-175)
        const bool push = rv > 0.0 && !(rv < eps * dm) && rv / dm >= thr;
        flag[q] = push ? 1u : 0u;
        cnt[q] = push ? (d > 0.0 ? g.ptr[v + 1] - g.ptr[v] : 1u) : 0u;
    }
}

// warp per candidate: apply the push (ranks below `budget` only) and write the node's pairs
__global__ void __launch_bounds__(kThreads) push_apply_kernel(const uint32_t *__restrict__ Q, uint32_t nq,
                                                              const uint32_t *__restrict__ flag, const uint32_t *__restrict__ rank,
                                                              const uint32_t *__restrict__ off, uint32_t budget,
                                                              double *__restrict__ r, double *__restrict__ est,
                                                              unsigned char *__restrict__ visited, PushView g, double alpha,
                                                              uint32_t *__restrict__ pkey, double *__restrict__ pval,
                                                              unsigned long long *__restrict__ nvisited) {
    const int lane = threadIdx.x & 31;
    const uint32_t nwarps = gridDim.x * (kThreads / 32);
    for (uint32_t q = blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5); q < nq; q += nwarps) {
        if (!flag[q] || rank[q] >= budget) continue;
        const uint32_t v = Q[q];
        const double rv = r[v], d = g.tdeg[v];
        const double remaining = (1.0 - alpha) * rv;  // forward_push.rs:194
        const uint32_t base = off[q];
        __syncwarp();
        if (lane == 0) {
            est[v] = est[v] + alpha * rv;             // :190-191
            r[v] = d > 0.0 ? 0.0 : remaining;         // :195 / :210-214 (no edges: the mass stays on the node)
            if (!visited[v]) {
                visited[v] = 1;
                atomicAdd(nvisited, 1ull);
            }
            if (!(d > 0.0)) {                         // stays a candidate: re-queued by push_if_threshold(node, .., 1.0)
                pkey[base] = v;
                pval[base] = 0.0;
            }
        }
        if (d > 0.0) {
            const uint32_t s = g.ptr[v], e = g.ptr[v + 1];
            const double dv = g.deg[v];
            for (uint32_t k = s + lane; k < e; k += 32) {
                const uint32_t nb = g.col[k];
                // forward: remaining * weight / deg_out(v) (:201-203); backward: remaining * (weight / max(deg_out(p), 1))
                // (backward_push.rs:201-205)
                const double c = g.backward ? remaining * (g.w[k] / fmax(g.deg[nb], 1.0)) : (remaining * g.w[k]) / dv;
                pkey[base + (k - s)] = nb;
                pval[base + (k - s)] = c;
            }
        }
    }
}

__global__ void __launch_bounds__(kThreads) head_flag_kernel(const uint32_t *__restrict__ key, uint32_t np,
                                                             uint32_t *__restrict__ head) {
    for (uint32_t i = blockIdx.x * kThreads + threadIdx.x; i < np; i += gridDim.x * kThreads)
        head[i] = (i == 0 || key[i] != key[i - 1]) ? 1u : 0u;
}

// one thread per distinct neighbour: its contributions, in (pushing node, edge) order, onto the residual
// scale: +1 (graph push) / -1 is folded into the values by the caller (A x = b push)
__global__ void __launch_bounds__(kThreads) reduce_apply_kernel(const uint32_t *__restrict__ key, const double *__restrict__ val,
                                                                const uint32_t *__restrict__ head,
                                                                const uint32_t *__restrict__ slot, uint32_t np,
                                                                double *__restrict__ r, uint32_t *__restrict__ Qnext) {
    for (uint32_t i = blockIdx.x * kThreads + threadIdx.x; i < np; i += gridDim.x * kThreads) {
        if (!head[i]) continue;
        const uint32_t v = key[i];
        double acc = r[v];
        uint32_t j = i;
        do {
            acc += val[j];  // residual[neighbor] += mass_to_transfer (:204), one contribution at a time
            j++;
        } while (j < np && key[j] == v);
        r[v] = acc;
        Qnext[slot[i]] = v;
    }
}

// ---- dense round (large frontiers) --------------------------------------------------------------------------------
// counters[0] = nodes pushed this round, counters[1] = nodes pushed for the first time
__global__ void __launch_bounds__(kThreads) push_select_kernel(double *__restrict__ r, double *__restrict__ est,
                                                               double *__restrict__ carry, const double *__restrict__ deg,
                                                               unsigned char *__restrict__ visited, uint64_t n, double alpha,
                                                               double eps, double thr, unsigned long long *counters) {
    __shared__ unsigned s_cnt[2];
    if (threadIdx.x < 2) s_cnt[threadIdx.x] = 0u;
    __syncthreads();
    unsigned pushed = 0, fresh = 0;
    for (uint64_t i = blockIdx.x * (uint64_t)kThreads + threadIdx.x; i < n; i += (uint64_t)gridDim.x * kThreads) {
        const double ri = r[i], d = deg[i], dm = fmax(d, 1.0);
        double c = 0.0;
        if (ri > 0.0 && !(ri < eps * dm) && ri / dm >= thr) {
            est[i] = est[i] + alpha * ri;
            const double remaining = (1.0 - alpha) * ri;
            if (d > 0.0) {
                c = remaining;
                r[i] = 0.0;
            } else {
                r[i] = remaining;
            }
            pushed++;
            if (!visited[i]) {
                visited[i] = 1;
                fresh++;
            }
        }
        carry[i] = c;
    }
    if (pushed) atomicAdd(&s_cnt[0], pushed);
    if (fresh) atomicAdd(&s_cnt[1], fresh);
    __syncthreads();
    if (threadIdx.x < 2 && s_cnt[threadIdx.x]) atomicAdd(&counters[threadIdx.x], (unsigned long long)s_cnt[threadIdx.x]);
}

// candidates after dense rounds: every node that could be pushed
__global__ void __launch_bounds__(kThreads) candidate_flag_kernel(const double *__restrict__ r, const double *__restrict__ deg,
                                                                  uint64_t n, double eps, uint32_t *__restrict__ flag) {
    for (uint64_t i = blockIdx.x * (uint64_t)kThreads + threadIdx.x; i < n; i += (uint64_t)gridDim.x * kThreads)
        flag[i] = (r[i] > 0.0 && !(r[i] < eps * fmax(deg[i], 1.0))) ? 1u : 0u;
}
__global__ void __launch_bounds__(kThreads) candidate_fill_kernel(const uint32_t *__restrict__ flag,
                                                                  const uint32_t *__restrict__ slot, uint64_t n,
                                                                  uint32_t *__restrict__ Q) {
    for (uint64_t i = blockIdx.x * (uint64_t)kThreads + threadIdx.x; i < n; i += (uint64_t)gridDim.x * kThreads)
        if (flag[i]) Q[slot[i]] = (uint32_t)i;
}

// exclusive scan that keeps the input: dst = scan(src) (dst has count + 1 slots)
int32_t scan_copy(const uint32_t *src, uint32_t *dst, uint64_t count, uint64_t *total, cudaStream_t st) {
    SB_CUDA(cudaMemcpyAsync(dst, src, count * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
    return device_exclusive_scan_u32(dst, count, total, st);
}

struct RoundBuffers {
    DevBuf<uint32_t> Q[2], flag, rank, cnt, off, head, slot, key[2];
    DevBuf<double> val[2];
    DevBuf<char> sort_tmp;
    uint64_t qcap = 0, pcap = 0;
    int32_t ensure_q(uint64_t nq) {
        if (nq <= qcap) return SB200_OK;
        const uint64_t c = std::max<uint64_t>(nq + nq / 2, 1024);
        SB_TRY(flag.alloc(c + 1));
        SB_TRY(rank.alloc(c + 1));
        SB_TRY(cnt.alloc(c + 1));
        SB_TRY(off.alloc(c + 1));
        qcap = c;
        return SB200_OK;
    }
    int32_t ensure_p(uint64_t np, cudaStream_t st) {
        if (np <= pcap) return SB200_OK;
        const uint64_t c = std::max<uint64_t>(np + np / 2, 4096);
        if (c >= 0x7FFFFFF0ull) return fail(SB200_ERR_MEMORY_ALLOCATION, "push round with %llu contributions", (unsigned long long)np);
        // Q[1 - cur] receives up to `np` distinct neighbours; Q[cur] is being read: grow both (contents of the current list
        // are preserved by the caller re-filling it only between rounds)
        for (int i = 0; i < 2; i++) {
            SB_TRY(key[i].alloc(c));
            SB_TRY(val[i].alloc(c));
        }
        SB_TRY(head.alloc(c + 1));
        SB_TRY(slot.alloc(c + 1));
        size_t bytes = 0;
        SB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const uint32_t *)nullptr, (uint32_t *)nullptr,
                                                (const double *)nullptr, (double *)nullptr, (int)c, 0, 32, st));
        SB_TRY(sort_tmp.alloc(bytes));
        pcap = c;
        return SB200_OK;
    }
};

int key_bits(uint64_t n) {
    int b = 1;
    while (b < 32 && (1ull << b) < n) b++;
    return b;
}

int32_t ensure_dense(sb200_push_graph *g, bool backward) {
    std::lock_guard<std::mutex> lk(g->mu);
    const uint64_t n = g->n, nnz = g->nnz;
    if (!backward && !g->fwd) {
        std::vector<double> tval(nnz);
        for (uint64_t v = 0; v < n; v++)
            for (uint64_t k = g->h_tptr[v]; k < g->h_tptr[v + 1]; k++) {
                const uint32_t u = g->h_tcol[k];
                tval[k] = g->h_deg[u] > 0.0 ? g->h_tw[k] / g->h_deg[u] : 0.0;
            }
        SB_TRY(matrix_from_host_csr(g->h_tptr.data(), nullptr, g->h_tcol.data(), tval.data(), n, n, nnz, false, &g->fwd));
    }
    if (backward && !g->bwd) {
        std::vector<double> bval(nnz);
        for (uint64_t p = 0; p < n; p++)
            for (uint64_t k = g->h_ptr[p]; k < g->h_ptr[p + 1]; k++) bval[k] = g->h_w[k] / std::fmax(g->h_deg[p], 1.0);
        SB_TRY(matrix_from_host_csr(g->h_ptr.data(), nullptr, g->h_col.data(), bval.data(), n, n, nnz, false, &g->bwd));
    }
    return SB200_OK;
}

struct PushQuery {
    bool backward = false;
    bool watch = false;          // solve_with_target / solve_with_source
    uint64_t watch_node = 0;
    double watch_precision = 0.0;
};

int32_t push_run(const sb200_push_graph *gc, const sb200_push_config *cfg, const uint64_t *seeds, uint64_t nseeds,
                 const PushQuery &query, double *est_out, double *res_out, sb200_push_stats *stats) {
    clear_error();
    if (!gc || !cfg || !stats || (nseeds && !seeds)) return fail(SB200_ERR_INVALID_INPUT, "null argument");
    memset(stats, 0, sizeof(*stats));
    sb200_push_graph *g = const_cast<sb200_push_graph *>(gc);
    const uint64_t n = g->n;
    if (n && (!est_out || !res_out)) return fail(SB200_ERR_INVALID_INPUT, "null output buffer");
    if (n >= 0xFFFFFFF0ull) return fail(SB200_ERR_INVALID_INPUT, "graph too large for 32-bit node ids");
    DeviceGuard guard(g->device);
    SB_TRY(require_device(g->device));
    const bool backward = query.backward;
    cudaStream_t st = g->stream;
    PushView view{backward ? g->d_tptr.p : g->d_ptr.p, backward ? g->d_tcol.p : g->d_col.p, backward ? g->d_tw.p : g->d_w.p,
                  backward ? g->d_rdeg.p : g->d_deg.p, g->d_deg.p, backward ? 1 : 0};

    // initial residual: unit mass at the seed, or 1/len per listed seed (forward_push.rs:66-84, 125-136)
    std::vector<uint32_t> q0;
    std::vector<double> r0v;
    bool any = false;
    if (nseeds == 1) {
        if (seeds[0] < n) { q0.push_back((uint32_t)seeds[0]); any = true; }
    } else if (nseeds > 1) {
        for (uint64_t s = 0; s < nseeds; s++)
            if (seeds[s] < n) { q0.push_back((uint32_t)seeds[s]); any = true; }
    }
    if (!any || n == 0) {
        for (uint64_t i = 0; i < n; i++) est_out[i] = res_out[i] = 0.0;
        return SB200_OK;
    }
    const double mass = nseeds == 1 ? 1.0 : 1.0 / (double)nseeds;
    std::vector<double> r_host(n, 0.0);
    for (uint32_t v : q0) r_host[v] += mass;  // a seed listed twice receives two shares
    std::sort(q0.begin(), q0.end());
    q0.erase(std::unique(q0.begin(), q0.end()), q0.end());

    DevBuf<double> d_r, d_est, d_carry;
    DevBuf<unsigned char> d_vis;
    DevBuf<unsigned long long> d_cnt;
    RoundBuffers B;
    SB_TRY(d_r.alloc(n));
    SB_TRY(d_est.alloc(n));
    SB_TRY(d_vis.alloc(n));
    SB_TRY(d_cnt.alloc(4));
    SB_TRY(copy_h2d(d_r.p, r_host.data(), n * 8, st));
    SB_CUDA(cudaMemsetAsync(d_est.p, 0, n * 8, st));
    SB_CUDA(cudaMemsetAsync(d_vis.p, 0, n, st));
    SB_CUDA(cudaMemsetAsync(d_cnt.p, 0, 32, st));
    uint64_t nq = q0.size();
    SB_TRY(B.Q[0].alloc(std::max<uint64_t>(nq, 1024)));
    SB_TRY(B.Q[1].alloc(1024));
    SB_TRY(copy_h2d(B.Q[0].p, q0.data(), nq * sizeof(uint32_t), st));
    int cur = 0;

    cudaEvent_t e0, e1;
    SB_CUDA(cudaEventCreate(&e0));
    SB_CUDA(cudaEventCreate(&e1));
    struct EvGuard {
        cudaEvent_t a, b;
        ~EvGuard() { cudaEventDestroy(a); cudaEventDestroy(b); }
    } evg{e0, e1};
    SB_CUDA(cudaEventRecord(e0, st));

    uint64_t push_count = 0, rounds = 0, dense_rounds = 0, launches = 0, edges_touched = 0;
    double thr = cfg->queue_threshold;
    const int bits = key_bits(n);
    const double alpha = cfg->alpha, eps = cfg->epsilon;
    bool dense = false;
    auto adapt_threshold = [&](uint64_t before, uint64_t after, uint64_t queue_len) {
        if (!cfg->adaptive_threshold) return;
        for (uint64_t k = before / 1000 + 1; k <= after / 1000; k++) {  // every 1 000 pushes (forward_push.rs:107-109)
            if (queue_len > 10000) thr *= 1.1;
            else if (queue_len < 100 && thr > 1e-12) thr *= 0.9;
        }
    };

    while (push_count < cfg->max_pushes) {  // `while !work_queue.is_empty() && push_count < max_pushes` (:93)
        if (query.watch) {  // solve_with_target (:260-263) / solve_with_source (backward_push.rs:263-266), once per round
            double w[2];
            SB_CUDA(cudaMemcpyAsync(&w[0], d_est.p + query.watch_node, 8, cudaMemcpyDeviceToHost, st));
            SB_CUDA(cudaMemcpyAsync(&w[1], d_r.p + query.watch_node, 8, cudaMemcpyDeviceToHost, st));
            SB_CUDA(cudaStreamSynchronize(st));
            if (w[0] > query.watch_precision && w[1] < query.watch_precision * 0.1) break;
        }
        const uint64_t budget64 = cfg->max_pushes - push_count;
        const uint32_t budget = (uint32_t)std::min<uint64_t>(budget64, 0xFFFFFFFFull);
        if (!dense) {
            if (nq == 0) break;  // empty queue
            SB_TRY(B.ensure_q(nq));
            push_flag_kernel<<<grid_for(nq), kThreads, 0, st>>>(B.Q[cur].p, (uint32_t)nq, d_r.p, view, eps, thr, B.flag.p, B.cnt.p);
            SB_CUDA(cudaGetLastError());
            uint64_t npush = 0, npairs = 0;
            SB_TRY(scan_copy(B.flag.p, B.rank.p, nq, &npush, st));
            launches += 5;
            if (npush == 0) break;  // nothing above its threshold: the reference's queue drains without a push
            if (npush > budget) {
                // max_pushes cuts inside this round: only ranks below the budget push; zero the counts of the others
                // (ranks are a prefix in Q order, so this is a plain host-side trim of the tail)
                std::vector<uint32_t> hf(nq), hc(nq);
                SB_CUDA(cudaMemcpyAsync(hf.data(), B.flag.p, nq * 4, cudaMemcpyDeviceToHost, st));
                SB_CUDA(cudaMemcpyAsync(hc.data(), B.cnt.p, nq * 4, cudaMemcpyDeviceToHost, st));
                SB_CUDA(cudaStreamSynchronize(st));
                uint64_t seen = 0;
                for (uint64_t q = 0; q < nq; q++) {
                    if (hf[q]) {
                        if (seen >= budget) hc[q] = 0;
                        seen++;
                    }
                }
                SB_TRY(copy_h2d(B.cnt.p, hc.data(), nq * 4, st));
                npush = budget;
            }
            SB_TRY(scan_copy(B.cnt.p, B.off.p, nq, &npairs, st));
            if (npairs > g->nnz / 4 && npairs > (1u << 20) && budget64 >= n) {
                // the frontier covers a large part of the graph: dense rounds from here (the same pushes, one SpMV each)
                dense = true;
                continue;
            }
            SB_TRY(B.ensure_p(npairs, st));
            if (B.Q[1 - cur].n < npairs) SB_TRY(B.Q[1 - cur].alloc(npairs + npairs / 2));
            push_apply_kernel<<<grid_for(nq, kThreads / 32), kThreads, 0, st>>>(B.Q[cur].p, (uint32_t)nq, B.flag.p, B.rank.p,
                                                                                B.off.p, budget, d_r.p, d_est.p, d_vis.p, view,
                                                                                alpha, B.key[0].p, B.val[0].p, d_cnt.p + 1);
            SB_CUDA(cudaGetLastError());
            size_t tmp_bytes = B.sort_tmp.n;
            SB_CUDA(cub::DeviceRadixSort::SortPairs(B.sort_tmp.p, tmp_bytes, B.key[0].p, B.key[1].p, B.val[0].p, B.val[1].p,
                                                    (int)npairs, 0, bits, st));
            head_flag_kernel<<<grid_for(npairs), kThreads, 0, st>>>(B.key[1].p, (uint32_t)npairs, B.head.p);
            SB_CUDA(cudaGetLastError());
            uint64_t nuniq = 0;
            SB_TRY(scan_copy(B.head.p, B.slot.p, npairs, &nuniq, st));
            reduce_apply_kernel<<<grid_for(npairs), kThreads, 0, st>>>(B.key[1].p, B.val[1].p, B.head.p, B.slot.p,
                                                                       (uint32_t)npairs, d_r.p, B.Q[1 - cur].p);
            SB_CUDA(cudaGetLastError());
            launches += 8;
            adapt_threshold(push_count, push_count + npush, nq);
            push_count += npush;
            edges_touched += npairs;
            rounds++;
            cur = 1 - cur;
            nq = nuniq;
        } else {
            SB_TRY(ensure_dense(g, backward));
            const sb200_matrix *M = backward ? g->bwd : g->fwd;
            if (!d_carry.p) SB_TRY(d_carry.alloc(n));
            if (budget64 < n) {
                // a dense round pushes every node above its threshold at once and could overshoot max_pushes: the last
                // pushes go back to the sparse form, which cuts by rank
                SB_TRY(B.flag.alloc(n + 1));
                SB_TRY(B.rank.alloc(n + 1));
                B.qcap = 0;
                candidate_flag_kernel<<<grid_for(n), kThreads, 0, st>>>(d_r.p, view.tdeg, n, eps, B.flag.p);
                uint64_t nc = 0;
                SB_TRY(scan_copy(B.flag.p, B.rank.p, n, &nc, st));
                if (B.Q[cur].n < nc) SB_TRY(B.Q[cur].alloc(nc + 1024));
                candidate_fill_kernel<<<grid_for(n), kThreads, 0, st>>>(B.flag.p, B.rank.p, n, B.Q[cur].p);
                SB_CUDA(cudaGetLastError());
                SB_TRY(B.flag.alloc(1));
                SB_TRY(B.rank.alloc(1));
                nq = nc;
                dense = false;
                launches += 5;
                continue;
            }
            SB_CUDA(cudaMemsetAsync(d_cnt.p, 0, 8, st));
            push_select_kernel<<<grid_for(n), kThreads, 0, st>>>(d_r.p, d_est.p, d_carry.p, view.tdeg, d_vis.p, n, alpha, eps, thr,
                                                                 d_cnt.p);
            SB_CUDA(cudaGetLastError());
            unsigned long long h[2] = {0, 0};
            SB_CUDA(cudaMemcpyAsync(h, d_cnt.p, 16, cudaMemcpyDeviceToHost, st));
            SB_CUDA(cudaStreamSynchronize(st));
            launches++;
            if (h[0] == 0) break;
            SB_TRY(matrix_spmv_dev(M, d_carry.p, d_r.p, 1, st));  // r += M carry
            launches += launches_per_pass(M);
            adapt_threshold(push_count, push_count + h[0], h[0]);
            push_count += h[0];
            edges_touched += g->nnz;
            rounds++;
            dense_rounds++;
            if (h[0] * 32 < n) {  // the frontier shrank: back to candidate lists
                SB_TRY(B.flag.alloc(n + 1));
                SB_TRY(B.rank.alloc(n + 1));
                B.qcap = 0;
                candidate_flag_kernel<<<grid_for(n), kThreads, 0, st>>>(d_r.p, view.tdeg, n, eps, B.flag.p);
                uint64_t nc = 0;
                SB_TRY(scan_copy(B.flag.p, B.rank.p, n, &nc, st));
                if (B.Q[cur].n < nc) SB_TRY(B.Q[cur].alloc(nc + 1024));
                candidate_fill_kernel<<<grid_for(n), kThreads, 0, st>>>(B.flag.p, B.rank.p, n, B.Q[cur].p);
                SB_CUDA(cudaGetLastError());
                SB_TRY(B.flag.alloc(1));
                SB_TRY(B.rank.alloc(1));
                nq = nc;
                dense = false;
                launches += 5;
            }
        }
    }
    SB_CUDA(cudaEventRecord(e1, st));
    unsigned long long hv[4] = {0, 0, 0, 0};
    SB_CUDA(cudaMemcpyAsync(hv, d_cnt.p, 32, cudaMemcpyDeviceToHost, st));
    SB_TRY(copy_d2h(est_out, d_est.p, n * 8, st));
    SB_TRY(copy_d2h(res_out, d_r.p, n * 8, st));
    SB_CUDA(cudaStreamSynchronize(st));
    float ms = 0.f;
    SB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    double nrm = 0.0;  // compute_residual_norm (:217-219): sequential sum of squares
    for (uint64_t i = 0; i < n; i++) nrm += res_out[i] * res_out[i];
    uint64_t visited = 0;
    {   // the dense rounds count first visits in counters[1] too; recount from the flags to cover both forms
        std::vector<unsigned char> vis(n);
        SB_CUDA(cudaMemcpy(vis.data(), d_vis.p, n, cudaMemcpyDeviceToHost));
        for (uint64_t i = 0; i < n; i++) visited += vis[i];
    }
    stats->push_count = push_count;
    stats->nodes_visited = visited;
    stats->residual_norm = std::sqrt(nrm);
    stats->rounds = rounds;
    stats->kernel_launches = launches;
    stats->device_time_ms = ms;
    stats->dense_rounds = dense_rounds;
    stats->edges_touched = edges_touched;
    return SB200_OK;
}

// =====================================================================================================================
// SublinearSolver.solveForwardPush for A x = b (src/core/solver.ts:437-522): residual r = b - A x is kept exact; a push
// of node i sets x_i += r_i / a_ii, r_i = 0 and r_j -= a_ji * (r_i / a_ii) down column i. The reference pushes the single
// node of largest |r| per iteration (Gauss-Southwell, an O(n) scan each); here every candidate with |r_i| >= epsilon is
// pushed in the same round (the parallel-Southwell relaxation converges for the diagonally dominant systems of this
// path), with the same sort + ordered reduce as the graph push. Stops when max |r| < epsilon over the candidates =
// over all nodes; `iterations` counts node pushes like the reference's loop counter.
// =====================================================================================================================
__global__ void __launch_bounds__(kThreads) axb_flag_kernel(const uint32_t *__restrict__ Q, uint32_t nq,
                                                            const double *__restrict__ r, const uint32_t *__restrict__ tptr,
                                                            double eps, uint32_t *__restrict__ flag, uint32_t *__restrict__ cnt) {
    for (uint32_t q = blockIdx.x * kThreads + threadIdx.x; q < nq; q += gridDim.x * kThreads) {
        const uint32_t v = Q[q];
        const bool push = !(fabs(r[v]) < eps);  // `if (maxResidual < epsilon) converged` (:463)
        flag[q] = push ? 1u : 0u;
        cnt[q] = push ? tptr[v + 1] - tptr[v] : 0u;
    }
}

__global__ void __launch_bounds__(kThreads) axb_apply_kernel(const uint32_t *__restrict__ Q, uint32_t nq,
                                                             const uint32_t *__restrict__ flag, const uint32_t *__restrict__ rank,
                                                             const uint32_t *__restrict__ off, uint32_t budget,
                                                             double *__restrict__ r, double *__restrict__ x,
                                                             const double *__restrict__ diag, const uint32_t *__restrict__ tptr,
                                                             const uint32_t *__restrict__ trow, const double *__restrict__ tval,
                                                             uint32_t *__restrict__ pkey, double *__restrict__ pval) {
    const int lane = threadIdx.x & 31;
    const uint32_t nwarps = gridDim.x * (kThreads / 32);
    for (uint32_t q = blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5); q < nq; q += nwarps) {
        if (!flag[q] || rank[q] >= budget) continue;
        const uint32_t i = Q[q];
        const double push = r[i] / diag[i];  // pushValue = residual[maxNode] / diagEntry (:474)
        const uint32_t base = off[q], s = tptr[i], e = tptr[i + 1];
        __syncwarp();
        if (lane == 0) {
            x[i] = x[i] + push;              // :475
            r[i] = 0.0;                      // :476
        }
        for (uint32_t k = s + lane; k < e; k += 32) {
            const uint32_t j = trow[k];
            pkey[base + (k - s)] = j;
            // residual[j] -= entry * pushValue for j != i (:479-484); the diagonal entry contributes nothing (+0.0 keeps i
            // in the candidate list, which is harmless: its residual is 0)
            pval[base + (k - s)] = j == i ? 0.0 : -(tval[k] * push);
        }
    }
}

}  // namespace

// transposed copy (columns of A) + diagonal, cached on the matrix handle
struct AxbCache {
    DevBuf<uint32_t> tptr, trow;
    DevBuf<double> tval, diag;
    std::vector<double> h_diag;
};

namespace sb200 {
void axb_cache_free(void *p) { delete static_cast<AxbCache *>(p); }
}  // namespace sb200

extern "C" {

// ForwardPushConfig::default = BackwardPushConfig::default (forward_push.rs:40-50, backward_push.rs:40-50)
void sb200_push_config_default(sb200_push_config *c) {
    if (!c) return;
    memset(c, 0, sizeof(*c));
    c->alpha = 0.15;
    c->epsilon = 1e-6;
    c->max_pushes = 1000000;
    c->queue_threshold = 1e-8;
    c->adaptive_threshold = 1;
}

// PushGraph::from_matrix (adjacency.rs:211-224): adjacency + transpose + row sums + column sums
int32_t sb200_push_graph_from_csr(const uint64_t *row_ptr, const uint32_t *col_indices, const double *weights, uint64_t n,
                                  sb200_push_graph **out) {
    clear_error();
    if (!out || !row_ptr) return fail(SB200_ERR_INVALID_INPUT, "null argument");
    *out = nullptr;
    const uint64_t nnz = row_ptr[n];
    if (nnz && (!col_indices || !weights)) return fail(SB200_ERR_INVALID_INPUT, "null CSR slice");
    if (n >= 0xFFFFFFF0ull || nnz >= 0xFFFFFFF0ull) return fail(SB200_ERR_INVALID_INPUT, "graph exceeds the u32 index type");
    for (uint64_t u = 0; u < n; u++)
        if (row_ptr[u] > row_ptr[u + 1]) return fail(SB200_ERR_INVALID_SPARSE_MATRIX, "row_ptr decreases at row %llu", (unsigned long long)u);
    for (uint64_t k = 0; k < nnz; k++) {
        if (col_indices[k] >= n)
            return fail(SB200_ERR_INDEX_OUT_OF_BOUNDS, "column index %u out of bounds (max %llu) at entry %llu", col_indices[k],
                        (unsigned long long)(n ? n - 1 : 0), (unsigned long long)k);
        if (!std::isfinite(weights[k])) return fail(SB200_ERR_INVALID_INPUT, "non-finite weight at entry %llu", (unsigned long long)k);
    }
    std::unique_ptr<sb200_push_graph> g(new sb200_push_graph());
    g->device = current_device();
    g->n = n;
    g->nnz = nnz;
    g->h_deg.assign(n, 0.0);
    g->h_rdeg.assign(n, 0.0);
    g->h_ptr.assign(row_ptr, row_ptr + n + 1);
    g->h_col.assign(col_indices, col_indices + nnz);
    g->h_w.assign(weights, weights + nnz);
    for (uint64_t u = 0; u < n; u++)  // row_sums (mod.rs:80-88) in row order; column sums in the transpose's row order
        for (uint64_t k = row_ptr[u]; k < row_ptr[u + 1]; k++) g->h_deg[u] += weights[k];
    // transpose by counting sort (mod.rs:93-127): within a column the entries keep ascending source order
    g->h_tptr.assign(n + 1, 0);
    for (uint64_t k = 0; k < nnz; k++) g->h_tptr[col_indices[k] + 1]++;
    for (uint64_t i = 0; i < n; i++) g->h_tptr[i + 1] += g->h_tptr[i];
    g->h_tcol.resize(nnz);
    g->h_tw.resize(nnz);
    {
        std::vector<uint64_t> pos(g->h_tptr.begin(), g->h_tptr.end() - 1);
        for (uint64_t u = 0; u < n; u++)
            for (uint64_t k = row_ptr[u]; k < row_ptr[u + 1]; k++) {
                const uint64_t p = pos[col_indices[k]]++;
                g->h_tcol[p] = (uint32_t)u;
                g->h_tw[p] = weights[k];
            }
    }
    for (uint64_t v = 0; v < n; v++)
        for (uint64_t k = g->h_tptr[v]; k < g->h_tptr[v + 1]; k++) g->h_rdeg[v] += g->h_tw[k];
    SB_TRY(require_device(g->device));
    DeviceGuard guard(g->device);
    SB_CUDA(cudaStreamCreateWithFlags(&g->stream, cudaStreamNonBlocking));
    std::vector<uint32_t> p32(n + 1), t32(n + 1);
    for (uint64_t i = 0; i <= n; i++) { p32[i] = (uint32_t)row_ptr[i]; t32[i] = (uint32_t)g->h_tptr[i]; }
    SB_TRY(g->d_ptr.alloc(n + 1));
    SB_TRY(g->d_tptr.alloc(n + 1));
    SB_TRY(g->d_col.alloc(nnz));
    SB_TRY(g->d_tcol.alloc(nnz));
    SB_TRY(g->d_w.alloc(nnz));
    SB_TRY(g->d_tw.alloc(nnz));
    SB_TRY(g->d_deg.alloc(n));
    SB_TRY(g->d_rdeg.alloc(n));
    SB_TRY(copy_h2d(g->d_ptr.p, p32.data(), (n + 1) * 4, g->stream));
    SB_TRY(copy_h2d(g->d_tptr.p, t32.data(), (n + 1) * 4, g->stream));
    SB_TRY(copy_h2d(g->d_col.p, col_indices, nnz * 4, g->stream));
    SB_TRY(copy_h2d(g->d_tcol.p, g->h_tcol.data(), nnz * 4, g->stream));
    SB_TRY(copy_h2d(g->d_w.p, weights, nnz * 8, g->stream));
    SB_TRY(copy_h2d(g->d_tw.p, g->h_tw.data(), nnz * 8, g->stream));
    SB_TRY(copy_h2d(g->d_deg.p, g->h_deg.data(), n * 8, g->stream));
    SB_TRY(copy_h2d(g->d_rdeg.p, g->h_rdeg.data(), n * 8, g->stream));
    SB_CUDA(cudaStreamSynchronize(g->stream));
    *out = g.release();
    return SB200_OK;
}

// PushGraph::from_edges (adjacency.rs:227-239): out-of-range edges are dropped, parallel edges kept, insertion order
int32_t sb200_push_graph_from_edges(uint64_t n, const uint64_t *from, const uint64_t *to, const double *weights,
                                    uint64_t nedges, sb200_push_graph **out) {
    clear_error();
    if (!out || (nedges && (!from || !to || !weights))) return fail(SB200_ERR_INVALID_INPUT, "null argument");
    std::vector<uint64_t> rp(n + 1, 0);
    for (uint64_t e = 0; e < nedges; e++)
        if (from[e] < n && to[e] < n) rp[from[e] + 1]++;
    for (uint64_t i = 0; i < n; i++) rp[i + 1] += rp[i];
    std::vector<uint32_t> ci(rp[n]);
    std::vector<double> w(rp[n]);
    std::vector<uint64_t> pos(rp.begin(), rp.end() - 1);
    for (uint64_t e = 0; e < nedges; e++)
        if (from[e] < n && to[e] < n) {
            const uint64_t p = pos[from[e]]++;
            ci[p] = (uint32_t)to[e];
            w[p] = weights[e];
        }
    return sb200_push_graph_from_csr(rp.data(), ci.data(), w.data(), n, out);
}

int32_t sb200_push_graph_info(const sb200_push_graph *g, uint64_t *num_nodes, uint64_t *num_edges) {
    if (!g) return fail(SB200_ERR_INVALID_INPUT, "null argument");
    if (num_nodes) *num_nodes = g->n;
    if (num_edges) *num_edges = g->nnz;
    return SB200_OK;
}

// PushGraph::out_degree / in_degree (adjacency.rs:262-276): 0 for out-of-range nodes
int32_t sb200_push_graph_degrees(const sb200_push_graph *g, uint64_t node, double *out_degree, double *in_degree) {
    if (!g) return fail(SB200_ERR_INVALID_INPUT, "null argument");
    if (out_degree) *out_degree = node < g->n ? g->h_deg[node] : 0.0;
    if (in_degree) *in_degree = node < g->n ? g->h_rdeg[node] : 0.0;
    return SB200_OK;
}

void sb200_push_graph_free(sb200_push_graph *g) {
    if (!g) return;
    DeviceGuard guard(g->device);
    delete g;
}

int32_t sb200_forward_push(const sb200_push_graph *g, const sb200_push_config *cfg, const uint64_t *sources,
                           uint64_t nsources, double *estimate, double *residual, sb200_push_stats *stats) {
    return push_run(g, cfg, sources, nsources, PushQuery{}, estimate, residual, stats);
}

int32_t sb200_backward_push(const sb200_push_graph *g, const sb200_push_config *cfg, const uint64_t *targets,
                            uint64_t ntargets, double *estimate, double *residual, sb200_push_stats *stats) {
    PushQuery q;
    q.backward = true;
    return push_run(g, cfg, targets, ntargets, q, estimate, residual, stats);
}

// ForwardPushSolver::solve_with_target (forward_push.rs:234-290): out-of-range source or target -> all-zero result
int32_t sb200_forward_push_with_target(const sb200_push_graph *g, const sb200_push_config *cfg, uint64_t source,
                                       uint64_t target, double target_precision, double *estimate, double *residual,
                                       sb200_push_stats *stats) {
    if (!g || !stats) return fail(SB200_ERR_INVALID_INPUT, "null argument");
    PushQuery q;
    q.watch = true;
    q.watch_node = target;
    q.watch_precision = target_precision;
    if (source >= g->n || target >= g->n) source = g->n;  // push_run answers an out-of-range seed with zeros
    if (source >= g->n) q.watch = false;
    return push_run(g, cfg, &source, 1, q, estimate, residual, stats);
}

// BackwardPushSolver::solve_with_source (backward_push.rs:238-290)
int32_t sb200_backward_push_with_source(const sb200_push_graph *g, const sb200_push_config *cfg, uint64_t source,
                                        uint64_t target, double source_precision, double *estimate, double *residual,
                                        sb200_push_stats *stats) {
    if (!g || !stats) return fail(SB200_ERR_INVALID_INPUT, "null argument");
    PushQuery q;
    q.backward = true;
    q.watch = true;
    q.watch_node = source;
    q.watch_precision = source_precision;
    if (source >= g->n || target >= g->n) target = g->n;
    if (target >= g->n) q.watch = false;
    return push_run(g, cfg, &target, 1, q, estimate, residual, stats);
}

// BackwardPushSolver::combine_with_forward (backward_push.rs:312-330): sequential accumulation in index order
int32_t sb200_push_combine_with_forward(double alpha, const double *backward_estimate, const double *backward_residual,
                                        uint64_t nbackward, const double *forward_estimate, const double *forward_residual,
                                        uint64_t nforward, double *out) {
    clear_error();
    if (!out || !backward_estimate || !backward_residual || !forward_estimate || !forward_residual)
        return fail(SB200_ERR_INVALID_INPUT, "null argument");
    double total = 0.0;
    const uint64_t m = std::min(nbackward, nforward);
    for (uint64_t i = 0; i < m; i++) {
        total += backward_estimate[i] * forward_estimate[i];
        total += backward_residual[i] * forward_estimate[i] * alpha;
        total += backward_estimate[i] * forward_residual[i] * alpha;
    }
    *out = total;
    return SB200_OK;
}

// BidirectionalPushSolver::solve_bidirectional (backward_push.rs:362-384)
int32_t sb200_bidirectional_push(const sb200_push_graph *g, const sb200_push_config *forward_cfg,
                                 const sb200_push_config *backward_cfg, uint64_t source, uint64_t target, double *out) {
    clear_error();
    if (!g || !forward_cfg || !backward_cfg || !out) return fail(SB200_ERR_INVALID_INPUT, "null argument");
    const uint64_t n = g->n;
    std::vector<double> fe(n), fr(n), be(n), br(n);
    sb200_push_stats st;
    SB_TRY(sb200_forward_push(g, forward_cfg, &source, 1, fe.data(), fr.data(), &st));
    SB_TRY(sb200_backward_push(g, backward_cfg, &target, 1, be.data(), br.data(), &st));
    return sb200_push_combine_with_forward(backward_cfg->alpha, be.data(), br.data(), n, fe.data(), fr.data(), n, out);
}

// BidirectionalPushSolver::adaptive_solve (backward_push.rs:387-420)
int32_t sb200_bidirectional_adaptive_push(const sb200_push_graph *g, const sb200_push_config *forward_cfg,
                                          const sb200_push_config *backward_cfg, uint64_t source, uint64_t target, double *out) {
    clear_error();
    if (!g || !forward_cfg || !backward_cfg || !out) return fail(SB200_ERR_INVALID_INPUT, "null argument");
    const uint64_t n = g->n;
    *out = 0.0;
    if (source >= n || target >= n) return SB200_OK;
    const double so = g->h_deg[source], ti = g->h_rdeg[target];
    std::vector<double> e(n), r(n);
    sb200_push_stats st;
    if (so > ti * 2.0) {         // backward push from the target, read the source's entry
        SB_TRY(sb200_backward_push(g, backward_cfg, &target, 1, e.data(), r.data(), &st));
        *out = e[source];
        return SB200_OK;
    }
    if (ti > so * 2.0) {         // forward push from the source, read the target's entry
        SB_TRY(sb200_forward_push(g, forward_cfg, &source, 1, e.data(), r.data(), &st));
        *out = e[target];
        return SB200_OK;
    }
    return sb200_bidirectional_push(g, forward_cfg, backward_cfg, source, target, out);
}

// SublinearSolver.solveForwardPush (src/core/solver.ts:437-522) for A x = b
int32_t sb200_forward_push_solve(const sb200_matrix *mc, const double *b, uint64_t blen, double epsilon,
                                 uint64_t max_iterations, double *x_out, sb200_axb_push_stats *stats) {
    clear_error();
    if (!mc || !stats || (blen && (!b || !x_out))) return fail(SB200_ERR_INVALID_INPUT, "null argument");
    memset(stats, 0, sizeof(*stats));
    sb200_matrix *m = const_cast<sb200_matrix *>(mc);
    if (m->distributed) return fail(SB200_ERR_INVALID_INPUT, "forward push needs whole columns: not available on a row block");
    if (m->nrows != m->ncols) return fail(SB200_ERR_INVALID_INPUT, "matrix must be square");
    const uint64_t n = m->nrows;
    if (blen != n) return fail(SB200_ERR_DIMENSION_MISMATCH, "expected %llu, actual %llu in forward_push", (unsigned long long)n, (unsigned long long)blen);
    DeviceGuard guard(m->device);
    SB_TRY(require_device(m->device));
    cudaStream_t st = m->stream;
    // columns of A (transpose by counting sort on the host, once per matrix) + diagonal (duplicates summed)
    AxbCache *c;
    {
        std::lock_guard<std::mutex> lk(m->mu);
        if (!m->axb_cache) {
            std::unique_ptr<AxbCache> nc(new AxbCache());
            const uint64_t nnz = m->nnz;
            std::vector<uint32_t> ci(nnz), trow(nnz), tptr(n + 1, 0);
            std::vector<double> cv(nnz), tval(nnz);
            SB_CUDA(cudaMemcpy(ci.data(), m->d_cols.p, nnz * 4, cudaMemcpyDeviceToHost));
            SB_CUDA(cudaMemcpy(cv.data(), m->d_vals.p, nnz * 8, cudaMemcpyDeviceToHost));
            nc->h_diag.assign(n, 0.0);
            const uint32_t *rp = m->h_row_ptr.data();
            for (uint64_t k = 0; k < nnz; k++) tptr[ci[k] + 1]++;
            for (uint64_t i = 0; i < n; i++) tptr[i + 1] += tptr[i];
            std::vector<uint32_t> pos(tptr.begin(), tptr.end() - 1);
            for (uint64_t i = 0; i < n; i++)
                for (uint32_t k = rp[i]; k < rp[i + 1]; k++) {
                    const uint32_t p = pos[ci[k]]++;
                    trow[p] = (uint32_t)i;
                    tval[p] = cv[k];
                    if (ci[k] == i) nc->h_diag[i] += cv[k];
                }
            SB_TRY(nc->tptr.alloc(n + 1));
            SB_TRY(nc->trow.alloc(nnz));
            SB_TRY(nc->tval.alloc(nnz));
            SB_TRY(nc->diag.alloc(n));
            SB_TRY(copy_h2d(nc->tptr.p, tptr.data(), (n + 1) * 4, st));
            SB_TRY(copy_h2d(nc->trow.p, trow.data(), nnz * 4, st));
            SB_TRY(copy_h2d(nc->tval.p, tval.data(), nnz * 8, st));
            SB_TRY(copy_h2d(nc->diag.p, nc->h_diag.data(), n * 8, st));
            SB_CUDA(cudaStreamSynchronize(st));
            m->axb_cache = nc.release();
            m->axb_cache_free = axb_cache_free;
        }
        c = static_cast<AxbCache *>(m->axb_cache);
    }
    DevBuf<double> d_r, d_x;
    RoundBuffers B;
    SB_TRY(d_r.alloc(n));
    SB_TRY(d_x.alloc(n));
    SB_TRY(copy_h2d(d_r.p, b, n * 8, st));           // residual = [...vector], approximate = 0 (:439-440)
    SB_CUDA(cudaMemsetAsync(d_x.p, 0, n * 8, st));
    // first candidates: every node whose right-hand side is not below epsilon
    std::vector<uint32_t> q0;
    for (uint64_t i = 0; i < n; i++)
        if (!(std::fabs(b[i]) < epsilon)) q0.push_back((uint32_t)i);
    uint64_t nq = q0.size();
    SB_TRY(B.Q[0].alloc(std::max<uint64_t>(nq, 1024)));
    SB_TRY(B.Q[1].alloc(1024));
    SB_TRY(copy_h2d(B.Q[0].p, q0.data(), nq * 4, st));
    int cur = 0;
    const int bits = key_bits(n);
    uint64_t pushes = 0, rounds = 0;
    bool converged = false;
    while (true) {
        if (nq == 0) { converged = true; break; }
        SB_TRY(B.ensure_q(nq));
        axb_flag_kernel<<<grid_for(nq), kThreads, 0, st>>>(B.Q[cur].p, (uint32_t)nq, d_r.p, c->tptr.p, epsilon, B.flag.p, B.cnt.p);
        SB_CUDA(cudaGetLastError());
        uint64_t npush = 0, npairs = 0;
        SB_TRY(scan_copy(B.flag.p, B.rank.p, nq, &npush, st));
        if (npush == 0) { converged = true; break; }
        if (pushes >= max_iterations) break;             // `for (iter < maxIterations)` exhausted (:452, :505-511)
        const uint64_t budget64 = max_iterations - pushes;
        const uint32_t budget = (uint32_t)std::min<uint64_t>(budget64, 0xFFFFFFFFull);
        if (npush > budget) {
            std::vector<uint32_t> hf(nq), hc(nq);
            SB_CUDA(cudaMemcpyAsync(hf.data(), B.flag.p, nq * 4, cudaMemcpyDeviceToHost, st));
            SB_CUDA(cudaMemcpyAsync(hc.data(), B.cnt.p, nq * 4, cudaMemcpyDeviceToHost, st));
            SB_CUDA(cudaStreamSynchronize(st));
            uint64_t seen = 0;
            for (uint64_t q = 0; q < nq; q++)
                if (hf[q]) {
                    if (seen >= budget) hc[q] = 0;
                    seen++;
                }
            SB_TRY(copy_h2d(B.cnt.p, hc.data(), nq * 4, st));
            npush = budget;
        }
        // zero-diagonal check for the nodes about to be pushed (`Zero diagonal at position i`, :468-471)
        SB_TRY(scan_copy(B.cnt.p, B.off.p, nq, &npairs, st));
        SB_TRY(B.ensure_p(npairs + 1, st));
        if (B.Q[1 - cur].n < npairs + nq) SB_TRY(B.Q[1 - cur].alloc(npairs + nq + 1024));
        axb_apply_kernel<<<grid_for(nq, kThreads / 32), kThreads, 0, st>>>(B.Q[cur].p, (uint32_t)nq, B.flag.p, B.rank.p, B.off.p,
                                                                           budget, d_r.p, d_x.p, c->diag.p, c->tptr.p, c->trow.p,
                                                                           c->tval.p, B.key[0].p, B.val[0].p);
        SB_CUDA(cudaGetLastError());
        pushes += npush;
        rounds++;
        if (npairs == 0) {  // pushed nodes without any column entry cannot happen (the diagonal is in the column)
            nq = 0;
            continue;
        }
        size_t tmp_bytes = B.sort_tmp.n;
        SB_CUDA(cub::DeviceRadixSort::SortPairs(B.sort_tmp.p, tmp_bytes, B.key[0].p, B.key[1].p, B.val[0].p, B.val[1].p,
                                                (int)npairs, 0, bits, st));
        head_flag_kernel<<<grid_for(npairs), kThreads, 0, st>>>(B.key[1].p, (uint32_t)npairs, B.head.p);
        SB_CUDA(cudaGetLastError());
        uint64_t nuniq = 0;
        SB_TRY(scan_copy(B.head.p, B.slot.p, npairs, &nuniq, st));
        reduce_apply_kernel<<<grid_for(npairs), kThreads, 0, st>>>(B.key[1].p, B.val[1].p, B.head.p, B.slot.p, (uint32_t)npairs,
                                                                   d_r.p, B.Q[1 - cur].p);
        SB_CUDA(cudaGetLastError());
        cur = 1 - cur;
        nq = nuniq;
    }
    std::vector<double> r(n);
    SB_TRY(copy_d2h(x_out, d_x.p, n * 8, st));
    SB_TRY(copy_d2h(r.data(), d_r.p, n * 8, st));
    SB_CUDA(cudaStreamSynchronize(st));
    double nrm = 0.0, mx = 0.0;
    bool finite = true;
    for (uint64_t i = 0; i < n; i++) {
        nrm += r[i] * r[i];
        mx = std::fmax(mx, std::fabs(r[i]));
        finite = finite && std::isfinite(r[i]) && std::isfinite(x_out[i]);
    }
    stats->iterations = pushes;
    stats->rounds = rounds;
    stats->residual_norm = std::sqrt(nrm);  // VectorOperations.norm2(residual) (:489)
    stats->max_residual = mx;
    stats->converged = converged && finite;
    if (!finite) {  // a zero (or ~0) diagonal under a pushed node: `Zero diagonal at position i` (:468-471)
        for (uint64_t i = 0; i < n; i++)
            if (std::fabs(c->h_diag[i]) < 1e-15 && !(std::fabs(b[i]) < epsilon))
                return fail(SB200_ERR_NUMERICAL_INSTABILITY, "Zero diagonal at position %llu", (unsigned long long)i);
        return fail(SB200_ERR_NUMERICAL_INSTABILITY, "non-finite residual in forward push");
    }
    if (!converged)
        return fail(SB200_ERR_CONVERGENCE_FAILURE, "Forward push failed to converge after %llu iterations (residual %.6e)",
                    (unsigned long long)max_iterations, stats->residual_norm);
    return SB200_OK;
}

}  // extern "C"
