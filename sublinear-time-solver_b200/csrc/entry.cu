// entry.cu — batched single-entry estimation (walk-per-thread Monte Carlo) and the PageRank system builder.
//
// The Rust crate has no solve_entry; the capability lives in the TS package
// (SublinearSolver.estimateEntry / performRandomWalk, ref src/core/solver.ts:550-659, 390-432), which builds a
// dense n x n transition table per call and uses an absorption rule that is only right for special matrices.
// This implements the unbiased absorbing-walk (Ulam-von Neumann) estimator specified in SURVEY.md Appendix C
// directly on the CSR rows, keeping the TS defaults: numSamples = max(100, ceil(1/eps^2)) (:587), 1000 steps (:399).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "matrix.hpp"

namespace sb200 {

__host__ __device__ __forceinline__ uint64_t splitmix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

constexpr int kWalkThreads = 256;
constexpr int kWalksPerThread = 4;

// grid = (chunks, nqueries). Thread handles walks chunk*1024 + k*256 + tid, k < 4.
// Walk w of query q draws u_d = (splitmix64(key + d) >> 11) * 2^-53 with
// key = splitmix64(splitmix64(seed ^ (q0+q+1)*0xA0761D6478BD642F) + w), q0 = position of this call's first query in the
// caller's batch: counter based, so the estimate does not depend on the launch geometry, on how a batch is cut into
// calls, or on the GPU count (sb200_solve_entry_replicas).
__global__ void __launch_bounds__(kWalkThreads) walk_kernel(const double *__restrict__ vals, const uint32_t *__restrict__ cols,
                                                            const uint32_t *__restrict__ row_ptr,
                                                            const double *__restrict__ dinv, const double *__restrict__ b,
                                                            const uint64_t *__restrict__ qrows, uint64_t nwalks,
                                                            uint32_t max_steps, uint64_t seed, uint64_t q0,
                                                            double *__restrict__ part) {
    __shared__ double s_a[kWalkThreads / 32], s_b[kWalkThreads / 32];
    const uint32_t q = blockIdx.y, chunk = blockIdx.x;
    const uint64_t qkey = splitmix64(seed ^ ((q0 + (uint64_t)q + 1ull) * 0xA0761D6478BD642Full));
    const uint32_t start = (uint32_t)qrows[q];
    double sum = 0.0, sumsq = 0.0;
    for (int k = 0; k < kWalksPerThread; k++) {
        const uint64_t w = (uint64_t)chunk * (kWalkThreads * kWalksPerThread) + (uint64_t)k * kWalkThreads + threadIdx.x;
        if (w >= nwalks) break;
        const uint64_t key = splitmix64(qkey + w);
        uint32_t s = start;
        double W = 1.0, acc = 0.0;
        for (uint32_t step = 0; step < max_steps; step++) {
            const double ds = dinv[s];
            acc += W * (b[s] * ds);  // every visited state pays W * c_s, c = D^-1 b
            const double u = (double)(splitmix64(key + step) >> 11) * (1.0 / 9007199254740992.0);
            double cum = 0.0;
            bool moved = false;
            const uint32_t re = row_ptr[s + 1];
            for (uint32_t p = row_ptr[s]; p < re; p++) {
                const uint32_t j = cols[p];
                if (j == s) continue;
                const double mv = -(vals[p] * ds);  // M_sj = -a_sj / a_ss
                cum += fabs(mv);
                if (cum > u) {  // move with probability |M_sj|, carry its sign
                    if (mv < 0.0) W = -W;
                    s = j;
                    moved = true;
                    break;
                }
            }
            if (!moved) break;  // absorbed with probability 1 - sum_j |M_sj|
        }
        sum += acc;
        sumsq += acc * acc;
    }
    // CTA partial (fixed tree) -> part[q][chunk]; a second kernel adds the chunks in order
    for (int o = 16; o > 0; o >>= 1) {
        sum += __shfl_xor_sync(0xffffffffu, sum, o);
        sumsq += __shfl_xor_sync(0xffffffffu, sumsq, o);
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { s_a[warp] = sum; s_b[warp] = sumsq; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, c = 0.0;
        for (int i = 0; i < kWalkThreads / 32; i++) { a += s_a[i]; c += s_b[i]; }
        part[((size_t)q * gridDim.x + chunk) * 2 + 0] = a;
        part[((size_t)q * gridDim.x + chunk) * 2 + 1] = c;
    }
}

__global__ void walk_finalize_kernel(const double *__restrict__ part, uint32_t nq, uint32_t chunks, uint64_t nwalks,
                                     double *__restrict__ est, double *__restrict__ var) {
    const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= nq) return;
    double a = 0.0, c = 0.0;
    for (uint32_t k = 0; k < chunks; k++) {
        a += part[((size_t)q * chunks + k) * 2 + 0];
        c += part[((size_t)q * chunks + k) * 2 + 1];
    }
    const double mean = a / (double)nwalks;
    est[q] = mean;
    var[q] = nwalks > 1 ? (c - (double)nwalks * mean * mean) / (double)(nwalks - 1) : 0.0;
}



}  // namespace sb200

using namespace sb200;

namespace {

// argument checks shared by the single-GPU call and the replica dispatcher (reference order: dimensions, epsilon, rows)
int32_t entry_precheck(const sb200_matrix *m, const double *b, uint64_t blen, const uint64_t *rows, uint64_t nqueries,
                       double eps, uint64_t &nwalks, uint64_t &max_steps, const double *est) {
    if (!m) return fail(SB200_ERR_INVALID_INPUT, "null matrix");
    if (m->distributed) return fail(SB200_ERR_INVALID_INPUT, "solve_entry needs the whole matrix on one GPU (replicate it)");
    if (m->nrows != m->ncols) return fail(SB200_ERR_INVALID_INPUT, "matrix must be square");
    if (blen != m->nrows)  // estimateEntry: INVALID_DIMENSIONS (src/core/solver.ts:575-581)
        return fail(SB200_ERR_DIMENSION_MISMATCH, "Vector length %llu does not match matrix rows %llu",
                    (unsigned long long)blen, (unsigned long long)m->nrows);
    if (nqueries && (!rows || !est)) return fail(SB200_ERR_INVALID_INPUT, "null query or output array");
    if (blen && !b) return fail(SB200_ERR_INVALID_INPUT, "b is null");
    if (nwalks == 0) {
        if (!(eps > 0.0)) return fail(SB200_ERR_INVALID_INPUT, "epsilon must be positive");  // validatePositiveNumber (:583)
        nwalks = (uint64_t)std::fmax(100.0, std::ceil(1.0 / (eps * eps)));  // :587
    }
    if (max_steps == 0) max_steps = 1000;  // :399
    for (uint64_t q = 0; q < nqueries; q++)
        if (rows[q] >= m->nrows)  // INVALID_PARAMETERS (:560-566)
            return fail(SB200_ERR_INDEX_OUT_OF_BOUNDS, "Row index %llu out of bounds. Matrix has %llu rows",
                        (unsigned long long)rows[q], (unsigned long long)m->nrows);
    return SB200_OK;
}

// queries [0, nqueries) of this call are queries q0 .. of the caller's batch (RNG keys); checked arguments
int32_t entry_run(const sb200_matrix *m, const double *b, uint64_t blen, const uint64_t *rows, uint64_t nqueries,
                  uint64_t nwalks, uint64_t max_steps, uint64_t seed, uint64_t q0, double *est, double *var) {
    if (nqueries == 0) return SB200_OK;
    DeviceGuard g(m->device);
    sb200_matrix *mm = const_cast<sb200_matrix *>(m);
    SB_TRY(matrix_analyse(mm, SB200_MODE_CORRECT, false));
    if (m->first_bad_diag[0] != kNone)  // "Zero diagonal at position i" (:369-372)
        return fail(SB200_ERR_INVALID_SPARSE_MATRIX, "Zero or missing diagonal at position %llu",
                    (unsigned long long)m->first_bad_diag[0]);
    if (m->first_bad_dd != kNone)
        return fail(SB200_ERR_MATRIX_NOT_DIAGONALLY_DOMINANT,
                    "the absorbing-walk estimator needs row dominance (first violating row %llu)",
                    (unsigned long long)m->first_bad_dd);
    const uint64_t per_block = (uint64_t)kWalkThreads * kWalksPerThread;
    const uint64_t chunks = (nwalks + per_block - 1) / per_block;
    if (chunks > 0x7FFFFFFFull) return fail(SB200_ERR_INVALID_INPUT, "at most 2^41 walks per query");
    cudaStream_t st = m->stream;
    DevBuf<double> d_b;
    SB_TRY(d_b.alloc(blen));
    SB_TRY(copy_h2d(d_b.p, b, blen * 8, st));
    constexpr uint64_t kMaxBatch = 65535;  // grid.y
    for (uint64_t off = 0; off < nqueries; off += kMaxBatch) {
        const uint64_t nq = std::min(kMaxBatch, nqueries - off);
        DevBuf<double> d_part, d_est, d_var;
        DevBuf<uint64_t> d_q;
        SB_TRY(d_part.alloc(nq * chunks * 2));
        SB_TRY(d_est.alloc(nq));
        SB_TRY(d_var.alloc(nq));
        SB_TRY(d_q.alloc(nq));
        SB_TRY(copy_h2d(d_q.p, rows + off, nq * 8, st));
        dim3 grid((unsigned)chunks, (unsigned)nq);
        walk_kernel<<<grid, kWalkThreads, 0, st>>>(m->d_vals.p, m->d_cols.p, m->d_row_ptr.p, m->d_dinv[0].p, d_b.p, d_q.p,
                                                   nwalks, (uint32_t)std::min<uint64_t>(max_steps, 0xFFFFFFFFull), seed,
                                                   q0 + off, d_part.p);
        SB_CUDA(cudaGetLastError());
        walk_finalize_kernel<<<(unsigned)((nq + 127) / 128), 128, 0, st>>>(d_part.p, (uint32_t)nq, (uint32_t)chunks, nwalks,
                                                                          d_est.p, d_var.p);
        SB_CUDA(cudaGetLastError());
        SB_TRY(copy_d2h(est + off, d_est.p, nq * 8, st));
        if (var) SB_TRY(copy_d2h(var + off, d_var.p, nq * 8, st));
        SB_CUDA(cudaStreamSynchronize(st));
    }
    return SB200_OK;
}

}  // namespace

extern "C" {

int32_t sb200_solve_entry(const sb200_matrix *m, const double *b, uint64_t blen, const uint64_t *rows,
                          uint64_t nqueries, double eps, uint64_t nwalks, uint64_t max_steps, uint64_t seed,
                          double *est, double *var) {
    clear_error();
    SB_TRY(entry_precheck(m, b, blen, rows, nqueries, eps, nwalks, max_steps, est));
    return entry_run(m, b, blen, rows, nqueries, nwalks, max_steps, seed, 0, est, var);
}

// The multi-GPU form of the batch (entry queries are embarrassingly parallel — replicas): every handle
// holds the whole matrix on its own GPU (build it once per device after sb200_set_device); the queries are cut into
// contiguous slices, one host thread drives each replica. The RNG keys use the position in the WHOLE batch, so the
// estimates are identical to a single sb200_solve_entry call whatever the number of replicas.
int32_t sb200_solve_entry_replicas(const sb200_matrix *const *replicas, int32_t nreplicas, const double *b, uint64_t blen,
                                   const uint64_t *rows, uint64_t nqueries, double eps, uint64_t nwalks, uint64_t max_steps,
                                   uint64_t seed, double *est, double *var) {
    clear_error();
    if (!replicas || nreplicas < 1) return fail(SB200_ERR_INVALID_INPUT, "need at least one replica");
    for (int32_t r = 0; r < nreplicas; r++) {
        if (!replicas[r]) return fail(SB200_ERR_INVALID_INPUT, "replica %d is null", r);
        if (replicas[r]->nrows != replicas[0]->nrows || replicas[r]->ncols != replicas[0]->ncols || replicas[r]->nnz != replicas[0]->nnz)
            return fail(SB200_ERR_INVALID_INPUT, "replica %d is not a copy of replica 0 (%llu x %llu, %llu entries)", r,
                        (unsigned long long)replicas[r]->nrows, (unsigned long long)replicas[r]->ncols, (unsigned long long)replicas[r]->nnz);
    }
    SB_TRY(entry_precheck(replicas[0], b, blen, rows, nqueries, eps, nwalks, max_steps, est));
    const uint64_t per = (nqueries + (uint64_t)nreplicas - 1) / (uint64_t)nreplicas;
    std::vector<int32_t> rc((size_t)nreplicas, SB200_OK);
    std::vector<std::string> msg((size_t)nreplicas);
    std::vector<std::thread> th;
    for (int32_t r = 0; r < nreplicas; r++) {
        const uint64_t q0 = std::min(nqueries, (uint64_t)r * per), q1 = std::min(nqueries, q0 + per);
        if (q1 == q0) continue;
        th.emplace_back([&, r, q0, q1] {
            rc[(size_t)r] = entry_run(replicas[r], b, blen, rows + q0, q1 - q0, nwalks, max_steps, seed, q0, est + q0,
                                      var ? var + q0 : nullptr);
            if (rc[(size_t)r] != SB200_OK) {  // the message lives in the worker's thread-local slot
                char buf[512];
                sb200_last_error(buf, sizeof(buf));
                msg[(size_t)r] = buf;
            }
        });
    }
    for (auto &t : th) t.join();
    for (int32_t r = 0; r < nreplicas; r++)
        if (rc[(size_t)r] != SB200_OK) return fail(rc[(size_t)r], "replica %d: %s", r, msg[(size_t)r].c_str());
    return SB200_OK;
}

}  // extern "C"
