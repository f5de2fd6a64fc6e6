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
// B200 mapping (HBM / L2-bound sparse gather-reduce; tensor cores are irrelevant here). Four bodies, all of which add
// the products of a row LEFT TO RIGHT — the reference's order — with FMA contraction off, so results are bit-identical
// to CSRStorage::multiply_vector and to each other:
//   * sell_kernel : SELL-32 copy of the matrix, lane r owns row r of a 32-row block, no shared memory at all
//                   (the whole 256 KB array serves as L1 for the in-flight gathers); default for vectors <= 48 MB,
//                   band-local matrices and the thin row blocks of the 8-GPU path;
//   * warp_kernel : CSR slices, coalesced 128/256-bit stream loads by the warp, products transposed through a 1 KB
//                   warp-private shared-memory chunk; used for ragged / power-law rows and the StreamingMatrix chunks;
//   * slab_kernel : (kernels_slab.cu) when the gathered vector does not fit the L2 partition of a die: ONE persistent
//                   launch walks <= 28 MB column slabs of the vector, the running row sums carried from slab to slab
//                   (2.4 -> 1.04 L2 sector operations per gather);
//   * tile_kernel : (kernels_tile.cu) TMA-staged tile pipeline (cp.async.bulk + mbarrier, LDGSTS gathers), kept
//                   selectable ($SUBLINEAR_B200_TILE_CFG) as the record of what was measured;
//   * long_rows_* : grid-wide pre-pass for hub rows (> 1024 entries) of power-law graphs.
// The stream (values, column indices) is read with L1::no_allocate + L2 evict_first, the gathered vector with L2
// evict_last (a persisting-L2 set-aside is reserved once per device). The epilogue (diagonal scale, term / solution
// update, squared norm, remote stores of the multi-GPU exchange) is fused; norms are reduced deterministically
// (fixed-shape tree per CTA or warp, fixed-order sum of the partials by the last CTA) and the last CTA also advances the
// device-resident loop state (LoopCtl) — after consuming the peers' flags in a multi-GPU run — so the host never has to
// synchronise per term.
#include <algorithm>

#include "device_util.cuh"

namespace sb200 {

// the TMA tile pipeline (kernels_tile.cu)
int32_t launch_tile_pipeline(int cfg, Epilogue epi, const TileKernelArgs &a, cudaStream_t stream, int *max_grid_out);

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
    if (rc == SB200_OK) rc = a.nslabs > 1 ? launch_slab_kernel(epi, p, stream, nullptr) : launch_warp_any(epi, p, stream, nullptr);
    cudaFreeAsync(scratch, stream);
    return rc;
}

int32_t launch_tile_kernel(int cfg, Epilogue epi, const TileKernelArgs &a, cudaStream_t stream) {
    if (cfg < 0 && a.nslabs > 1) {
        // column slabs: the gather source of one phase is a slab of the vector small enough to stay in the L2 partition of
        // each die; one fused launch, the row sums continue from slab to slab in column = CSR order
        if (!a.acc && !a.out) return fail(SB200_ERR_ALGORITHM, "column-slab passes need a buffer for the partial row sums");
        if (a.nlong > 0) return launch_with_long_rows(epi, a, stream);
        return launch_slab_kernel(epi, a, stream, nullptr);
    }
    if (cfg < 0 && a.sell_ptr != nullptr) return launch_sell_any(epi, a, stream, nullptr);
    if (cfg < 0 && a.nlong > 0) return launch_with_long_rows(epi, a, stream);
    if (cfg < 0) return launch_warp_any(epi, a, stream, nullptr);
    if (epi == EPI_CG) return fail(SB200_ERR_INVALID_INPUT, "the CG epilogue exists in the warp-stream kernel only");
    if (reinterpret_cast<uintptr_t>(a.xin) & 15u)  // 16-byte gathers and the TMA window copy
        return fail(SB200_ERR_INVALID_INPUT, "device vectors must be 16-byte aligned");
    return launch_tile_pipeline(cfg, epi, a, stream, nullptr);
}

// upper bound of the number of partial sums a launch of this configuration writes (callers size `partials` from it)
int tile_kernel_max_grid(int cfg, Epilogue epi) {
    int mg = 0;
    TileKernelArgs dummy{};
    if (cfg >= 0) return launch_tile_pipeline(cfg, epi, dummy, nullptr, &mg) == SB200_OK ? mg : 0;
    int ms = 0, mf = 0;
    if (launch_warp_any(epi, dummy, nullptr, &mg) != SB200_OK) return 0;
    if (launch_sell_any(epi, dummy, nullptr, &ms) != SB200_OK) return 0;  // one partial per warp
    if (launch_slab_kernel(epi, dummy, nullptr, &mf) != SB200_OK) return 0;
    return std::max(mg, std::max(ms, mf)) + 16;  // + scratch of the SELL kernel's last-CTA reduction (2 * 8 doubles)
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
        if (active) row_operands<EPI>(a, row, own, dv, xs, rh);
        double acc = (EPI == EPI_SPMV && a.accumulate) ? xs : 0.0;
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
        if (active) row_epilogue<EPI>(a, row, acc, own, dv, xs, rh, sq, aux);
    }
    if (EPI != EPI_SPMV) {
        const int kind = EPI == EPI_PUSH ? TAIL_TERM : (EPI == EPI_CG ? TAIL_CG_PAP : TAIL_RESID);
        grid_reduce_and_tail<NT>(sq, aux, a.ctl, a.partials, kind, a.it, a.last_in_iter, a.identity_res, a.defer_tail,
                                 a.norm_log, s_red, &s_flag, &a.px);
    }
}

template <int EPI, int NT, int EPL>
static int32_t launch_warp_one(const TileKernelArgs &a, cudaStream_t stream, int *max_grid_out) {
    static int max_grid[64] = {0};
    static std::mutex mu;
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) return fail(SB200_ERR_ALGORITHM, "device index %d out of range", dev);
    std::lock_guard<std::mutex> lk(mu);  // the occupancy cache is filled once per device
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
// warp partial -> global partial array -> the last CTA sums all partials in a fixed order. No shared memory.
template <int NT>
__device__ __forceinline__ void grid_reduce_and_tail_regs(double sq, double aux, LoopCtl *ctl, double *partials, int kind,
                                                          uint32_t it, int last_in_iter, int identity_res, int defer,
                                                          double *norm_log, const PeerExchange *px) {
    constexpr int WARPS = NT / 32;
    const bool p2p = px != nullptr && px->world > 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned nparts = gridDim.x * WARPS;
    sq = warp_sum(sq);
    if (identity_res) aux = warp_sum(aux);
    if (lane == 0) {
        partials[blockIdx.x * WARPS + warp] = sq;
        if (identity_res) partials[nparts + blockIdx.x * WARPS + warp] = aux;
    }
    __syncthreads();
    int last = 0;
    if (threadIdx.x == 0) {
        // one cumulative fence after the CTA barrier orders every thread's stores (partials, peer memory) before the ticket
        if (p2p) __threadfence_system(); else __threadfence();
        last = atomicAdd(&ctl->ticket, 1u) == gridDim.x - 1;
    }
    last = __syncthreads_or(last);
    if (last) {
        // the whole last CTA adds the partials (fixed order for a given grid: thread t takes t, t + NT, ... on four
        // independent accumulators — one warp walking 9 000 partials with a dependent add per L2 round trip cost ~10 us per
        // launch), then the warps' sums go through 2 * WARPS doubles of global scratch behind the partial arrays
        __threadfence();
        double s4[4] = {0.0, 0.0, 0.0, 0.0}, a4[4] = {0.0, 0.0, 0.0, 0.0};
        unsigned i = threadIdx.x;
        for (; i + 3u * NT < nparts; i += 4u * NT) {
#pragma unroll
            for (int q = 0; q < 4; q++) s4[q] += __ldcg(partials + i + q * NT);
            if (identity_res) {
#pragma unroll
                for (int q = 0; q < 4; q++) a4[q] += __ldcg(partials + nparts + i + q * NT);
            }
        }
        for (; i < nparts; i += NT) {
            s4[0] += __ldcg(partials + i);
            if (identity_res) a4[0] += __ldcg(partials + nparts + i);
        }
        double s = warp_sum((s4[0] + s4[1]) + (s4[2] + s4[3]));
        double a2 = identity_res ? warp_sum((a4[0] + a4[1]) + (a4[2] + a4[3])) : 0.0;
        double *scratch = partials + 2u * nparts;
        if (lane == 0) {
            __stcg(scratch + warp, s);
            __stcg(scratch + WARPS + warp, a2);
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            s = 0.0;
            a2 = 0.0;
            for (int w = 0; w < WARPS; w++) {
                s += __ldcg(scratch + w);
                a2 += __ldcg(scratch + WARPS + w);
            }
            ctl->ticket = 0;
            if (p2p) {
                __threadfence_system();
                peer_signal(ctl, *px, s, a2);
                peer_consume(ctl, *px, kind, it, last_in_iter, identity_res, norm_log);  // global sums -> loop logic
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
    static int max_grid[64] = {0};
    static std::mutex mu;
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) return fail(SB200_ERR_ALGORITHM, "device index %d out of range", dev);
    std::lock_guard<std::mutex> lk(mu);  // the occupancy cache is filled once per device
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

}  // namespace sb200
