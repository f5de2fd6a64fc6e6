// kernels_slab.cu — the fused column-slab kernel: one launch per pass over the matrix when the gathered vector does not
// fit the L2 partition of a die.
//
// Why slabs: a B200 die keeps copies of the lines its SMs fetch from the other die's L2 partition, so a vector that all
// SMs gather from uniformly must fit ONE ~63 MB partition, not the 126 MB total; beyond that every far-homed gather
// costs 2.4 L2 sector operations instead of 1.04 (profiles/r1_gather_footprint_lts_ops.txt) and the kernel sits on the
// L2 sector-throughput cap. Ingest therefore regroups the entries into S column ranges ("slabs") of <= 28 MB of the
// vector; rows are column-sorted, so adding slab after slab onto a carried row sum is exactly the CSR order of
// CSRStorage::multiply_vector (ref src/matrix/sparse.rs:193-203) — results stay bit-identical to the single-pass kernels.
//
// Round 1 ran one warp_kernel launch per slab (3 launches, 3 row_ptr arrays, serialised tails). This kernel does the
// whole pass set in ONE persistent launch:
//   * slab-major order per warp: every warp owns the 32-row blocks gw, gw + nwarps, ... and walks them once per slab, so
//     the whole grid gathers from one slab window at a time without any grid barrier (the carried sums of a block are
//     written and read back by the same thread: program order is all the synchronisation the hand-over needs);
//   * per (slab, block): a u32 entry offset (warp-uniform) and one u16 row length per lane replace the per-slab row_ptr
//     arrays (4 -> 2.1 bytes per row and slab); the row starts come from a warp-shuffle prefix sum;
//   * software pipeline across blocks: offsets two blocks ahead, row lengths + the first 128 entries (column indices
//     128-bit, values 256-bit coalesced, L1::no_allocate, L2 evict_first) one block ahead of the gathers they feed, so a
//     block costs one dependent memory round trip (the gathers) instead of three;
//   * the x[col] gathers go to registers (L2 evict_last), the products pass through a 1 KB warp-private shared-memory
//     chunk so that lane r adds the products of row r left to right;
//   * the last slab runs the fused epilogue (diagonal scale, term / solution update, squared norms, remote stores of the
//     multi-GPU exchange) and the deterministic grid reduction + device-side loop control.
// Hub rows (> kLongRow entries, marked 65535 in the length arrays) hold no entries in the slabs: the long_rows_*
// pre-pass (kernels.cu) sums them grid-wide from the CSR slices and the epilogue only looks the sums up.
#include "device_util.cuh"

namespace sb200 {

namespace {

constexpr int kSlabThreads = 256;
constexpr uint32_t kChunk = 128;  // entries per warp step: 4 per lane

struct StreamRegs {
    uint32_t c[4];
    double v[4];
};

// this lane's 4 entries of the chunk that starts at c0 (multiple of 4): 16-byte / 32-byte aligned vector loads
__device__ __forceinline__ void load_stream(const TileKernelArgs &a, uint32_t c0, uint32_t end, int lane, uint64_t pol,
                                            StreamRegs &r) {
    const uint32_t e = c0 + 4u * (uint32_t)lane;
    if (e < end) {
        asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
                     : "=r"(r.c[0]), "=r"(r.c[1]), "=r"(r.c[2]), "=r"(r.c[3])
                     : "l"(a.slab_cols + e), "l"(pol));
        asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f64 {%0,%1,%2,%3}, [%4], %5;"
                     : "=d"(r.v[0]), "=d"(r.v[1]), "=d"(r.v[2]), "=d"(r.v[3])
                     : "l"(a.slab_vals + e), "l"(pol));
    }
}

// entry range [o0, o1) of block `blk` in slab `s`: two adjacent u32, one load instruction
__device__ __forceinline__ uint32_t load_off_pair(const TileKernelArgs &a, uint32_t s, uint32_t blk, uint32_t nb1, int lane) {
    return __ldg(a.slab_blk + (size_t)s * nb1 + blk + (uint32_t)(lane & 1));
}

}  // namespace

template <int EPI, int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB) slab_kernel(const TileKernelArgs a) {
    constexpr int WARPS = NT / 32;
    __shared__ __align__(16) double s_prod[WARPS][kChunk];
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
    const uint32_t nb1 = nblocks + 1u;
    const uint32_t nwarps = gridDim.x * WARPS;
    const uint32_t gw = blockIdx.x * WARPS + warp;
    const uint32_t S = (uint32_t)a.nslabs;
    double *__restrict__ sp = s_prod[warp];
    double *__restrict__ accbuf = a.acc ? a.acc : a.out;

    double sq = 0.0, aux = 0.0;
    if (gw < nblocks) {
        // the warp's walk: for every slab, its blocks gw, gw + nwarps, ...; (ns, nb) is the step being prefetched
        uint32_t s = 0, blk = gw;                      // current step
        uint32_t ns = 0, nb = gw;                      // next step (lengths + first chunk in flight)
        uint32_t fs = 0, fb = gw;                      // step after next (offsets in flight)
        auto advance = [&](uint32_t &ss, uint32_t &bb) {
            bb += nwarps;
            if (bb >= nblocks) { bb = gw; ss++; }
        };
        // prologue: offsets of step 0 and 1, lengths + stream of step 0
        uint32_t off_cur = load_off_pair(a, s, blk, nb1, lane);
        advance(ns, nb);
        fs = ns; fb = nb;
        uint32_t off_nxt = ns < S ? load_off_pair(a, ns, nb, nb1, lane) : 0u;
        advance(fs, fb);
        uint32_t o0 = __shfl_sync(0xffffffffu, off_cur, 0), o1 = __shfl_sync(0xffffffffu, off_cur, 1);
        uint32_t len_cur = 0;
        {
            const uint32_t row = (blk << 5) + lane;
            if (row < nrows) len_cur = a.slab_len[(size_t)s * a.slab_len_stride + row];
        }
        StreamRegs cur;
        load_stream(a, o0 & ~3u, o1, lane, pol_stream, cur);

        while (s < S) {
            // ---- prefetch: offsets of the step after next, lengths + first chunk of the next step ----
            uint32_t off_far = 0u;
            if (fs < S) off_far = load_off_pair(a, fs, fb, nb1, lane);
            uint32_t n0 = 0, n1 = 0, len_nxt = 0;
            StreamRegs nxt;
            if (ns < S) {
                n0 = __shfl_sync(0xffffffffu, off_nxt, 0);
                n1 = __shfl_sync(0xffffffffu, off_nxt, 1);
                const uint32_t nrow = (nb << 5) + lane;
                if (nrow < nrows) len_nxt = a.slab_len[(size_t)ns * a.slab_len_stride + nrow];
                load_stream(a, n0 & ~3u, n1, lane, pol_stream, nxt);
            }

            // ---- the current block ----
            const uint32_t row = (blk << 5) + lane;
            const bool active = row < nrows;
            const bool first = s == 0, last = s + 1 == S;
            const bool is_long = len_cur == 65535u;
            const uint32_t len = is_long ? 0u : len_cur;
            uint32_t incl = len;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += y;
            }
            const uint32_t re = o0 + incl, rs = re - len;

            double own = 0.0, dv = 0.0, xs = 0.0, rh = 0.0;
            // operands of the epilogue (last slab), or the base of y += A x (first slab)
            if (active && (last || (EPI == EPI_SPMV && first))) row_operands<EPI>(a, row, own, dv, xs, rh);
            double acc = 0.0;
            if (first) {
                if (EPI == EPI_SPMV && a.accumulate) acc = xs;
            } else if (active) {
                acc = a.acc_keep ? __ldcg(accbuf + row) : ld_once_f64_hint(accbuf + row, pol_stream);
            }

            for (uint32_t c0 = o0 & ~3u; c0 < o1; c0 += kChunk) {
                const uint32_t e = c0 + 4u * (uint32_t)lane;
                if (c0 != (o0 & ~3u)) load_stream(a, c0, o1, lane, pol_stream, cur);  // blocks above 128 entries
                double p[4] = {0.0, 0.0, 0.0, 0.0};
                if (e < o1) {
                    // entries outside [o0, o1) belong to neighbouring blocks or to the zero padding: their columns are
                    // valid, their products are never summed
                    double xg[4];
#pragma unroll
                    for (int q = 0; q < 4; q++) xg[q] = ld_gather(a.xin + cur.c[q], pol_gather);
#pragma unroll
                    for (int q = 0; q < 4; q++) p[q] = cur.v[q] * xg[q];
                }
                *reinterpret_cast<double2 *>(sp + 4 * lane) = make_double2(p[0], p[1]);
                *reinterpret_cast<double2 *>(sp + 4 * lane + 2) = make_double2(p[2], p[3]);
                __syncwarp();
                // left-to-right accumulation, the order of CSRStorage::multiply_vector_add (sparse.rs:193-203)
                const uint32_t lo = max(rs, c0), hi = min(re, c0 + kChunk);
                for (uint32_t k = lo; k < hi; k++) acc += sp[k - c0];
                __syncwarp();
            }

            if (active) {
                if (!last) {
                    // hand the running sum to the next slab (same thread reads it back)
                    if (first || o0 != o1) {
                        if (a.acc_keep) __stcg(accbuf + row, acc);
                        else st_stream_f64_hint(accbuf + row, acc, pol_stream);
                    }
                } else {
                    if (is_long) {  // a hub row: summed by the whole grid before this launch (kernels.cu long_rows_*)
                        uint32_t lo = 0, hi = a.nlong;
                        while (lo < hi) {
                            const uint32_t mid = (lo + hi) >> 1;
                            if (a.long_rows[mid] < row) lo = mid + 1; else hi = mid;
                        }
                        acc += a.long_sum[lo];
                    }
                    row_epilogue<EPI>(a, row, acc, own, dv, xs, rh, sq, aux);
                }
            }

            // ---- rotate the pipeline ----
            s = ns; blk = nb;
            ns = fs; nb = fb;
            advance(fs, fb);
            o0 = n0; o1 = n1;
            len_cur = len_nxt;
            cur = nxt;
            off_nxt = off_far;
        }
    }
    if (EPI != EPI_SPMV) {
        const int kind = EPI == EPI_PUSH ? TAIL_TERM : (EPI == EPI_CG ? TAIL_CG_PAP : TAIL_RESID);
        grid_reduce_and_tail<NT>(sq, aux, a.ctl, a.partials, kind, a.it, a.last_in_iter, a.identity_res, a.defer_tail,
                                 a.norm_log, s_red, &s_flag, &a.px);
    }
}

template <int EPI, int MINB>
static int32_t launch_slab_one(const TileKernelArgs &a, cudaStream_t stream, int *max_grid_out) {
    static int max_grid[64] = {0};
    static std::mutex mu;
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) return fail(SB200_ERR_ALGORITHM, "device index %d out of range", dev);
    int mg;
    {
        std::lock_guard<std::mutex> lk(mu);
        if (max_grid[dev] == 0) {
            int per_sm = 0, sms = 0;
            SB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, slab_kernel<EPI, kSlabThreads, MINB>, kSlabThreads, 0));
            SB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
            if (per_sm < 1) return fail(SB200_ERR_ALGORITHM, "slab kernel does not fit on an SM");
            if (per_sm > MINB) per_sm = MINB;
            max_grid[dev] = per_sm * sms;  // persistent grid: a whole number of CTAs per SM (148 SMs on B200)
        }
        mg = max_grid[dev];
    }
    if (max_grid_out) {
        *max_grid_out = mg;
        return SB200_OK;
    }
    if (a.nrows == 0 && EPI == EPI_SPMV) return SB200_OK;
    const uint32_t nblocks = (a.nrows + 31u) / 32u;
    unsigned need = (nblocks + kSlabThreads / 32 - 1) / (kSlabThreads / 32);
    unsigned grid = need < (unsigned)mg ? need : (unsigned)mg;
    if (grid == 0) grid = 1;
    slab_kernel<EPI, kSlabThreads, MINB><<<grid, kSlabThreads, 0, stream>>>(a);
    SB_CUDA(cudaGetLastError());
    return SB200_OK;
}

template <int MINB>
static int32_t launch_slab_b(Epilogue epi, const TileKernelArgs &a, cudaStream_t stream, int *max_grid_out) {
    switch (epi) {
        case EPI_SPMV: return launch_slab_one<EPI_SPMV, MINB>(a, stream, max_grid_out);
        case EPI_PUSH: return launch_slab_one<EPI_PUSH, MINB>(a, stream, max_grid_out);
        case EPI_CG: return launch_slab_one<EPI_CG, MINB>(a, stream, max_grid_out);
        default: return launch_slab_one<EPI_RESID, MINB>(a, stream, max_grid_out);
    }
}

int32_t launch_slab_kernel(Epilogue epi, const TileKernelArgs &a, cudaStream_t stream, int *max_grid_out) {
    // resident CTAs per SM: 3 (default, 85 registers per thread) or 4 ($SUBLINEAR_B200_SLAB_CTAS=4: 64 registers;
    // measurement aid, same results bit for bit)
    static int ctas = [] { const char *e = getenv("SUBLINEAR_B200_SLAB_CTAS"); return e ? atoi(e) : 3; }();
    if (max_grid_out) {  // callers size the partial-sum array for either variant
        int m3 = 0, m4 = 0;
        SB_TRY(launch_slab_b<3>(epi, a, stream, &m3));
        SB_TRY(launch_slab_b<4>(epi, a, stream, &m4));
        *max_grid_out = m3 > m4 ? m3 : m4;
        return SB200_OK;
    }
    return ctas == 4 ? launch_slab_b<4>(epi, a, stream, nullptr) : launch_slab_b<3>(epi, a, stream, nullptr);
}

}  // namespace sb200
