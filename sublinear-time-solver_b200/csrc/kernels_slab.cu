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
#include <algorithm>

#include "device_util.cuh"

namespace sb200 {

namespace {

constexpr int kSlabThreads = 256;
constexpr uint32_t kChunk = 160;  // entries per warp step: 4 per lane (vector loads) + a tail of 1 per lane (scalar loads).
                                  // A 32-row block holds ~107 entries per slab (10 per row, 3 slabs) and ~128 in the slab
                                  // of its diagonal: with 128-entry steps every other such block needed a second,
                                  // nearly empty step (measured: 1.2 steps per block)
constexpr uint32_t kMain = 128;

struct StreamRegs {
    uint32_t c[4];
    double v[4];
    uint32_t ct;  // tail entry
    double vt;
};

// this lane's 4 entries of the chunk that starts at c0 (multiple of 4): 16-byte / 32-byte aligned vector loads
template <bool HINT>
__device__ __forceinline__ void load_stream_cols(const TileKernelArgs &a, uint32_t c0, uint32_t end, int lane, uint64_t pol,
                                                 StreamRegs &r) {
    const uint32_t e = c0 + 4u * (uint32_t)lane;
    if (e < end) {
        if (HINT)
            asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u32 {%0,%1,%2,%3}, [%4], %5;"
                         : "=r"(r.c[0]), "=r"(r.c[1]), "=r"(r.c[2]), "=r"(r.c[3])
                         : "l"(a.slab_cols + e), "l"(pol));
        else
            asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                         : "=r"(r.c[0]), "=r"(r.c[1]), "=r"(r.c[2]), "=r"(r.c[3])
                         : "l"(a.slab_cols + e));
    }
}
template <bool HINT>
__device__ __forceinline__ void load_stream_vals(const TileKernelArgs &a, uint32_t c0, uint32_t end, int lane, uint64_t pol,
                                                 StreamRegs &r) {
    const uint32_t e = c0 + 4u * (uint32_t)lane;
    if (e < end) {
        if (HINT)
            asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f64 {%0,%1,%2,%3}, [%4], %5;"
                         : "=d"(r.v[0]), "=d"(r.v[1]), "=d"(r.v[2]), "=d"(r.v[3])
                         : "l"(a.slab_vals + e), "l"(pol));
        else
            asm volatile("ld.global.nc.L1::no_allocate.v4.f64 {%0,%1,%2,%3}, [%4];"
                         : "=d"(r.v[0]), "=d"(r.v[1]), "=d"(r.v[2]), "=d"(r.v[3])
                         : "l"(a.slab_vals + e));
    }
}
// this lane's 4 entries of the chunk that starts at c0 (multiple of 4): 16-byte / 32-byte aligned vector loads
template <bool HINT>
__device__ __forceinline__ void load_stream(const TileKernelArgs &a, uint32_t c0, uint32_t end, int lane, uint64_t pol,
                                            StreamRegs &r) {
    load_stream_cols<HINT>(a, c0, end, lane, pol, r);
    load_stream_vals<HINT>(a, c0, end, lane, pol, r);
    const uint32_t t = c0 + kMain + (uint32_t)lane;
    if (t < end) {
        r.ct = ld_stream_u32_hint(a.slab_cols + t, pol);
        r.vt = ld_stream_f64_hint(a.slab_vals + t, pol);
    }
}

// the four x[col] gathers of a lane as ONE asm statement: all four loads are issued back to back (left to the compiler,
// register pressure makes it wait for the first two before issuing the others: three serialised round trips per chunk)
template <bool HINT>
__device__ __forceinline__ void gather4(const double *x, const uint32_t (&c)[4], uint64_t pol, double (&g)[4]) {
    if (!HINT) {
        asm volatile(
            "{\n\t.reg .u64 a0, a1, a2, a3;\n\t"
            "mad.wide.u32 a0, %4, 8, %8;\n\t"
            "mad.wide.u32 a1, %5, 8, %8;\n\t"
            "mad.wide.u32 a2, %6, 8, %8;\n\t"
            "mad.wide.u32 a3, %7, 8, %8;\n\t"
            "ld.global.nc.f64 %0, [a0];\n\t"
            "ld.global.nc.f64 %1, [a1];\n\t"
            "ld.global.nc.f64 %2, [a2];\n\t"
            "ld.global.nc.f64 %3, [a3];\n\t}"
            : "=d"(g[0]), "=d"(g[1]), "=d"(g[2]), "=d"(g[3])
            : "r"(c[0]), "r"(c[1]), "r"(c[2]), "r"(c[3]), "l"(x));
        return;
    }
    asm volatile(
        "{\n\t.reg .u64 a0, a1, a2, a3;\n\t"
        "mad.wide.u32 a0, %4, 8, %8;\n\t"
        "mad.wide.u32 a1, %5, 8, %8;\n\t"
        "mad.wide.u32 a2, %6, 8, %8;\n\t"
        "mad.wide.u32 a3, %7, 8, %8;\n\t"
        "ld.global.nc.L2::cache_hint.f64 %0, [a0], %9;\n\t"
        "ld.global.nc.L2::cache_hint.f64 %1, [a1], %9;\n\t"
        "ld.global.nc.L2::cache_hint.f64 %2, [a2], %9;\n\t"
        "ld.global.nc.L2::cache_hint.f64 %3, [a3], %9;\n\t}"
        : "=d"(g[0]), "=d"(g[1]), "=d"(g[2]), "=d"(g[3])
        : "r"(c[0]), "r"(c[1]), "r"(c[2]), "r"(c[3]), "l"(x), "l"(pol));
}

}  // namespace

// shared-memory position of product i of a chunk: 16 bytes of padding after every 128 bytes, so that the two 128-bit
// stores of a lane (elements 4 lane .. 4 lane + 3, 32 bytes apart from the neighbouring lane's) hit each bank once
__device__ __forceinline__ uint32_t prod_pos(uint32_t i) { return i + ((i >> 3) & ~1u); }
constexpr uint32_t kProdSlots = kChunk + 2 * (kChunk / 16);  // doubles per warp incl. padding

template <int EPI, int NT, int MINB, bool PF, bool HINT>
__global__ void __launch_bounds__(NT, MINB) slab_kernel(const TileKernelArgs a) {
    constexpr int WARPS = NT / 32;
    __shared__ __align__(16) double s_prod[WARPS][kProdSlots];
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
    const double *__restrict__ sp = s_prod[warp];
    double2 *const my_slot = reinterpret_cast<double2 *>(s_prod[warp] + prod_pos(4u * (uint32_t)lane));
    double *const my_tail = s_prod[warp] + prod_pos(kMain + (uint32_t)lane);
    double *__restrict__ accbuf = a.acc ? a.acc : a.out;

    double sq = 0.0, aux = 0.0;
    // the warp's walk: slab after slab, inside a slab its blocks gw, gw + nwarps, ... (all warps gather from one slab
    // window of the vector at a time; the carried row sums of a block are written and read back by the same thread)
    for (uint32_t s = 0; s < S && gw < nblocks; s++) {
        const uint32_t *__restrict__ blkp = a.slab_blk + (size_t)s * nb1 + (uint32_t)(lane & 1);
        const uint16_t *__restrict__ relp = a.slab_rel + (size_t)s * a.slab_rel_stride + lane;
        const bool first = s == 0, last = s + 1 == S;
        // block offsets + row offsets run one block ahead of their use [PF: the first chunk of the stream too, issued
        // once the registers of the current chunk are free]
        uint32_t off_nxt = __ldg(blkp + gw);
        uint32_t rel_nxt = __ldg(relp + ((size_t)gw << 5));
        StreamRegs st;
        uint32_t pf0 = 0, pf1 = 0;  // [PF] entry range whose first chunk is in `st`
        if (PF) {
            pf0 = __shfl_sync(0xffffffffu, off_nxt, 0);
            pf1 = __shfl_sync(0xffffffffu, off_nxt, 1);
            load_stream<HINT>(a, pf0 & ~3u, pf1, lane, pol_stream, st);
        }
        for (uint32_t blk = gw; blk < nblocks; blk += nwarps) {
            const uint32_t o0 = PF ? pf0 : __shfl_sync(0xffffffffu, off_nxt, 0);
            const uint32_t o1 = PF ? pf1 : __shfl_sync(0xffffffffu, off_nxt, 1);
            const uint32_t rel_cur = rel_nxt;
            if (!PF) load_stream<HINT>(a, o0 & ~3u, o1, lane, pol_stream, st);
            const uint32_t nblk = blk + nwarps;
            if (nblk < nblocks) {
                off_nxt = __ldg(blkp + nblk);
                rel_nxt = __ldg(relp + ((size_t)nblk << 5));
            }
            const uint32_t row = (blk << 5) + lane;
            const bool active = row < nrows;
            const bool is_long = (rel_cur & 0x8000u) != 0u;
            const uint32_t rs = o0 + (rel_cur & 0x7FFFu);   // row r starts where the rows before it (in this block) end
            uint32_t re = __shfl_down_sync(0xffffffffu, rs, 1);
            if (lane == 31) re = o1;

            double own = 0.0, dv = 0.0, xs = 0.0, rh = 0.0;
            // operands of the epilogue (last slab), or the base of y += A x (first slab)
            if (active && (last || (EPI == EPI_SPMV && first))) row_operands<EPI>(a, row, own, dv, xs, rh);
            double acc = 0.0;
            if (first) {
                if (EPI == EPI_SPMV && a.accumulate) acc = xs;
            } else if (active) {
                acc = (a.acc_keep || !HINT) ? __ldcg(accbuf + row) : ld_once_f64_hint(accbuf + row, pol_stream);
            }

            for (uint32_t c0 = o0 & ~3u; c0 < o1; c0 += kChunk) {
                const uint32_t e = c0 + 4u * (uint32_t)lane;
                if (c0 != (o0 & ~3u)) load_stream<HINT>(a, c0, o1, lane, pol_stream, st);  // blocks above 128 entries
                double p[4] = {0.0, 0.0, 0.0, 0.0};
                double xg[4];
                const bool has_tail = c0 + kMain + (uint32_t)lane < o1;
                double xt = 0.0;
                if (e < o1) gather4<HINT>(a.xin, st.c, pol_gather, xg);
                if (has_tail) xt = ld_gather(a.xin + st.ct, pol_gather);
                if (PF && c0 + kChunk >= o1 && nblk < nblocks) {
                    // the column-index registers are free: start on the next block while the gathers are in flight
                    pf0 = __shfl_sync(0xffffffffu, off_nxt, 0);
                    pf1 = __shfl_sync(0xffffffffu, off_nxt, 1);
                    load_stream_cols<HINT>(a, pf0 & ~3u, pf1, lane, pol_stream, st);
                }
                if (e < o1) {
                    // entries outside [o0, o1) belong to neighbouring blocks or to the zero padding: their columns are
                    // valid, their products are never summed
#pragma unroll
                    for (int q = 0; q < 4; q++) p[q] = st.v[q] * xg[q];
                }
                const double tail_prod = has_tail ? st.vt * xt : 0.0;
                if (PF && c0 + kChunk >= o1 && nblk < nblocks) {
                    load_stream_vals<HINT>(a, pf0 & ~3u, pf1, lane, pol_stream, st);
                    const uint32_t t = (pf0 & ~3u) + kMain + (uint32_t)lane;
                    if (t < pf1) {
                        st.ct = ld_stream_u32_hint(a.slab_cols + t, pol_stream);
                        st.vt = ld_stream_f64_hint(a.slab_vals + t, pol_stream);
                    }
                }
                my_slot[0] = make_double2(p[0], p[1]);
                my_slot[1] = make_double2(p[2], p[3]);
                my_tail[0] = tail_prod;
                __syncwarp();
                // left-to-right accumulation, the order of CSRStorage::multiply_vector_add (sparse.rs:193-203); the trip
                // count is the longest row of the warp (uniform loop, predicated adds)
                const uint32_t lo = max(rs, c0), hi = min(re, c0 + kChunk);
                const uint32_t cnt = hi > lo ? hi - lo : 0u;
                const uint32_t longest = __reduce_max_sync(0xffffffffu, cnt);
                const uint32_t j0 = cnt ? lo - c0 : 0u;
#pragma unroll 2
                for (uint32_t k = 0; k < longest; k++)
                    if (k < cnt) acc += sp[prod_pos(j0 + k)];
                __syncwarp();
            }

            if (active) {
                if (!last) {
                    // hand the running sum to the next slab (same thread reads it back)
                    if (first || o0 != o1) {
                        if (a.acc_keep || !HINT) __stcg(accbuf + row, acc);
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
        }
    }
    if (EPI != EPI_SPMV) {
        const int kind = EPI == EPI_PUSH ? TAIL_TERM : (EPI == EPI_CG ? TAIL_CG_PAP : TAIL_RESID);
        grid_reduce_and_tail<NT>(sq, aux, a.ctl, a.partials, kind, a.it, a.last_in_iter, a.identity_res, a.defer_tail,
                                 a.norm_log, s_red, &s_flag, &a.px);
    }
}

template <int EPI, int MINB, bool PF, bool HINT>
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
            SB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, slab_kernel<EPI, kSlabThreads, MINB, PF, HINT>, kSlabThreads, 0));
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
    slab_kernel<EPI, kSlabThreads, MINB, PF, HINT><<<grid, kSlabThreads, 0, stream>>>(a);
    SB_CUDA(cudaGetLastError());
    return SB200_OK;
}

template <int MINB, bool PF, bool HINT>
static int32_t launch_slab_b(Epilogue epi, const TileKernelArgs &a, cudaStream_t stream, int *max_grid_out) {
    switch (epi) {
        case EPI_SPMV: return launch_slab_one<EPI_SPMV, MINB, PF, HINT>(a, stream, max_grid_out);
        case EPI_PUSH: return launch_slab_one<EPI_PUSH, MINB, PF, HINT>(a, stream, max_grid_out);
        case EPI_CG: return launch_slab_one<EPI_CG, MINB, PF, HINT>(a, stream, max_grid_out);
        default: return launch_slab_one<EPI_RESID, MINB, PF, HINT>(a, stream, max_grid_out);
    }
}

int32_t launch_slab_kernel(Epilogue epi, const TileKernelArgs &a, cudaStream_t stream, int *max_grid_out) {
    // variants ($SUBLINEAR_B200_SLAB_VARIANT; measurement aid, same results bit for bit):
    //   CTAs per SM 3 / 4, late stream prefetch on / off, L2 cache-policy hints on / off
    static int variant = [] { const char *e = getenv("SUBLINEAR_B200_SLAB_VARIANT"); return e ? atoi(e) : 1; }();
    if (max_grid_out) {  // callers size the partial-sum array for any variant (one partial per CTA)
        int m = 0;
        SB_TRY((launch_slab_b<4, false, true>(epi, a, stream, &m)));
        *max_grid_out = m;
        return SB200_OK;
    }
    switch (variant) {
        case 0: return launch_slab_b<3, true, true>(epi, a, stream, nullptr);
        case 2: return launch_slab_b<4, false, false>(epi, a, stream, nullptr);
        case 3: return launch_slab_b<4, true, true>(epi, a, stream, nullptr);
        case 4: return launch_slab_b<3, true, false>(epi, a, stream, nullptr);
        case 5: return launch_slab_b<4, true, false>(epi, a, stream, nullptr);
        default: return launch_slab_b<4, false, true>(epi, a, stream, nullptr);
    }
}

}  // namespace sb200
