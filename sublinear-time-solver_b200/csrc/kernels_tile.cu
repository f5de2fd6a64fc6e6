// kernels_tile.cu — the TMA-staged tile pipeline ($SUBLINEAR_B200_TILE_CFG = 0..5): kept selectable as the record of
// what was measured; the default kernels live in kernels.cu.
#include "device_util.cuh"

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

// entry point for kernels.cu (launch_tile_kernel / tile_kernel_max_grid)
int32_t launch_tile_pipeline(int cfg, Epilogue epi, const TileKernelArgs &a, cudaStream_t stream, int *max_grid_out) {
    return launch_any(cfg, epi, a, stream, max_grid_out);
}

}  // namespace sb200
