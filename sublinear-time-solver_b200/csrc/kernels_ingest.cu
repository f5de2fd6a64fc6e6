// kernels_ingest.cu — one-off device passes at matrix ingest: SELL-32 layout, column-slab split (+ prefix sum), and the
// K4 setup pass (dominance, diagonal, D^-1).
#include "device_util.cuh"

namespace sb200 {

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
// column-slab split at ingest (one-off; matrix.cu build_slabs). Warp per 32-row block, lane r <-> row r.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) slab_count_kernel(const uint32_t *__restrict__ cols, const uint32_t *__restrict__ row_ptr,
                                                         uint32_t nrows, uint32_t slab_width, int nslabs, uint32_t long_row,
                                                         uint16_t *__restrict__ rel, uint64_t rel_stride,
                                                         uint32_t *__restrict__ blk, unsigned long long *flags) {
    const int lane = threadIdx.x & 31;
    const uint32_t nblocks = (nrows + 31u) >> 5, nb1 = nblocks + 1u;
    const uint32_t nwarps = gridDim.x * (blockDim.x >> 5);
    for (uint32_t b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); b < nblocks; b += nwarps) {
        const uint32_t row = (b << 5) + lane;
        uint32_t cnt[kMaxSlabs];
#pragma unroll
        for (int s = 0; s < kMaxSlabs; s++) cnt[s] = 0u;
        bool is_long = false;
        if (row < nrows) {
            const uint32_t rs = row_ptr[row], re = row_ptr[row + 1];
            is_long = re - rs > long_row;
            uint32_t prev = 0;
            bool bad = false;
            for (uint32_t k = rs; k < re; k++) {
                const uint32_t c = cols[k];
                bad |= c < prev;
                prev = c;
                const uint32_t s = min(c / slab_width, (uint32_t)(nslabs - 1));
#pragma unroll
                for (int q = 0; q < kMaxSlabs; q++) cnt[q] += (uint32_t)(q == (int)s);
            }
            int used = 0;
#pragma unroll
            for (int s = 0; s < kMaxSlabs; s++) used += cnt[s] != 0u;
            if (bad) flags[0] = 1ull;
            if (used > 1) atomicAdd(flags + 1, (unsigned long long)(re - rs));  // entries of rows that reach into several slabs
        }
#pragma unroll
        for (int s = 0; s < kMaxSlabs; s++) {
            if (s >= nslabs) break;
            const uint32_t c = is_long ? 0u : cnt[s];  // hub rows keep their entries in the CSR slices only
            uint32_t incl = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t y = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += y;
            }
            // offset of the row's first entry inside the block (< 32 * kLongRow = 2^15), bit 15 marks a hub row; lanes
            // past the last row store the block total, i.e. an empty row
            rel[(size_t)s * rel_stride + row] = (uint16_t)((incl - c) | (is_long ? 0x8000u : 0u));
            const uint32_t tot = __shfl_sync(0xffffffffu, incl, 31);
            if (lane == 0) blk[(size_t)s * nb1 + b] = tot;
        }
    }
}

__global__ void __launch_bounds__(256) slab_fill_kernel(const double *__restrict__ vals, const uint32_t *__restrict__ cols,
                                                        const uint32_t *__restrict__ row_ptr, uint32_t nrows,
                                                        uint32_t slab_width, int nslabs, const uint16_t *__restrict__ rel,
                                                        uint64_t rel_stride, const uint32_t *__restrict__ blk,
                                                        uint32_t *__restrict__ sc, double *__restrict__ sv) {
    const int lane = threadIdx.x & 31;
    const uint32_t nblocks = (nrows + 31u) >> 5, nb1 = nblocks + 1u;
    const uint32_t nwarps = gridDim.x * (blockDim.x >> 5);
    for (uint32_t b = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); b < nblocks; b += nwarps) {
        const uint32_t row = (b << 5) + lane;
        uint32_t pos[kMaxSlabs];
        bool is_long = false;
#pragma unroll
        for (int s = 0; s < kMaxSlabs; s++) {
            pos[s] = 0u;
            if (s >= nslabs) continue;
            const uint32_t r = rel[(size_t)s * rel_stride + row];
            is_long |= (r & 0x8000u) != 0u;
            pos[s] = blk[(size_t)s * nb1 + b] + (r & 0x7FFFu);
        }
        if (row >= nrows || is_long) continue;
        for (uint32_t k = row_ptr[row]; k < row_ptr[row + 1]; k++) {  // in CSR order: the order inside a slab row is kept
            const uint32_t c = cols[k];
            const double v = vals[k];
            const uint32_t s = min(c / slab_width, (uint32_t)(nslabs - 1));
            uint32_t p = 0;
#pragma unroll
            for (int q = 0; q < kMaxSlabs; q++)
                if (q == (int)s) { p = pos[q]; pos[q]++; }
            sc[p] = c;
            sv[p] = v;
        }
    }
}

static unsigned block_grid(uint32_t nrows) {
    const uint32_t nblocks = (nrows + 31u) / 32u;
    unsigned grid = (nblocks + 7) / 8;
    if (grid > 148 * 16) grid = 148 * 16;
    return grid ? grid : 1;
}

int32_t launch_slab_count(const uint32_t *cols, const uint32_t *row_ptr, uint32_t nrows, uint32_t slab_width, int nslabs,
                          uint32_t long_row, uint16_t *rel, uint64_t rel_stride, uint32_t *blk, unsigned long long *flags,
                          cudaStream_t stream) {
    slab_count_kernel<<<block_grid(nrows), 256, 0, stream>>>(cols, row_ptr, nrows, slab_width, nslabs, long_row, rel,
                                                             rel_stride, blk, flags);
    SB_CUDA(cudaGetLastError());
    return SB200_OK;
}

int32_t launch_slab_fill(const double *vals, const uint32_t *cols, const uint32_t *row_ptr, uint32_t nrows,
                         uint32_t slab_width, int nslabs, const uint16_t *rel, uint64_t rel_stride, const uint32_t *blk,
                         uint32_t *slab_cols, double *slab_vals, cudaStream_t stream) {
    slab_fill_kernel<<<block_grid(nrows), 256, 0, stream>>>(vals, cols, row_ptr, nrows, slab_width, nslabs, rel, rel_stride,
                                                            blk, slab_cols, slab_vals);
    SB_CUDA(cudaGetLastError());
    return SB200_OK;
}

__global__ void add_u32_kernel(uint32_t *data, uint64_t n, uint32_t v) {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) data[i] += v;
}

int32_t launch_add_u32(uint32_t *data, uint64_t n, uint32_t v, cudaStream_t stream) {
    if (n == 0 || v == 0) return SB200_OK;
    uint64_t g = (n + 255) / 256;
    add_u32_kernel<<<(unsigned)(g > 148ull * 8 ? 148ull * 8 : g), 256, 0, stream>>>(data, n, v);
    SB_CUDA(cudaGetLastError());
    return SB200_OK;
}

// ---------------------------------------------------------------------------------------------------------
// exclusive prefix sum of n u32 counts in place (data has n + 1 entries: data[n] receives the total). Three launches over
// tiles of 4096 elements: per-tile sums, a one-CTA scan of the (<= a few thousand) tile sums, tile-local scans + the tile's
// base. Sums are carried in 64 bits; the stored offsets are u32 (callers check the total).
// ---------------------------------------------------------------------------------------------------------
constexpr int kScanThreads = 256;
constexpr int kScanPer = 16;                       // consecutive elements per thread
constexpr uint64_t kScanTile = (uint64_t)kScanThreads * kScanPer;

__device__ __forceinline__ unsigned long long block_exclusive_scan_u64(unsigned long long v, unsigned long long *s_warp,
                                                                      unsigned long long *total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned long long x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= o) x += y;
    }
    __syncthreads();  // s_warp may still be read from a previous call
    if (lane == 31) s_warp[warp] = x;
    __syncthreads();
    if (warp == 0) {
        unsigned long long w = lane < (int)(blockDim.x >> 5) ? s_warp[lane] : 0ull;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long y = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o) w += y;
        }
        s_warp[lane] = w;  // inclusive over warps
    }
    __syncthreads();
    if (total) *total = s_warp[(blockDim.x >> 5) - 1];
    return (warp ? s_warp[warp - 1] : 0ull) + (x - v);
}

__global__ void __launch_bounds__(kScanThreads) scan_tile_sums_kernel(const uint32_t *__restrict__ data, uint64_t n,
                                                                      unsigned long long *__restrict__ tile_sums) {
    __shared__ unsigned long long s_warp[32];
    const uint64_t base = blockIdx.x * kScanTile + (uint64_t)threadIdx.x * kScanPer;
    unsigned long long v = 0;
#pragma unroll
    for (int q = 0; q < kScanPer; q++)
        if (base + q < n) v += data[base + q];
    unsigned long long tot;
    block_exclusive_scan_u64(v, s_warp, &tot);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = tot;
}

// one CTA: exclusive scan of the tile sums in place; tile_sums[ntiles] receives the grand total
__global__ void __launch_bounds__(1024) scan_tile_bases_kernel(unsigned long long *tile_sums, uint64_t ntiles) {
    __shared__ unsigned long long s_warp[32];
    __shared__ unsigned long long s_carry;
    if (threadIdx.x == 0) s_carry = 0ull;
    __syncthreads();
    for (uint64_t base = 0; base < ntiles; base += 1024) {
        const uint64_t i = base + threadIdx.x;
        const unsigned long long v = i < ntiles ? tile_sums[i] : 0ull;
        unsigned long long tot;
        const unsigned long long ex = block_exclusive_scan_u64(v, s_warp, &tot);
        const unsigned long long carry = s_carry;
        if (i < ntiles) tile_sums[i] = carry + ex;
        __syncthreads();
        if (threadIdx.x == 0) s_carry = carry + tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) tile_sums[ntiles] = s_carry;
}

__global__ void __launch_bounds__(kScanThreads) scan_apply_kernel(uint32_t *__restrict__ data, uint64_t n,
                                                                  const unsigned long long *__restrict__ tile_sums,
                                                                  uint64_t ntiles) {
    __shared__ unsigned long long s_warp[32];
    const uint64_t base = blockIdx.x * kScanTile + (uint64_t)threadIdx.x * kScanPer;
    uint32_t e[kScanPer];
    unsigned long long v = 0;
#pragma unroll
    for (int q = 0; q < kScanPer; q++) {
        e[q] = base + q < n ? data[base + q] : 0u;
        v += e[q];
    }
    unsigned long long run = tile_sums[blockIdx.x] + block_exclusive_scan_u64(v, s_warp, nullptr);
#pragma unroll
    for (int q = 0; q < kScanPer; q++) {
        if (base + q < n) data[base + q] = (uint32_t)run;
        run += e[q];
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) data[n] = (uint32_t)tile_sums[ntiles];
}

int32_t device_exclusive_scan_u32(uint32_t *data, uint64_t n, uint64_t *total, cudaStream_t stream) {
    const uint64_t ntiles = (n + kScanTile - 1) / kScanTile;
    if (ntiles > 0x7FFFFFFFull) return fail(SB200_ERR_INVALID_INPUT, "prefix sum over %llu elements", (unsigned long long)n);
    DevBuf<unsigned long long> t;
    SB_TRY(t.alloc(ntiles + 1));
    if (ntiles == 0) {
        SB_CUDA(cudaMemsetAsync(t.p, 0, 8, stream));
        SB_CUDA(cudaMemsetAsync(data, 0, 4, stream));
    } else {
        scan_tile_sums_kernel<<<(unsigned)ntiles, kScanThreads, 0, stream>>>(data, n, t.p);
        scan_tile_bases_kernel<<<1, 1024, 0, stream>>>(t.p, ntiles);
        scan_apply_kernel<<<(unsigned)ntiles, kScanThreads, 0, stream>>>(data, n, t.p, ntiles);
        SB_CUDA(cudaGetLastError());
    }
    unsigned long long h = 0;
    SB_CUDA(cudaMemcpyAsync(&h, t.p + ntiles, 8, cudaMemcpyDeviceToHost, stream));
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

}  // namespace sb200
