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

}  // namespace sb200
