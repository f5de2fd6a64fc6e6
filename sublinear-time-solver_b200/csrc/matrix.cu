// matrix.cu — SparseMatrix ingest (host side, mirrors src/matrix/mod.rs + src/matrix/sparse.rs of the
// reference), upload to HBM, tile table construction, cached analysis and the Matrix trait entry points.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <numeric>

#include "matrix.hpp"

using namespace sb200;

sb200_matrix::~sb200_matrix() {
    int prev = -1;
    cudaGetDevice(&prev);
    cudaSetDevice(device);
    pool.clear();
    if (axb_cache && axb_cache_free) axb_cache_free(axb_cache);
    d_vals.release();
    d_cols.release();
    d_row_ptr.release();
    d_tiles.release();
    d_sell_ptr.release();
    d_sell_cols.release();
    d_sell_vals.release();
    d_long_rows.release();
    d_long_first.release();
    d_long_chunks.release();
    d_slab_vals.release();
    d_slab_cols.release();
    d_slab_blk.release();
    d_slab_rel.release();
    d_dinv[0].release();
    d_dinv[1].release();
    if (stream) cudaStreamDestroy(stream);
    if (prev >= 0) cudaSetDevice(prev);
}

namespace sb200 {

Workspace::~Workspace() {
    if (h_ctl) cudaFreeHost(h_ctl);
    if (ev0) cudaEventDestroy(ev0);
    if (ev1) cudaEventDestroy(ev1);
}

int32_t Workspace::ensure(uint64_t nl, uint64_t nf, size_t npartials) {
    if (nl > n_local || x.p == nullptr) {
        SB_TRY(x.alloc(nl));
        SB_TRY(c.alloc(nl));
        SB_TRY(b.alloc(nl));
        SB_TRY(tmp.alloc(nl));
        n_local = nl;
    }
    if (nf > n_full || t[0].p == nullptr) {
        SB_TRY(t[0].alloc(nf));
        SB_TRY(t[1].alloc(nf));
        n_full = nf;
    }
    if (partials.n < npartials) SB_TRY(partials.alloc(npartials));
    if (!ctl.p) SB_TRY(ctl.alloc(1));
    if (!h_ctl) SB_CUDA(cudaHostAlloc((void **)&h_ctl, sizeof(LoopCtl), cudaHostAllocDefault));
    if (!ev0) SB_CUDA(cudaEventCreate(&ev0));
    if (!ev1) SB_CUDA(cudaEventCreate(&ev1));
    return SB200_OK;
}

uint64_t Workspace::bytes() const {
    return 8ull * (x.n + c.n + b.n + tmp.n + t[0].n + t[1].n + partials.n + norm_log.n) + sizeof(LoopCtl);
}

void matrix_retain(sb200_matrix *m) {
    if (m) m->refcount.fetch_add(1, std::memory_order_relaxed);
}

void matrix_release(sb200_matrix *m) {
    if (m && m->refcount.fetch_sub(1, std::memory_order_acq_rel) == 1) delete m;
}

std::unique_ptr<Workspace> matrix_acquire_ws(sb200_matrix *m) {
    std::lock_guard<std::mutex> lk(m->mu);
    if (!m->pool.empty()) {
        auto ws = std::move(m->pool.back());
        m->pool.pop_back();
        return ws;
    }
    return std::unique_ptr<Workspace>(new Workspace());
}

void matrix_release_ws(sb200_matrix *m, std::unique_ptr<Workspace> ws) {
    std::lock_guard<std::mutex> lk(m->mu);
    if (m->pool.size() < 2) m->pool.push_back(std::move(ws));
}

// Group rows into tiles of <= threads rows and <= cap non-zeros; a row above cap becomes its own tile.
static void build_tiles(const uint32_t *row_ptr, uint64_t nrows, TileCfg cfg, std::vector<TileDesc> &tiles) {
    tiles.clear();
    tiles.reserve(nrows / (size_t)cfg.rows + 16);
    uint64_t r = 0;
    while (r < nrows) {
        uint64_t r1 = r;
        const uint64_t rmax = std::min<uint64_t>(nrows, r + (uint64_t)cfg.rows);
        const uint64_t base = row_ptr[r];
        while (r1 < rmax && (uint64_t)row_ptr[r1 + 1] - base <= (uint64_t)cfg.cap) r1++;
        if (r1 == r) r1 = r + 1;  // long row
        tiles.push_back(TileDesc{(uint32_t)r, row_ptr[r]});
        r = r1;
    }
    tiles.push_back(TileDesc{(uint32_t)nrows, row_ptr[nrows]});
}

// SELL-32 copy for the hot kernels: per block of 32 rows `width` = its longest row, slab offsets by prefix sum on the
// host (O(n) over the host row_ptr copy), the transposition itself on the device from the uploaded CSR slices.
// $SUBLINEAR_B200_SELL = 1 / 0 forces / forbids the layout; default: used when it costs <= 25 % extra slots.
static int32_t build_sell(sb200_matrix *m) {
    m->use_sell = false;
    if (m->tile_cfg >= 0 || m->nrows == 0) return SB200_OK;  // TMA tile pipeline selected: CSR only
    const char *e = getenv("SUBLINEAR_B200_SELL");
    const int force = e ? atoi(e) : -1;
    if (force == 0) return SB200_OK;
    const uint32_t *rp = m->h_row_ptr.data();
    const uint64_t nrows = m->nrows, nblocks = (nrows + 31) / 32;
    std::vector<uint32_t> sp(nblocks + 1);
    uint64_t slabs = 0, max_width = 0;
    for (uint64_t b = 0; b < nblocks; b++) {
        sp[b] = (uint32_t)slabs;
        uint32_t w = 0;
        const uint64_t r1 = std::min(nrows, b * 32 + 32);
        for (uint64_t r = b * 32; r < r1; r++) w = std::max(w, rp[r + 1] - rp[r]);
        slabs += w;
        max_width = std::max<uint64_t>(max_width, w);
        if (slabs >= 0xFFFFFFF0ull) return SB200_OK;  // slab offsets are u32: keep the CSR kernels
    }
    sp[nblocks] = (uint32_t)slabs;
    const uint64_t slots = slabs * 32;
    if (force != 1 && (slots > m->nnz + m->nnz / 4 + 2048 || max_width > 4096)) return SB200_OK;
    SB_TRY(m->d_sell_ptr.alloc(nblocks + 1));
    SB_TRY(m->d_sell_cols.alloc(slots));
    SB_TRY(m->d_sell_vals.alloc(slots));
    SB_TRY(copy_h2d(m->d_sell_ptr.p, sp.data(), (nblocks + 1) * sizeof(uint32_t), m->stream));
    SB_TRY(launch_csr_to_sell(m->d_vals.p, m->d_cols.p, m->d_row_ptr.p, (uint32_t)nrows, m->d_sell_ptr.p, m->d_sell_cols.p,
                              m->d_sell_vals.p, m->stream));
    SB_CUDA(cudaStreamSynchronize(m->stream));  // `sp` is staged asynchronously
    m->sell_slabs = slabs;
    m->use_sell = true;
    return SB200_OK;
}

// Hub rows: rows with more than `threshold` entries get a chunk table so that the whole grid can sum them before the
// row-block kernel runs (kernels.cu long_rows_*). O(n) over the host row_ptr copy.
static int32_t build_long_rows(sb200_matrix *m) {
    const uint32_t threshold = kLongRow;
    m->nlong = m->nlong_chunks = 0;
    const uint32_t *rp = m->h_row_ptr.data();
    std::vector<uint32_t> rows, first;
    std::vector<uint2> chunks;
    for (uint64_t r = 0; r < m->nrows; r++) {
        const uint32_t rs = rp[r], re = rp[r + 1];
        if (re - rs <= threshold) continue;
        rows.push_back((uint32_t)r);
        first.push_back((uint32_t)chunks.size());
        for (uint64_t s = rs; s < re; s += kLongChunk)
            chunks.push_back(make_uint2((uint32_t)s, (uint32_t)std::min<uint64_t>(s + kLongChunk, re)));
    }
    if (rows.empty()) return SB200_OK;
    first.push_back((uint32_t)chunks.size());
    SB_TRY(m->d_long_rows.alloc(rows.size()));
    SB_TRY(m->d_long_first.alloc(first.size()));
    SB_TRY(m->d_long_chunks.alloc(chunks.size()));
    SB_TRY(copy_h2d(m->d_long_rows.p, rows.data(), rows.size() * sizeof(uint32_t), m->stream));
    SB_TRY(copy_h2d(m->d_long_first.p, first.data(), first.size() * sizeof(uint32_t), m->stream));
    SB_TRY(copy_h2d(m->d_long_chunks.p, chunks.data(), chunks.size() * sizeof(uint2), m->stream));
    SB_CUDA(cudaStreamSynchronize(m->stream));
    m->nlong = (uint32_t)rows.size();
    m->nlong_chunks = (uint32_t)chunks.size();
    return SB200_OK;
}

// Column-slab layout for the hot kernels. Measured on a B200 (profiles/r1_slab_timing.log): random 8-byte
// gathers cost one L2 sector operation while the gather source fits the L2 partition of each die (<= ~40 MB) and 2.4
// once it does not (80 MB: every far-homed line is looked up near, fetched over the fabric and filled again), and the
// push kernel runs at the chip's L2 sector-throughput cap. Regrouping the entries into column slabs of <= 28 MB of the
// vector and walking slab after slab keeps the gathers at one operation each; rows are column-sorted, so carrying the
// row sum from slab to slab adds the products in exactly the CSR order.
// $SUBLINEAR_B200_SLABS = 0 forbids, 2..8 forces that many slabs. Default: the gathered vector is > 48 MB and <= 8 * 28 MB
// and the matrix holds enough entries per vector sector for the window to pay: $SUBLINEAR_B200_SLAB_MIN_DENSITY entries
// per 32-byte sector of the vector, default 8. A row block of the multi-GPU path gathers from the full-length vector with
// only its share of the entries, and the slab walk produces its outputs (and with them the remote stores of the
// exchange) only in the last slab phase: measured at 8 GPUs (profiles/r2_multi_gpu.md), 5 entries per sector, 202 us per
// iteration with slabs vs 173 us single-pass although the slab kernel alone is the faster one (119 vs 146 us). Needs every row sorted by column (checked on the device), otherwise the split is dropped.
static int32_t build_slabs(sb200_matrix *m) {
    m->nslabs = 0;
    if (m->tile_cfg >= 0 || m->nrows == 0 || m->nnz == 0) return SB200_OK;
    const char *e = getenv("SUBLINEAR_B200_SLABS");
    const int force = e ? atoi(e) : -1;
    if (force == 0 || force == 1) return SB200_OK;
    // power-law graphs (hub rows present): the slab walk repeats the load imbalance of the few heavy row blocks once per
    // slab — measured on the C3 PageRank system 1.86 ms per push against 1.30 ms single-pass — so the split is only taken
    // when forced (the hub-row marks of the slab layout are exercised by the layout tests)
    if (force < 2) {
        const uint32_t *rp = m->h_row_ptr.data();
        for (uint64_t r = 0; r < m->nrows; r++)
            if (rp[r + 1] - rp[r] > kLongRow) return SB200_OK;
    }
    int S = 0;
    const double vec_bytes = 8.0 * (double)m->ncols;
    if (force >= 2) {
        S = force > kMaxSlabs ? kMaxSlabs : force;
        if ((uint64_t)S > m->ncols) return SB200_OK;
    } else {
        if (vec_bytes <= 48e6 || vec_bytes > kMaxSlabs * 28e6) return SB200_OK;
        const char *d = getenv("SUBLINEAR_B200_SLAB_MIN_DENSITY");
        const double min_density = d ? atof(d) : 8.0;
        if ((double)m->nnz < min_density * vec_bytes / 32.0) return SB200_OK;
        S = (int)std::ceil(vec_bytes / 28e6);
        if (S > kMaxSlabs) S = kMaxSlabs;
    }
    const uint32_t width = (uint32_t)((m->ncols + S - 1) / S);
    const uint64_t n = m->nrows, nblocks = (n + 31) / 32, nb1 = nblocks + 1;
    const uint64_t len_stride = (n + 127) & ~127ull;  // >= 32 * nblocks: every lane of the last block has a slot
    DevBuf<unsigned long long> d_flags;
    SB_TRY(d_flags.alloc(2));
    SB_CUDA(cudaMemsetAsync(d_flags.p, 0, 2 * sizeof(unsigned long long), m->stream));
    SB_TRY(m->d_slab_blk.alloc((size_t)S * nb1));
    SB_TRY(m->d_slab_rel.alloc((size_t)S * len_stride));
    SB_CUDA(cudaMemsetAsync(m->d_slab_rel.p, 0, (size_t)S * len_stride * sizeof(uint16_t), m->stream));
    SB_TRY(launch_slab_count(m->d_cols.p, m->d_row_ptr.p, (uint32_t)n, width, S, kLongRow, m->d_slab_rel.p, len_stride,
                             m->d_slab_blk.p, d_flags.p, m->stream));
    unsigned long long flags[2] = {0, 0};
    SB_CUDA(cudaMemcpyAsync(flags, d_flags.p, sizeof(flags), cudaMemcpyDeviceToHost, m->stream));
    SB_CUDA(cudaStreamSynchronize(m->stream));
    const bool unsorted = flags[0] != 0;
    // banded / block-local matrices gather from a window of the vector that stays cached anyway: extra passes would
    // only add hand-over traffic. Split only when most ENTRIES sit in rows that really reach into several slabs (counting
    // rows instead would drop the split for power-law graphs, where most rows hold little more than their diagonal).
    const bool local = force < 2 && flags[1] * 2 < m->nnz;
    if (unsorted || local) {  // unsorted rows (from_csr input): the split would reorder the sums
        m->d_slab_blk.release();
        m->d_slab_rel.release();
        return SB200_OK;
    }
    // per-slab prefix sums over the block counts, then the slab bases (multiples of 4 entries: 32-byte aligned value loads)
    uint64_t base = 0;
    for (int s = 0; s < S; s++) {
        uint64_t total = 0;
        uint32_t *blk = m->d_slab_blk.p + (size_t)s * nb1;
        SB_TRY(device_exclusive_scan_u32(blk, nblocks, &total, m->stream));
        if (base + total >= 0xFFFFFFF0ull) {
            m->d_slab_blk.release();
            m->d_slab_rel.release();
            return SB200_OK;
        }
        SB_TRY(launch_add_u32(blk, nb1, (uint32_t)base, m->stream));
        base = (base + total + 3) & ~3ull;
    }
    const size_t entries = base + 8;  // over-read slack of the last chunk
    SB_TRY(m->d_slab_cols.alloc(entries));
    SB_TRY(m->d_slab_vals.alloc(entries));
    SB_CUDA(cudaMemsetAsync(m->d_slab_cols.p, 0, entries * sizeof(uint32_t), m->stream));  // padding: column 0, value 0
    SB_CUDA(cudaMemsetAsync(m->d_slab_vals.p, 0, entries * sizeof(double), m->stream));
    SB_TRY(launch_slab_fill(m->d_vals.p, m->d_cols.p, m->d_row_ptr.p, (uint32_t)n, width, S, m->d_slab_rel.p, len_stride,
                            m->d_slab_blk.p, m->d_slab_cols.p, m->d_slab_vals.p, m->stream));
    SB_CUDA(cudaStreamSynchronize(m->stream));
    m->nslabs = S;
    m->slab_width = width;
    m->slab_rel_stride = len_stride;
    m->slab_entries = entries;
    return SB200_OK;
}

// Upload a validated host CSR. Exactly one of row_ptr64 / row_ptr32 is non-null.
int32_t matrix_from_host_csr(const uint64_t *row_ptr64, const uint32_t *row_ptr32, const uint32_t *cols,
                             const double *vals, uint64_t nrows, uint64_t ncols, uint64_t nnz, bool validate,
                             sb200_matrix **out, bool allow_slabs) {
    if (!out) return fail(SB200_ERR_INVALID_INPUT, "out is null");
    *out = nullptr;
    if (nrows >= 0xFFFFFFF0ull || ncols >= 0xFFFFFFF0ull)
        return fail(SB200_ERR_INVALID_INPUT, "dimension exceeds the u32 IndexType of the reference (src/types.rs:22)");
    if (nnz >= 0xFFFFFFF0ull)
        return fail(SB200_ERR_MEMORY_ALLOCATION,
                    "nnz %llu does not fit the u32 row_ptr of CSRStorage (src/matrix/sparse.rs:22); "
                    "row-partition the system across GPUs",
                    (unsigned long long)nnz);
    if ((nnz && (!cols || !vals)) || (!row_ptr64 && !row_ptr32))
        return fail(SB200_ERR_INVALID_INPUT, "null CSR slice");

    std::unique_ptr<sb200_matrix> m(new sb200_matrix());
    m->device = current_device();
    m->nrows = nrows;
    m->ncols = ncols;
    m->nnz = nnz;
    m->n_global = ncols;
    m->h_row_ptr.resize(nrows + 1);
    if (row_ptr64) {
        for (uint64_t i = 0; i <= nrows; i++) {
            if (row_ptr64[i] > nnz) return fail(SB200_ERR_INVALID_SPARSE_MATRIX, "row_ptr[%llu] exceeds nnz", (unsigned long long)i);
            m->h_row_ptr[i] = (uint32_t)row_ptr64[i];
        }
    } else {
        memcpy(m->h_row_ptr.data(), row_ptr32, (nrows + 1) * sizeof(uint32_t));
    }
    const uint32_t *rp = m->h_row_ptr.data();
    if (validate) {
        if (rp[0] != 0 || rp[nrows] != nnz)
            return fail(SB200_ERR_INVALID_SPARSE_MATRIX, "row_ptr must start at 0 and end at nnz");
        long long bad_rp = -1, bad_col = -1, bad_val = -1;
#pragma omp parallel for reduction(max : bad_rp)
        for (long long i = 0; i < (long long)nrows; i++)
            if (rp[i] > rp[i + 1]) bad_rp = std::max(bad_rp, i);
        if (bad_rp >= 0) return fail(SB200_ERR_INVALID_SPARSE_MATRIX, "row_ptr decreases at row %lld", bad_rp);
#pragma omp parallel for reduction(max : bad_col, bad_val)
        for (long long k = 0; k < (long long)nnz; k++) {
            if (cols[k] >= ncols) bad_col = std::max(bad_col, k);
            if (!std::isfinite(vals[k])) bad_val = std::max(bad_val, k);
        }
        if (bad_col >= 0)  // IndexOutOfBounds, as SparseMatrix::from_triplets (src/matrix/mod.rs:174-180)
            return fail(SB200_ERR_INDEX_OUT_OF_BOUNDS, "column index %u out of bounds (max %llu) at entry %lld",
                        cols[bad_col], (unsigned long long)(ncols ? ncols - 1 : 0), bad_col);
        if (bad_val >= 0)  // InvalidInput (src/matrix/mod.rs:181-186)
            return fail(SB200_ERR_INVALID_INPUT, "non-finite value at entry %lld", bad_val);
    }

    SB_TRY(require_device(m->device));  // after host-side validation: input errors do not need a GPU to be reported
    // device arrays are padded so that the 16-byte-granular bulk copies may over-read past nnz
    const size_t nnz_pad = ((nnz + 3) & ~(size_t)3) + 8;
    SB_TRY(m->d_vals.alloc(nnz_pad));
    SB_TRY(m->d_cols.alloc(nnz_pad));
    SB_TRY(m->d_row_ptr.alloc(nrows + 1));
    SB_CUDA(cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking));
    SB_CUDA(cudaMemsetAsync(m->d_vals.p + nnz, 0, (nnz_pad - nnz) * sizeof(double), m->stream));
    SB_CUDA(cudaMemsetAsync(m->d_cols.p + nnz, 0, (nnz_pad - nnz) * sizeof(uint32_t), m->stream));
    SB_TRY(copy_h2d(m->d_vals.p, vals, nnz * sizeof(double), m->stream));
    SB_TRY(copy_h2d(m->d_cols.p, cols, nnz * sizeof(uint32_t), m->stream));
    SB_TRY(copy_h2d(m->d_row_ptr.p, rp, (nrows + 1) * sizeof(uint32_t), m->stream));
    SB_TRY(matrix_finish(m.get(), allow_slabs));
    *out = m.release();
    return SB200_OK;
}

// derived layouts of a handle whose CSR slices are resident (d_vals / d_cols padded, d_row_ptr, h_row_ptr, stream set)
int32_t matrix_finish(sb200_matrix *m, bool allow_slabs) {
    const uint32_t *rp = m->h_row_ptr.data();
    m->tile_cfg = default_tile_cfg();
    std::vector<TileDesc> tiles;
    build_tiles(rp, m->nrows, kTileCfgs[m->tile_cfg < 0 ? 0 : m->tile_cfg], tiles);  // unused by the warp-stream kernel
    m->ntiles = (uint32_t)(tiles.size() - 1);
    SB_TRY(m->d_tiles.alloc(tiles.size()));
    SB_TRY(copy_h2d(m->d_tiles.p, tiles.data(), tiles.size() * sizeof(TileDesc), m->stream));
    if (allow_slabs) SB_TRY(build_slabs(m));
    SB_TRY(build_long_rows(m));
    if (m->nslabs == 0) SB_TRY(build_sell(m));
    SB_CUDA(cudaStreamSynchronize(m->stream));
    return SB200_OK;
}

// a handle over CSR slices that already live on the device (csrc/ingest.cu): takes ownership of nothing, copies the
// slices into padded arrays of its own
int32_t matrix_from_device_csr(const uint32_t *d_row_ptr, const uint32_t *d_cols, const double *d_vals, uint64_t nrows,
                               uint64_t ncols, uint64_t nnz, cudaStream_t src_stream, sb200_matrix **out) {
    if (!out) return fail(SB200_ERR_INVALID_INPUT, "out is null");
    *out = nullptr;
    if (nrows >= 0xFFFFFFF0ull || ncols >= 0xFFFFFFF0ull)
        return fail(SB200_ERR_INVALID_INPUT, "dimension exceeds the u32 IndexType of the reference (src/types.rs:22)");
    if (nnz >= 0xFFFFFFF0ull)
        return fail(SB200_ERR_MEMORY_ALLOCATION,
                    "nnz %llu does not fit the u32 row_ptr of CSRStorage (src/matrix/sparse.rs:22); "
                    "row-partition the system across GPUs",
                    (unsigned long long)nnz);
    std::unique_ptr<sb200_matrix> m(new sb200_matrix());
    m->device = current_device();
    m->nrows = nrows;
    m->ncols = ncols;
    m->nnz = nnz;
    m->n_global = ncols;
    m->h_row_ptr.resize(nrows + 1);
    const size_t nnz_pad = ((nnz + 3) & ~(size_t)3) + 8;
    SB_TRY(m->d_vals.alloc(nnz_pad));
    SB_TRY(m->d_cols.alloc(nnz_pad));
    SB_TRY(m->d_row_ptr.alloc(nrows + 1));
    SB_CUDA(cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking));
    SB_CUDA(cudaStreamSynchronize(src_stream));  // the producer's work on the source slices
    SB_CUDA(cudaMemsetAsync(m->d_vals.p + nnz, 0, (nnz_pad - nnz) * sizeof(double), m->stream));
    SB_CUDA(cudaMemsetAsync(m->d_cols.p + nnz, 0, (nnz_pad - nnz) * sizeof(uint32_t), m->stream));
    SB_CUDA(cudaMemcpyAsync(m->d_vals.p, d_vals, nnz * sizeof(double), cudaMemcpyDeviceToDevice, m->stream));
    SB_CUDA(cudaMemcpyAsync(m->d_cols.p, d_cols, nnz * sizeof(uint32_t), cudaMemcpyDeviceToDevice, m->stream));
    SB_CUDA(cudaMemcpyAsync(m->d_row_ptr.p, d_row_ptr, (nrows + 1) * sizeof(uint32_t), cudaMemcpyDeviceToDevice, m->stream));
    SB_TRY(copy_d2h(m->h_row_ptr.data(), d_row_ptr, (nrows + 1) * sizeof(uint32_t), m->stream));
    SB_CUDA(cudaStreamSynchronize(m->stream));
    SB_TRY(matrix_finish(m.get(), true));
    *out = m.release();
    return SB200_OK;
}

void fill_tile_args(const sb200_matrix *m, TileKernelArgs &a) {
    a.vals = m->d_vals.p;
    a.cols = m->d_cols.p;
    a.row_ptr = m->d_row_ptr.p;
    a.tiles = m->d_tiles.p;
    a.ntiles = m->ntiles;
    a.sell_ptr = m->use_sell ? m->d_sell_ptr.p : nullptr;
    a.sell_cols = m->d_sell_cols.p;
    a.sell_vals = m->d_sell_vals.p;
    a.nlong = m->nlong;
    a.nlong_chunks = m->nlong_chunks;
    a.long_rows = m->d_long_rows.p;
    a.long_first = m->d_long_first.p;
    a.long_chunks = m->d_long_chunks.p;
    a.long_sum = nullptr;
    a.nslabs = m->nslabs;
    a.slab_vals = m->d_slab_vals.p;
    a.slab_cols = m->d_slab_cols.p;
    a.slab_blk = m->d_slab_blk.p;
    a.slab_rel = m->d_slab_rel.p;
    a.slab_rel_stride = m->slab_rel_stride;
    a.acc_keep = m->nrows * 16 <= (24ull << 20);  // carried sums small enough to live in L2 next to the slab window
    a.nrows = (uint32_t)m->nrows;
    a.row_base = (uint32_t)m->row_base;
    a.xin_len = m->ncols;
}

int32_t matrix_spmv_dev(const sb200_matrix *m, const double *x_dev, double *y_dev, int accumulate, cudaStream_t st) {
    TileKernelArgs a{};
    fill_tile_args(m, a);
    a.xin = x_dev;
    a.xin_own = x_dev;
    a.out = y_dev;
    a.accumulate = accumulate;
    return launch_tile_kernel(m->tile_cfg, EPI_SPMV, a, st);
}

// K4, cached per mode. mode 1 (ref_compat) extracts the diagonal like CSRStorage::get, mode 0 sums duplicates.
int32_t matrix_analyse(sb200_matrix *m, int mode, bool need_cols, const ColReduce &reduce_cols) {
    std::lock_guard<std::mutex> lk(m->mu);
    const bool have = m->analysed[mode] && (!need_cols || m->col_analysed);
    if (have) return SB200_OK;
    const uint64_t n = m->nrows;
    if (!m->d_dinv[mode].p) SB_TRY(m->d_dinv[mode].alloc(n));
    DevBuf<unsigned long long> scal;
    SB_TRY(scal.alloc(4));
    unsigned long long init[4] = {kNone, kNone, kNone, 0x7FF0000000000000ull /* +inf */};
    SB_CUDA(cudaMemcpyAsync(scal.p, init, sizeof(init), cudaMemcpyHostToDevice, m->stream));
    DevBuf<double> col_diag, col_off;
    SetupOut o{};
    o.dinv = m->d_dinv[mode].p;
    o.first_bad_dd = scal.p;
    o.first_bad_diag = scal.p + 1;
    o.min_factor_bits = reinterpret_cast<double *>(scal.p + 3);
    if (need_cols) {
        SB_TRY(col_diag.alloc(m->ncols));
        SB_TRY(col_off.alloc(m->ncols));
        SB_CUDA(cudaMemsetAsync(col_diag.p, 0, std::max<uint64_t>(m->ncols, 1) * 8, m->stream));
        SB_CUDA(cudaMemsetAsync(col_off.p, 0, std::max<uint64_t>(m->ncols, 1) * 8, m->stream));
        o.col_diag = col_diag.p;
        o.col_off = col_off.p;
    }
    SB_TRY(launch_setup_rows(m->d_vals.p, m->d_cols.p, m->d_row_ptr.p, (uint32_t)n, (uint32_t)m->row_base,
                             mode == SB200_MODE_REF_COMPAT, o, m->stream));
    if (need_cols && reduce_cols) SB_TRY(reduce_cols(col_diag.p, col_off.p, m->ncols, m->stream));
    if (need_cols) SB_TRY(launch_col_dominance(col_diag.p, col_off.p, (uint32_t)m->ncols, scal.p + 2, m->stream));
    unsigned long long res[4];
    SB_CUDA(cudaMemcpyAsync(res, scal.p, sizeof(res), cudaMemcpyDeviceToHost, m->stream));
    SB_CUDA(cudaStreamSynchronize(m->stream));
    m->first_bad_dd = res[0];
    m->first_bad_diag[mode] = res[1];
    if (need_cols) {
        m->first_bad_col = res[2];
        m->col_analysed = true;
    }
    double f;
    memcpy(&f, &res[3], 8);
    m->has_factor = std::isfinite(f);
    m->min_factor = f;
    m->analysed[mode] = true;
    return SB200_OK;
}

}  // namespace sb200

// -------------------------------------------------------------------------------------------------------------
// C ABI
// -------------------------------------------------------------------------------------------------------------
extern "C" {

int32_t sb200_matrix_from_triplets_ex(const uint64_t *rows, const uint64_t *cols, const double *vals,
                                      uint64_t ntriplets, uint64_t nrows, uint64_t ncols, int32_t dup_policy,
                                      sb200_matrix **out) {
    clear_error();
    if (!out) return fail(SB200_ERR_INVALID_INPUT, "out is null");
    *out = nullptr;
    // SparseMatrix::from_triplets (src/matrix/mod.rs:160-199): validated on the host in triplet order, converted on the
    // device (csrc/ingest.cu)
    SB_TRY(validate_triplets(rows, cols, vals, ntriplets, nrows, ncols));
    return matrix_from_triplets_device(rows, cols, vals, ntriplets, nrows, ncols, dup_policy, out);
}

int32_t sb200_matrix_from_triplets(const uint64_t *rows, const uint64_t *cols, const double *vals,
                                   uint64_t ntriplets, uint64_t nrows, uint64_t ncols, sb200_matrix **out) {
    return sb200_matrix_from_triplets_ex(rows, cols, vals, ntriplets, nrows, ncols, SB200_DUP_KEEP, out);
}

int32_t sb200_matrix_from_csr(const uint32_t *row_ptr, const uint32_t *col_indices, const double *values,
                              uint64_t nrows, uint64_t ncols, uint64_t nnz, sb200_matrix **out) {
    clear_error();
    return matrix_from_host_csr(nullptr, row_ptr, col_indices, values, nrows, ncols, nnz, true, out);
}

int32_t sb200_matrix_from_csr64(const uint64_t *row_ptr, const uint32_t *col_indices, const double *values,
                                uint64_t nrows, uint64_t ncols, uint64_t nnz, sb200_matrix **out) {
    clear_error();
    return matrix_from_host_csr(row_ptr, nullptr, col_indices, values, nrows, ncols, nnz, true, out);
}

// SparseMatrix::from_dense (src/matrix/mod.rs:202-223)
int32_t sb200_matrix_from_dense(const double *data, uint64_t nrows, uint64_t ncols, sb200_matrix **out) {
    clear_error();
    if (!out) return fail(SB200_ERR_INVALID_INPUT, "out is null");
    *out = nullptr;
    if (nrows * ncols != 0 && !data) return fail(SB200_ERR_INVALID_INPUT, "data is null");
    std::vector<uint64_t> r, c;
    std::vector<double> v;
    for (uint64_t i = 0; i < nrows * ncols; i++)
        if (data[i] != 0.0) {
            r.push_back(i / ncols);
            c.push_back(i % ncols);
            v.push_back(data[i]);
        }
    return sb200_matrix_from_triplets(r.data(), c.data(), v.data(), v.size(), nrows, ncols, out);
}

// SparseMatrix::identity / diagonal (src/matrix/mod.rs:226-239)
int32_t sb200_matrix_diagonal(const double *diag, uint64_t size, sb200_matrix **out) {
    clear_error();
    if (size && !diag) return fail(SB200_ERR_INVALID_INPUT, "diag is null");
    std::vector<uint64_t> r, c;
    std::vector<double> v;
    for (uint64_t i = 0; i < size; i++)
        if (diag[i] != 0.0) {
            r.push_back(i);
            c.push_back(i);
            v.push_back(diag[i]);
        }
    return sb200_matrix_from_triplets(r.data(), c.data(), v.data(), v.size(), size, size, out);
}

int32_t sb200_matrix_identity(uint64_t size, sb200_matrix **out) {
    std::vector<double> d(size, 1.0);
    return sb200_matrix_diagonal(d.data(), size, out);
}

// The handle is reference counted: a solver state (csrc/state.cu) shares ownership, so callers (and garbage collectors)
// may free the matrix handle and its states in any order.
void sb200_matrix_free(sb200_matrix *m) { sb200::matrix_release(m); }

int32_t sb200_matrix_rows(const sb200_matrix *m, uint64_t *out) {
    if (!m || !out) return fail(SB200_ERR_INVALID_INPUT, "null argument");
    *out = m->nrows;
    return SB200_OK;
}
int32_t sb200_matrix_cols(const sb200_matrix *m, uint64_t *out) {
    if (!m || !out) return fail(SB200_ERR_INVALID_INPUT, "null argument");
    *out = m->ncols;
    return SB200_OK;
}
int32_t sb200_matrix_nnz(const sb200_matrix *m, uint64_t *out) {
    if (!m || !out) return fail(SB200_ERR_INVALID_INPUT, "null argument");
    *out = m->nnz;
    return SB200_OK;
}

int32_t sb200_matrix_storage_info(const sb200_matrix *m, int32_t *layout, uint64_t *slots, uint64_t *device_bytes) {
    if (!m) return fail(SB200_ERR_INVALID_INPUT, "null argument");
    if (layout) *layout = m->nslabs > 1 ? SB200_LAYOUT_CSR_SLABS : (m->use_sell ? SB200_LAYOUT_SELL32 : SB200_LAYOUT_CSR);
    if (slots) *slots = m->use_sell ? m->sell_slabs * 32 : m->nnz;
    if (device_bytes) {
        uint64_t b = m->d_vals.n * 8 + m->d_cols.n * 4 + m->d_row_ptr.n * 4 + m->d_tiles.n * sizeof(TileDesc) +
                     m->d_sell_ptr.n * 4 + m->d_sell_cols.n * 4 + m->d_sell_vals.n * 8 + m->d_dinv[0].n * 8 +
                     m->d_dinv[1].n * 8;
        b += m->d_slab_vals.n * 8 + m->d_slab_cols.n * 4 + m->d_slab_blk.n * 4 + m->d_slab_rel.n * 2;
        *device_bytes = b;
    }
    return SB200_OK;
}

// Matrix::get -> CSRStorage::get (src/matrix/mod.rs:394-404, src/matrix/sparse.rs:142-155)
int32_t sb200_matrix_get(const sb200_matrix *m, uint64_t row, uint64_t col, double *value, int32_t *present) {
    clear_error();
    if (!m || !present) return fail(SB200_ERR_INVALID_INPUT, "null argument");
    *present = 0;
    if (row >= m->nrows || col >= m->ncols) return SB200_OK;
    const uint32_t s = m->h_row_ptr[row], e = m->h_row_ptr[row + 1];
    if (e == s) return SB200_OK;
    DeviceGuard g(m->device);
    std::vector<uint32_t> ci(e - s);
    std::vector<double> cv(e - s);
    SB_CUDA(cudaMemcpy(ci.data(), m->d_cols.p + s, (e - s) * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    SB_CUDA(cudaMemcpy(cv.data(), m->d_vals.p + s, (e - s) * sizeof(double), cudaMemcpyDeviceToHost));
    uint32_t lo = 0, hi = e - s;
    while (lo < hi) {
        const uint32_t mid = lo + (hi - lo) / 2;
        if (ci[mid] == (uint32_t)col) {
            if (value) *value = cv[mid];
            *present = 1;
            return SB200_OK;
        }
        if (ci[mid] < (uint32_t)col) lo = mid + 1; else hi = mid;
    }
    return SB200_OK;
}

int32_t sb200_matrix_is_diagonally_dominant(const sb200_matrix *m, int32_t dominance, int32_t *out) {
    clear_error();
    if (!m || !out) return fail(SB200_ERR_INVALID_INPUT, "null argument");
    sb200_matrix *mm = const_cast<sb200_matrix *>(m);
    DeviceGuard g(m->device);
    SB_TRY(matrix_analyse(mm, SB200_MODE_CORRECT, dominance == SB200_DOMINANCE_ROW_OR_COL));
    *out = (m->first_bad_dd == kNone) || (dominance == SB200_DOMINANCE_ROW_OR_COL && m->first_bad_col == kNone);
    return SB200_OK;
}

int32_t sb200_matrix_diagonal_dominance_factor(const sb200_matrix *m, double *factor, int32_t *present) {
    clear_error();
    if (!m || !present) return fail(SB200_ERR_INVALID_INPUT, "null argument");
    sb200_matrix *mm = const_cast<sb200_matrix *>(m);
    DeviceGuard g(m->device);
    SB_TRY(matrix_analyse(mm, SB200_MODE_CORRECT, false));
    *present = m->has_factor;
    if (factor) *factor = m->has_factor ? m->min_factor : 0.0;
    return SB200_OK;
}

int32_t sb200_matrix_export_csr(const sb200_matrix *m, uint64_t *row_ptr, uint32_t *col_indices, double *values) {
    clear_error();
    if (!m) return fail(SB200_ERR_INVALID_INPUT, "null matrix");
    DeviceGuard g(m->device);
    if (row_ptr)
        for (uint64_t i = 0; i <= m->nrows; i++) row_ptr[i] = m->h_row_ptr[i];
    if (col_indices) SB_TRY(copy_d2h(col_indices, m->d_cols.p, m->nnz * sizeof(uint32_t), m->stream));
    if (values) SB_TRY(copy_d2h(values, m->d_vals.p, m->nnz * sizeof(double), m->stream));
    SB_CUDA(cudaStreamSynchronize(m->stream));
    return SB200_OK;
}

// Matrix::multiply_vector / multiply_vector_add (src/matrix/mod.rs:415-465): dimension checks, then the kernel.
static int32_t multiply_host(const sb200_matrix *m, const double *x, uint64_t xlen, double *y, uint64_t ylen, int add) {
    clear_error();
    if (!m) return fail(SB200_ERR_INVALID_INPUT, "null matrix");
    if (xlen != m->ncols)
        return fail(SB200_ERR_DIMENSION_MISMATCH, "expected %llu, actual %llu in %s", (unsigned long long)m->ncols,
                    (unsigned long long)xlen, add ? "matrix_vector_multiply_add" : "matrix_vector_multiply");
    if (ylen != m->nrows)
        return fail(SB200_ERR_DIMENSION_MISMATCH, "expected %llu, actual %llu in %s", (unsigned long long)m->nrows,
                    (unsigned long long)ylen, add ? "matrix_vector_multiply_add" : "matrix_vector_multiply");
    if ((xlen && !x) || (ylen && !y)) return fail(SB200_ERR_INVALID_INPUT, "null vector");
    DeviceGuard g(m->device);
    DevBuf<double> dx, dy;
    SB_TRY(dx.alloc(xlen));
    SB_TRY(dy.alloc(ylen));
    SB_TRY(copy_h2d(dx.p, x, xlen * 8, m->stream));
    if (add) SB_TRY(copy_h2d(dy.p, y, ylen * 8, m->stream));
    SB_TRY(matrix_spmv_dev(m, dx.p, dy.p, add, m->stream));
    SB_TRY(copy_d2h(y, dy.p, ylen * 8, m->stream));
    SB_CUDA(cudaStreamSynchronize(m->stream));
    return SB200_OK;
}

int32_t sb200_matrix_multiply_vector(const sb200_matrix *m, const double *x, uint64_t xlen, double *y, uint64_t ylen) {
    return multiply_host(m, x, xlen, y, ylen, 0);
}
int32_t sb200_matrix_multiply_vector_add(const sb200_matrix *m, const double *x, uint64_t xlen, double *y, uint64_t ylen) {
    return multiply_host(m, x, xlen, y, ylen, 1);
}

int32_t sb200_matrix_multiply_vector_dev(const sb200_matrix *m, const double *x_dev, uint64_t xlen, double *y_dev,
                                         uint64_t ylen, int32_t accumulate, void *stream) {
    clear_error();
    if (!m) return fail(SB200_ERR_INVALID_INPUT, "null matrix");
    if (xlen != m->ncols || ylen != m->nrows)
        return fail(SB200_ERR_DIMENSION_MISMATCH, "expected %llu x %llu, actual %llu x %llu in matrix_vector_multiply",
                    (unsigned long long)m->nrows, (unsigned long long)m->ncols, (unsigned long long)ylen,
                    (unsigned long long)xlen);
    DeviceGuard g(m->device);
    return matrix_spmv_dev(m, x_dev, y_dev, accumulate, (cudaStream_t)stream);
}

// SparseMatrix::scale (src/matrix/mod.rs:345-352)
int32_t sb200_matrix_scale(sb200_matrix *m, double factor) {
    clear_error();
    if (!m) return fail(SB200_ERR_INVALID_INPUT, "null matrix");
    DeviceGuard g(m->device);
    SB_TRY(launch_scale(m->d_vals.p, m->nnz, factor, m->stream));
    if (m->use_sell) SB_TRY(launch_scale(m->d_sell_vals.p, m->sell_slabs * 32, factor, m->stream));
    if (m->nslabs > 1) SB_TRY(launch_scale(m->d_slab_vals.p, m->slab_entries, factor, m->stream));
    SB_CUDA(cudaStreamSynchronize(m->stream));
    std::lock_guard<std::mutex> lk(m->mu);
    m->analysed[0] = m->analysed[1] = m->col_analysed = false;  // cached D^-1 is stale
    return SB200_OK;
}

// create_test_matrix + create_test_rhs (benches/performance_benchmarks.rs:12-43), emitted directly as the CSR
// that from_triplets would build (stable by-column order inside the row, exact zeros dropped).
int32_t sb200_gen_bench_csr(uint64_t size, double sparsity, uint64_t row0, uint64_t row1, uint64_t *row_ptr,
                            uint32_t *col_indices, double *values, double *b, uint64_t *nnz_out) {
    clear_error();
    if (row1 > size || row0 > row1) return fail(SB200_ERR_INVALID_INPUT, "bad row range");
    const uint64_t k = std::min<uint64_t>((uint64_t)std::max((double)size * sparsity, 3.0), size);
    const uint64_t nloc = row1 - row0;
    const double two64 = 18446744073709551616.0;  // `u64::MAX as f64`
    auto gen_row = [&](uint64_t i, uint32_t *c, double *v) -> uint64_t {
        const double diag = 10.0 + ((double)i * 0.01);
        uint64_t cnt = 0;
        c[cnt] = (uint32_t)i;
        v[cnt++] = diag;
        const double max_off = diag / ((double)k * 2.0);
        uint64_t rng = i * 1664525ull + 1013904223ull;
        for (uint64_t j = 1; j < k; j++) {
            rng = rng * 1664525ull + 1013904223ull;
            const uint64_t col = rng % size;
            if (col != i) {
                rng = rng * 1664525ull + 1013904223ull;
                c[cnt] = (uint32_t)col;
                v[cnt++] = ((double)rng / two64) * max_off;
            }
        }
        for (uint64_t a = 1; a < cnt; a++) {  // stable insertion sort by column
            const uint32_t cc = c[a];
            const double vv = v[a];
            uint64_t p = a;
            while (p > 0 && c[p - 1] > cc) { c[p] = c[p - 1]; v[p] = v[p - 1]; p--; }
            c[p] = cc;
            v[p] = vv;
        }
        uint64_t w = 0;
        for (uint64_t a = 0; a < cnt; a++)
            if (v[a] != 0.0) { c[w] = c[a]; v[w] = v[a]; w++; }
        return w;
    };
    // pass 1: row lengths (needed for a parallel fill); pass 2: fill
    std::vector<uint64_t> rp(nloc + 1, 0);
#pragma omp parallel
    {
        std::vector<uint32_t> c(k + 1);
        std::vector<double> v(k + 1);
#pragma omp for schedule(static)
        for (long long r = 0; r < (long long)nloc; r++) rp[r + 1] = gen_row(row0 + r, c.data(), v.data());
    }
    for (uint64_t r = 0; r < nloc; r++) rp[r + 1] += rp[r];
    if (nnz_out) *nnz_out = rp[nloc];
    if (row_ptr) memcpy(row_ptr, rp.data(), (nloc + 1) * sizeof(uint64_t));
    if (col_indices && values) {
#pragma omp parallel
        {
            std::vector<uint32_t> c(k + 1);
            std::vector<double> v(k + 1);
#pragma omp for schedule(static)
            for (long long r = 0; r < (long long)nloc; r++) {
                const uint64_t cnt = gen_row(row0 + r, c.data(), v.data());
                memcpy(col_indices + rp[r], c.data(), cnt * sizeof(uint32_t));
                memcpy(values + rp[r], v.data(), cnt * sizeof(double));
            }
        }
    }
    if (b)
        for (uint64_t r = 0; r < nloc; r++) b[r] = 1.0 + ((double)(row0 + r) * 0.001);
    return SB200_OK;
}

}  // extern "C"
