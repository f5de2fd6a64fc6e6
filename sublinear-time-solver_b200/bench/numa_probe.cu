// numa_probe.cu — which L2 half is "near" for which SM, and at what address granularity?
//
// Why: the push kernel on uniform-random columns runs at the chip's L2 sector-throughput cap: 314 M LTS
// sector operations per launch for 150 M sectors actually delivered, because a sector homed on the other die is looked
// up in the near slice (miss), in the far slice (hit) and filled into the near slice again. A die-aware schedule
// (stream each row block from an SM of the die its slices live on, keep one copy of the gather source per die) needs two
// maps that CUDA does not expose: SM -> die and address -> home die. This probe measures both from L2-hit latency
// (B300_MICROARCH.md: 234 cycles near / 262 far): one CTA per SM pointer-chases inside every 2 KB chunk of a buffer
// with L1-bypassing loads and records the mean latency per (SM, chunk).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o numa_probe numa_probe.cu ; run: ./numa_probe out.bin [MB]
// Output: header {nsm, nchunks, chunk_bytes} (3 x u32), smid[nsm] (u32), lat[nsm][nchunks] (u16, cycles per load).
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                          \
    do {                                                                               \
        cudaError_t e = (x);                                                           \
        if (e != cudaSuccess) {                                                        \
            printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__);     \
            exit(1);                                                                   \
        }                                                                              \
    } while (0)

constexpr int kChunkBytes = 2048, kChunkWords = kChunkBytes / 4, kLines = kChunkBytes / 128;

__device__ __forceinline__ uint32_t ld_cg(const uint32_t *p) {
    uint32_t v;
    asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__global__ void init_chase(uint32_t *buf, uint32_t nchunks) {
    // within each chunk: line l -> line (l + 5) % 16 (a 16-cycle), stored as a word offset inside the chunk
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < (uint64_t)nchunks * kChunkWords;
         i += (uint64_t)gridDim.x * blockDim.x) {
        const uint32_t w = (uint32_t)(i % kChunkWords), l = w / 32;
        buf[i] = ((l + 5) % kLines) * 32;
    }
}

__global__ void probe(const uint32_t *buf, uint32_t nchunks, uint16_t *lat, uint32_t *smid_out) {
    extern __shared__ unsigned char force_one_cta_per_sm[];
    if (threadIdx.x != 0) return;
    uint32_t smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    smid_out[blockIdx.x] = smid;
    uint32_t sink = 0;
    for (uint32_t c = 0; c < nchunks; c++) {
        const uint32_t *p = buf + (size_t)c * kChunkWords;
        uint32_t off = 0;
        for (int i = 0; i < kLines; i++) off = ld_cg(p + off);  // warm the 16 lines into L2
        const long long t0 = clock64();
        for (int r = 0; r < 2 * kLines; r++) off = ld_cg(p + off);
        const long long t1 = clock64();
        sink += off;
        lat[(size_t)blockIdx.x * nchunks + c] = (uint16_t)((t1 - t0) / (2 * kLines));
    }
    if (sink == 0xFFFFFFFFu) smid_out[blockIdx.x] = sink;
}

// ---- second experiment: what does a die-local gather cost in L2 sector operations? -------------------------------
// Every thread reads its indices coalesced (like a column-index stream) and gathers 8-byte values, U in flight.
// The CTAs of each die take their indices from that die's index array, in CTA-rank order within the die.
// Run under `ncu --metrics lts__t_sectors.sum,lts__t_sectors_srcunit_ltcfabric.sum,lts__t_sectors_srcunit_tex.sum`.
constexpr int kU = 8;
__global__ void __launch_bounds__(256) gather_by_die(const double *__restrict__ src, const uint32_t *__restrict__ idx_a,
                                                     const uint32_t *__restrict__ idx_b, uint64_t n_idx,
                                                     const uint8_t *__restrict__ die_of_sm, unsigned *rank_ctr,
                                                     unsigned ctas_die0, unsigned ctas_die1, int only_die, double *out) {
    __shared__ unsigned s_rank;
    uint32_t smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    const int die = die_of_sm[smid];
    if (only_die >= 0 && die != only_die) return;  // single-die run: the other die's SMs stay idle
    if (threadIdx.x == 0) s_rank = atomicAdd(&rank_ctr[die], 1u);
    __syncthreads();
    const uint32_t *__restrict__ idx = die ? idx_b : idx_a;
    const unsigned ctas_per_die = die ? ctas_die1 : ctas_die0;  // 8 resident CTAs per SM; the split is 70 / 78 or 74 / 74
    double acc = 0.0;
    for (uint64_t base = ((uint64_t)s_rank * 256 + threadIdx.x) * kU; base + kU <= n_idx;
         base += (uint64_t)ctas_per_die * 256 * kU) {
        uint32_t c[kU];
        double v[kU];
#pragma unroll
        for (int u = 0; u < kU; u++) c[u] = idx[base + u];
#pragma unroll
        for (int u = 0; u < kU; u++) asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(v[u]) : "l"(src + c[u]));
#pragma unroll
        for (int u = 0; u < kU; u++) acc += v[u];
    }
    if (acc == 12345.678) out[blockIdx.x * 256 + threadIdx.x] = acc;
}

static uint64_t splitmix(uint64_t &s) {
    uint64_t z = (s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

// mode: 0 = uniform over the buffer, 1 = only chunks near the gathering die, 2 = only chunks far from it
static void run_gather(const char *name, int mode, const double *src, uint32_t nchunks, const std::vector<uint8_t> &home,
                       const uint8_t *d_die_of_sm, int sms, int sms_die1, uint64_t n_idx, int only_die = -1) {
    std::vector<uint32_t> lists[2];
    for (uint32_t c = 0; c < nchunks; c++) lists[home[c]].push_back(c);
    std::vector<uint32_t> h[2];
    uint64_t seed = 42 + mode;
    for (int die = 0; die < 2; die++) {
        h[die].resize(n_idx);
        const std::vector<uint32_t> &pick = mode == 1 ? lists[die] : lists[die ^ 1];
        for (uint64_t i = 0; i < n_idx; i++) {
            const uint64_t r = splitmix(seed);
            const uint32_t chunk = mode == 0 ? (uint32_t)(r % nchunks) : pick[r % pick.size()];
            h[die][i] = chunk * (kChunkBytes / 8) + (uint32_t)((r >> 40) % (kChunkBytes / 8));
        }
    }
    uint32_t *d_idx[2];
    unsigned *ctr;
    double *out;
    for (int die = 0; die < 2; die++) {
        CK(cudaMalloc(&d_idx[die], n_idx * 4));
        CK(cudaMemcpy(d_idx[die], h[die].data(), n_idx * 4, cudaMemcpyHostToDevice));
    }
    CK(cudaMalloc(&ctr, 8));
    CK(cudaMalloc(&out, (size_t)sms * 8 * 256 * 8));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    float best = 1e9f;
    for (int rep = 0; rep < 3; rep++) {
        CK(cudaMemset(ctr, 0, 8));
        CK(cudaEventRecord(e0));
        gather_by_die<<<sms * 8, 256>>>(src, d_idx[0], d_idx[1], n_idx, d_die_of_sm, ctr, 8u * (sms - sms_die1), 8u * sms_die1, only_die, out);
        CK(cudaEventRecord(e1));
        CK(cudaDeviceSynchronize());
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    printf("gather_by_die %-14s: %.1f M gathers, %.3f ms, %.1f G gathers/s (+ %.1f GB/s index stream)\n", name, (only_die >= 0 ? 1.0 : 2.0) * n_idx / 1e6,
           best, (only_die >= 0 ? 1.0 : 2.0) * n_idx / best / 1e6, (only_die >= 0 ? 1.0 : 2.0) * n_idx * 4 / best / 1e6);
    cudaFree(d_idx[0]); cudaFree(d_idx[1]); cudaFree(ctr); cudaFree(out);
}

int main(int argc, char **argv) {
    const char *path = argc > 1 ? argv[1] : "numa_probe.bin";
    const size_t mb = argc > 2 ? atoi(argv[2]) : 16;
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const uint32_t nchunks = (uint32_t)(mb * 1024 * 1024 / kChunkBytes);
    uint32_t *buf, *smid;
    uint16_t *lat;
    CK(cudaMalloc(&buf, (size_t)nchunks * kChunkBytes));
    CK(cudaMalloc(&lat, (size_t)sms * nchunks * 2));
    CK(cudaMalloc(&smid, sms * 4));
    init_chase<<<1024, 256>>>(buf, nchunks);
    CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    probe<<<sms, 32, 200 * 1024>>>(buf, nchunks, lat, smid);  // 200 KB of shared memory: one CTA per SM
    CK(cudaDeviceSynchronize());
    std::vector<uint16_t> h((size_t)sms * nchunks);
    std::vector<uint32_t> hs(sms);
    CK(cudaMemcpy(h.data(), lat, h.size() * 2, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hs.data(), smid, sms * 4, cudaMemcpyDeviceToHost));
    FILE *f = fopen(path, "wb");
    uint32_t hdr[3] = {(uint32_t)sms, nchunks, (uint32_t)kChunkBytes};
    fwrite(hdr, 4, 3, f);
    fwrite(hs.data(), 4, sms, f);
    fwrite(h.data(), 2, h.size(), f);
    fclose(f);
    double s = 0;
    for (auto v : h) s += v;
    printf("numa_probe: %d SMs x %u chunks of %d B, mean latency %.1f cycles -> %s\n", sms, nchunks, kChunkBytes, s / h.size(), path);

    // ---- classify on the host: SM -> die by the sign of the correlation with CTA 0's centred latency vector over the
    // first chunks, chunk -> home die by which SM group sees it faster ----
    const uint32_t ncls = nchunks < 2048 ? nchunks : 2048;
    std::vector<double> mean(sms, 0.0);
    for (int b = 0; b < sms; b++) {
        for (uint32_t c = 0; c < ncls; c++) mean[b] += h[(size_t)b * nchunks + c];
        mean[b] /= ncls;
    }
    std::vector<uint8_t> die_of_cta(sms, 0), die_of_sm(256, 0);
    int count1 = 0;
    for (int b = 0; b < sms; b++) {
        double dot = 0.0;
        for (uint32_t c = 0; c < ncls; c++)
            dot += (h[(size_t)b * nchunks + c] - mean[b]) * (h[c] - mean[0]);
        die_of_cta[b] = dot < 0.0;
        die_of_sm[hs[b]] = die_of_cta[b];
        count1 += die_of_cta[b];
    }
    std::vector<uint8_t> home(nchunks);
    uint32_t home1 = 0, ambiguous = 0;
    for (uint32_t c = 0; c < nchunks; c++) {
        double l0 = 0.0, l1 = 0.0;
        for (int b = 0; b < sms; b++) (die_of_cta[b] ? l1 : l0) += h[(size_t)b * nchunks + c];
        l0 /= (sms - count1);
        l1 /= count1;
        home[c] = l1 < l0;
        home1 += home[c];
        ambiguous += std::abs(l1 - l0) < 10.0;
    }
    printf("classification: %d / %d SMs per die, %.1f %% of the chunks homed on die 1, %u ambiguous (<10 cycles)\n", sms - count1,
           count1, 100.0 * home1 / nchunks, ambiguous);
    if (count1 < sms / 4 || count1 > 3 * sms / 4) {
        printf("unexpected die split, skipping the gather experiment\n");
        return 0;
    }
    uint8_t *d_die;
    CK(cudaMalloc(&d_die, 256));
    CK(cudaMemcpy(d_die, die_of_sm.data(), 256, cudaMemcpyHostToDevice));
    const uint64_t n_idx = 50ull * 1000 * 1000;  // per die: 100 M gathers per launch, the push kernel's count
    const double *src = reinterpret_cast<const double *>(buf);
    run_gather("uniform", 0, src, nchunks, home, d_die, sms, count1, n_idx);
    run_gather("near", 1, src, nchunks, home, d_die, sms, count1, n_idx);
    run_gather("far", 2, src, nchunks, home, d_die, sms, count1, n_idx);
    // footprint sweep (uniform over the first 1 / 8 / 20 MB of the buffer) and single-die runs
    for (uint32_t mbs : {1u, 8u, 20u}) {
        char nm[32];
        snprintf(nm, sizeof(nm), "uniform_%uMB", mbs);
        if (mbs * 512u <= nchunks) run_gather(nm, 0, src, mbs * 512u, home, d_die, sms, count1, n_idx);
    }
    run_gather("die0_near", 1, src, nchunks, home, d_die, sms, count1, n_idx, 0);
    run_gather("die0_far", 2, src, nchunks, home, d_die, sms, count1, n_idx, 0);
    run_gather("die0_uniform", 0, src, nchunks, home, d_die, sms, count1, n_idx, 0);
    return 0;
}
