// gather_probe.cu — micro-benchmark behind the design of the push kernel's gather stage.
// Measures random 8-byte gathers/s on a B200 as a function of: load flavour, source footprint (L2 residency),
// resident threads per SM, gathers in flight per thread, and the shared-memory carve-out (which shrinks L1).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_probe gather_probe.cu ; run: ./gather_probe
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                                  \
    do {                                                                                       \
        cudaError_t e = (x);                                                                   \
        if (e != cudaSuccess) {                                                                \
            printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__);             \
            exit(1);                                                                           \
        }                                                                                      \
    } while (0)

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

enum Flavour { LDG_DEFAULT = 0, LDG_NOALLOC = 1, LDG_EVICT_LAST = 2, LDG_CG = 3, LDGSTS_CA8 = 4, LDGSTS_CG16 = 5, LDG_NOALLOC_EL = 6 };

template <int FL>
__device__ __forceinline__ double gload(const double *p, uint64_t pol) {
    double v;
    if (FL == LDG_DEFAULT) asm volatile("ld.global.nc.f64 %0, [%1];" : "=d"(v) : "l"(p));
    else if (FL == LDG_NOALLOC) asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
    else if (FL == LDG_EVICT_LAST) asm volatile("ld.global.nc.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol));
    else if (FL == LDG_NOALLOC_EL) asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f64 %0, [%1], %2;" : "=d"(v) : "l"(p), "l"(pol));
    else asm volatile("ld.global.cg.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}

// each thread performs `per_thread` gathers, U in flight at a time; indices are hashed (no index stream)
template <int FL, int U>
__global__ void gather_kernel(const double *__restrict__ src, uint32_t n, uint32_t per_thread, double *out) {
    extern __shared__ __align__(16) unsigned char smem[];
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    const uint32_t gtid = blockIdx.x * blockDim.x + threadIdx.x;
    double acc = 0.0;
    if (FL == LDGSTS_CA8 || FL == LDGSTS_CG16) {
        // per-thread ring of U slots in shared memory
        double *ring = reinterpret_cast<double *>(smem) + (size_t)threadIdx.x * U * (FL == LDGSTS_CG16 ? 2 : 1);
        for (uint32_t it = 0; it < per_thread; it += U) {
#pragma unroll
            for (int u = 0; u < U; u++) {
                const uint32_t idx = hash32(gtid * 0x9E3779B9u + it + u) % n;
                if (FL == LDGSTS_CA8)
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(ring + u)), "l"(src + idx) : "memory");
                else
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(ring + 2 * u)), "l"(src + (idx & ~1u)) : "memory");
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
            asm volatile("cp.async.wait_group 0;" ::: "memory");
#pragma unroll
            for (int u = 0; u < U; u++) acc += ring[(FL == LDGSTS_CG16 ? 2 : 1) * u];
        }
    } else {
        for (uint32_t it = 0; it < per_thread; it += U) {
            double v[U];
#pragma unroll
            for (int u = 0; u < U; u++) {
                const uint32_t idx = hash32(gtid * 0x9E3779B9u + it + u) % n;
                v[u] = gload<FL>(src + idx, pol);
            }
#pragma unroll
            for (int u = 0; u < U; u++) acc += v[u];
        }
    }
    if (acc == 12345.678) out[gtid] = acc;
}

template <int FL, int U>
double run(const double *src, uint32_t n, int threads_per_sm, size_t smem_per_block, int block, double *out, int sms) {
    const int blocks_per_sm = threads_per_sm / block;
    const int grid = sms * blocks_per_sm;
    size_t smem = smem_per_block;
    if (FL == LDGSTS_CA8) smem = std::max(smem, (size_t)block * U * 8);
    if (FL == LDGSTS_CG16) smem = std::max(smem, (size_t)block * U * 16);
    CK(cudaFuncSetAttribute(gather_kernel<FL, U>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, gather_kernel<FL, U>, block, smem));
    if (occ < blocks_per_sm) return -1.0;
    const uint64_t total = 200ull * 1000 * 1000;
    uint32_t per_thread = (uint32_t)(total / ((uint64_t)grid * block));
    per_thread = per_thread / U * U;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    gather_kernel<FL, U><<<grid, block, smem>>>(src, n, per_thread, out);  // warm
    CK(cudaEventRecord(e0));
    gather_kernel<FL, U><<<grid, block, smem>>>(src, n, per_thread, out);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    return (double)per_thread * grid * block / (ms * 1e-3) / 1e9;  // G gathers / s
}

int main() {
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    const size_t max_n = 100ull * 1000 * 1000;  // 800 MB
    double *src, *out;
    CK(cudaMalloc(&src, max_n * 8));
    CK(cudaMemset(src, 0, max_n * 8));
    CK(cudaMalloc(&out, 64ull * 1024 * 1024));
    const char *names[] = {"ldg", "ldg.noalloc", "ldg.evict_last", "ldg.cg", "ldgsts.ca8", "ldgsts.cg16", "ldg.noalloc.el"};
    printf("flavour,src_MB,threads_per_SM,U,smem_KB_per_SM,Ggather_per_s,cycles_per_gather_per_SM@1.965GHz\n");
    const uint32_t sizes[] = {2000000u, 5000000u, 10000000u, 100000000u};  // 16 / 40 / 80 / 800 MB
    for (uint32_t n : sizes) {
        for (int tps : {256, 512, 1024, 2048}) {
            for (int smem_kb_sm : {0, 64, 128, 192}) {
                const int block = 256;
                const int bps = tps / block;
                const size_t smem_blk = (size_t)smem_kb_sm * 1024 / bps;
#define ROW(FL, U)                                                                                              \
    {                                                                                                           \
        double g = run<FL, U>(src, n, tps, smem_blk, block, out, sms);                                          \
        if (g > 0)                                                                                              \
            printf("%s,%u,%d,%d,%d,%.2f,%.2f\n", names[FL], (unsigned)(n / 125000), tps, U, smem_kb_sm, g,     \
                   1.965 * sms / g);                                                                            \
    }
                ROW(LDG_DEFAULT, 4) ROW(LDG_DEFAULT, 12) ROW(LDG_NOALLOC, 4) ROW(LDG_NOALLOC, 12) ROW(LDG_EVICT_LAST, 12)
                ROW(LDG_NOALLOC_EL, 12) ROW(LDG_CG, 12) ROW(LDGSTS_CA8, 12) ROW(LDGSTS_CG16, 12)
                fflush(stdout);
            }
        }
    }
    return 0;
}
