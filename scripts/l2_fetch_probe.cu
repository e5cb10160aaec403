// l2_fetch_probe.cu -- measurement aid (not part of the library): random 32-byte gathers from a table much larger
// than L2 through several load forms.  Run under ncu to read DRAM bytes / L2 sectors per load:
//   ncu --metrics dram__bytes_read.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_requests_srcunit_tex_op_read.sum,gpu__time_duration.sum ./l2_fetch_probe [granularity]
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t mix64(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
    return x;
}
template <int MODE>
__device__ __forceinline__ uint64_t load32(const char* p) {
    uint64_t a, b, c, d;
    if (MODE == 0) {
        uint64_t pol;
        asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
        asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.u64 {%0,%1,%2,%3}, [%4], %5;" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p), "l"(pol));
    } else if (MODE == 1) {
        asm volatile("ld.global.nc.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
    } else if (MODE == 2) {
        asm volatile("ld.global.nc.v2.u64 {%0,%1}, [%2];" : "=l"(a), "=l"(b) : "l"(p));
        asm volatile("ld.global.nc.v2.u64 {%0,%1}, [%2];" : "=l"(c), "=l"(d) : "l"(p + 16));
    } else if (MODE == 3) {
        asm volatile("ld.global.nc.u64 %0, [%1];" : "=l"(a) : "l"(p));
        asm volatile("ld.global.nc.u64 %0, [%1];" : "=l"(b) : "l"(p + 8));
        asm volatile("ld.global.nc.u64 %0, [%1];" : "=l"(c) : "l"(p + 16));
        asm volatile("ld.global.nc.u64 %0, [%1];" : "=l"(d) : "l"(p + 24));
    } else if (MODE == 4) {
        uint64_t pol;
        asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
        asm volatile("ld.global.nc.L2::cache_hint.v4.u64 {%0,%1,%2,%3}, [%4], %5;" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p), "l"(pol));
    } else if (MODE == 5) {
        asm volatile("ld.global.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
    } else if (MODE == 6) {
        asm volatile("ld.global.nc.L1::no_allocate.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
    } else if (MODE == 7) {   // 128-bit halves, evict-first, no L1 allocation
        uint64_t pol;
        asm("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
        asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.u64 {%0,%1}, [%2], %3;" : "=l"(a), "=l"(b) : "l"(p), "l"(pol));
        asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v2.u64 {%0,%1}, [%2], %3;" : "=l"(c), "=l"(d) : "l"(p + 16), "l"(pol));
    } else {   // 8: one 8-byte word only (what a plain u64 gather costs)
        asm volatile("ld.global.nc.u64 %0, [%1];" : "=l"(a) : "l"(p));
        b = c = d = 0;
    }
    return a ^ b ^ c ^ d;
}
template <int MODE>
__global__ void __launch_bounds__(256) k_gather(const char* table, uint64_t n_units, uint64_t* out, int iters) {
    const uint64_t tid = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    uint64_t acc = 0;
#pragma unroll 1
    for (int i = 0; i < iters; i++) {
        const uint64_t u = mix64(tid * 1315423911ULL + i) % n_units;
        acc ^= load32<MODE>(table + 32 * u);
    }
    if (acc == 0x1234567) out[0] = acc;
}
template <int MODE>
static void run(const char* name, const char* table, uint64_t n_units, uint64_t* out) {
    const int iters = 8, grid = 148 * 32, block = 256;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k_gather<MODE><<<grid, block>>>(table, n_units, out, iters);
    cudaEventRecord(e0);
    k_gather<MODE><<<grid, block>>>(table, n_units, out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double loads = (double)grid * block * iters;
    printf("mode %d %-44s %.3f ms  %.1f G loads/s  %.0f GB/s useful (32 B each)\n", MODE, name, ms, loads / ms / 1e6, loads * 32 / ms / 1e6);
}
int main(int argc, char** argv) {
    if (argc > 1) {
        cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, (size_t)atoi(argv[1]));
        size_t g = 0;
        cudaDeviceGetLimit(&g, cudaLimitMaxL2FetchGranularity);
        printf("cudaLimitMaxL2FetchGranularity set %s -> %zu (%s)\n", argv[1], g, cudaGetErrorString(e));
    } else {
        size_t g = 0;
        cudaDeviceGetLimit(&g, cudaLimitMaxL2FetchGranularity);
        printf("cudaLimitMaxL2FetchGranularity default %zu\n", g);
    }
    const uint64_t bytes = 1ull << 30;
    char* table; uint64_t* out;
    cudaMalloc(&table, bytes); cudaMalloc(&out, 8);
    cudaMemset(table, 1, bytes);
    run<0>("v4.u64 nc no_allocate evict_first (dict)", table, bytes / 32, out);
    run<1>("v4.u64 nc", table, bytes / 32, out);
    run<2>("2 x v2.u64 nc", table, bytes / 32, out);
    run<3>("4 x u64 nc", table, bytes / 32, out);
    run<4>("v4.u64 nc evict_last", table, bytes / 32, out);
    run<5>("v4.u64 (not nc)", table, bytes / 32, out);
    run<6>("v4.u64 nc no_allocate", table, bytes / 32, out);
    run<7>("2 x v2.u64 nc no_allocate evict_first", table, bytes / 32, out);
    run<8>("1 x u64 nc (8 of 32 bytes)", table, bytes / 32, out);
    cudaDeviceSynchronize();
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
