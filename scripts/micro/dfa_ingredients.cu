// Microbenchmark 2: start from the pure dependent-LDS chain (ILP chains per thread) and add the
// ingredients of the real scan one at a time, to see which one costs the throughput.
//   F_BRANCH : per step, branch on min(entries) < bound (never taken)
//   F_TEXT   : classes come from text bytes held in registers (PRMT + min/sub), text loaded with
//              strided 16-byte global loads (one 512-byte slice per chain), prefetched one group ahead
//   F_TEXTSM : like F_TEXT but the class bytes come from a linear congruential generator (no loads)
#include <cstdio>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ uint4 ldt(const uint8_t *p)
{
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}

template <int ILP, bool F_BRANCH, int F_TEXT>
__global__ void k(const uint16_t *gtab, uint32_t entries, uint32_t ncls, uint32_t nstates, const uint8_t *text,
                  uint32_t slice, uint32_t nslices, unsigned long long *cycles, uint32_t *sink)
{
    extern __shared__ __align__(16) unsigned char raw[];
    uint16_t *tab = (uint16_t *)raw;
    for (uint32_t i = threadIdx.x; i < entries; i += blockDim.x) tab[i] = gtab[i];
    __syncthreads();
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(tab);
    const uint32_t row = ncls * 2;
    uint32_t s[ILP];
    uint32_t pos[ILP];
    const uint32_t lane = threadIdx.x & 31, warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    for (int q = 0; q < ILP; ++q) {
        s[q] = (threadIdx.x * 2654435761u + blockIdx.x * 40503u + q * 977u) % nstates;
        pos[q] = ((warp * ILP + q) * 32 + lane) % nslices * slice;
    }
    uint32_t x = threadIdx.x * 747796405u + blockIdx.x * 2891336453u + 1u;
    uint32_t hits = 0;
    unsigned long long t0 = clock64();
    const int groups = slice / 16;
    uint4 cur[ILP], nxt[ILP];
    if (F_TEXT == 1) for (int q = 0; q < ILP; ++q) cur[q] = ldt(text + pos[q]);
    for (int g = 0; g < groups; ++g) {
        if (F_TEXT == 1) {
            for (int q = 0; q < ILP; ++q) { nxt[q] = cur[q]; if (g + 1 < groups) nxt[q] = ldt(text + pos[q] + 16 * (g + 1)); }
        } else {
            for (int q = 0; q < ILP; ++q) { x = x * 1664525u + 1013904223u; cur[q] = make_uint4(x, x * 3u, x * 5u, x * 7u); }
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            uint32_t e[ILP];
#pragma unroll
            for (int q = 0; q < ILP; ++q) {
                const uint32_t w = (j < 4) ? cur[q].x : (j < 8) ? cur[q].y : (j < 12) ? cur[q].z : cur[q].w;
                uint32_t b = __byte_perm(w, 0, 0x4440 | (j & 3));
                uint32_t c = (F_TEXT == 1) ? min(b - 97u, ncls - 1) : (b % ncls);
                uint32_t t = base + c * 2;
                uint32_t addr;
                asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(addr) : "r"(s[q]), "r"(row), "r"(t));
                asm volatile("ld.shared.u16 %0, [%1];" : "=r"(e[q]) : "r"(addr));
            }
            if (F_BRANCH) {
                uint32_t m = e[0];
#pragma unroll
                for (int q = 1; q < ILP; ++q) m = min(m, e[q]);
                if (m >= nstates) { hits++; }
            }
#pragma unroll
            for (int q = 0; q < ILP; ++q) s[q] = e[q];
        }
        if (F_TEXT == 1) for (int q = 0; q < ILP; ++q) cur[q] = nxt[q];
    }
    unsigned long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    uint32_t acc = hits;
    for (int q = 0; q < ILP; ++q) acc ^= s[q];
    if (acc == 0xffffffffu) sink[0] = acc;
}

template <int ILP, bool F_BRANCH, int F_TEXT>
void run(int threads, uint32_t slice)
{
    const int nstates = 16000, ncls = 7;
    const uint32_t entries = nstates * ncls;
    std::vector<uint16_t> h(entries);
    uint32_t x = 12345;
    for (uint32_t i = 0; i < entries; ++i) { x = x * 1103515245u + 12345u; h[i] = (uint16_t)((x >> 8) % nstates); }
    const uint32_t nslices = 148u * threads * ILP;      // every chain its own slice
    const size_t tbytes = (size_t)nslices * slice;
    std::vector<uint8_t> ht(tbytes);
    for (size_t i = 0; i < tbytes; ++i) { x = x * 1103515245u + 12345u; ht[i] = 97 + (x >> 16) % 6; }
    uint16_t *d; uint8_t *dt; unsigned long long *cyc; uint32_t *sink;
    cudaMalloc(&d, entries * 2); cudaMalloc(&dt, tbytes + 64); cudaMalloc(&cyc, 148 * 8); cudaMalloc(&sink, 4);
    cudaMemcpy(d, h.data(), entries * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(dt, ht.data(), tbytes, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(k<ILP, F_BRANCH, F_TEXT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 230000);
    k<ILP, F_BRANCH, F_TEXT><<<148, threads, entries * 2>>>(d, entries, ncls, nstates, dt, slice, nslices, cyc, sink);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    k<ILP, F_BRANCH, F_TEXT><<<148, threads, entries * 2>>>(d, entries, ncls, nstates, dt, slice, nslices, cyc, sink);
    cudaEventRecord(b); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, a, b);
    double steps = 148.0 * threads * ILP * slice;
    printf("ILP=%d threads=%4d branch=%d text=%d slice=%4u : %.2f Tsteps/s (%.2f steps/cyc/SM @1.965GHz)  %s\n", ILP, threads,
           (int)F_BRANCH, F_TEXT, slice, steps / (ms * 1e-3) / 1e12, steps / (ms * 1e-3) / 148 / 1.965e9,
           cudaGetErrorString(cudaGetLastError()));
    cudaFree(d); cudaFree(dt); cudaFree(cyc); cudaFree(sink);
}

int main()
{
    run<4, false, 0>(512, 4096);
    run<4, true, 0>(512, 4096);
    run<4, false, 1>(512, 4096);
    run<4, true, 1>(512, 4096);
    run<4, true, 1>(512, 512);
    run<2, true, 1>(1024, 512);
    run<1, true, 1>(1024, 512);
    run<8, true, 1>(256, 512);
    run<4, true, 1>(256, 512);
    return 0;
}
