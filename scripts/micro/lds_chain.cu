// Microbenchmark: dependent shared-memory lookup chains (the DFA inner loop in isolation).
// Each thread walks  s = tab[s*ncls + c_k]  with pseudo-random classes; reports cycles per step per warp
// and aggregate steps/cycle/SM for several warp counts, table sizes and entry widths.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <vector>

template <typename E, int ILP>
__global__ void chain(const E *gtab, uint32_t entries, uint32_t ncls, uint32_t nstates, int steps, int mode,
                      unsigned long long *cycles, uint32_t *sink)
{
    extern __shared__ __align__(16) unsigned char raw[];
    E *tab = (E *)raw;
    for (uint32_t i = threadIdx.x; i < entries; i += blockDim.x) tab[i] = gtab[i];
    __syncthreads();
    uint32_t base = (uint32_t)__cvta_generic_to_shared(tab);
    uint32_t s[ILP];
    for (int q = 0; q < ILP; ++q) s[q] = (threadIdx.x * 2654435761u + blockIdx.x * 40503u + q * 977u) % nstates;
    uint32_t x = threadIdx.x * 747796405u + blockIdx.x * 2891336453u + 1u;
    unsigned long long t0 = clock64();
    for (int k = 0; k < steps; k += 8) {
        x = x * 1664525u + 1013904223u;
        uint32_t w = x;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            uint32_t c = (mode == 1) ? 0u : ((w >> (4 * j)) & 15u) % ncls;
#pragma unroll
            for (int q = 0; q < ILP; ++q) {
                uint32_t addr = base + (s[q] * ncls + c) * (uint32_t)sizeof(E);
                uint32_t v;
                if (sizeof(E) == 2) asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(addr));
                else asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
                s[q] = v;
            }
        }
    }
    unsigned long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    uint32_t acc = 0; for (int q = 0; q < ILP; ++q) acc ^= s[q];
    if (acc == 0xffffffffu) sink[0] = acc;
}

template <typename E, int ILP>
void run(int nstates, int ncls, int threads, int mode)
{
    const uint32_t entries = (uint32_t)nstates * ncls;
    std::vector<E> h(entries);
    uint32_t x = 12345;
    for (uint32_t i = 0; i < entries; ++i) {
        x = x * 1103515245u + 12345u;
        h[i] = (mode == 1) ? (E)0 : (E)((x >> 8) % nstates);
    }
    E *d; unsigned long long *cyc; uint32_t *sink;
    cudaMalloc(&d, entries * sizeof(E)); cudaMalloc(&cyc, 148 * 8); cudaMalloc(&sink, 4);
    cudaMemcpy(d, h.data(), entries * sizeof(E), cudaMemcpyHostToDevice);
    size_t smem = entries * sizeof(E);
    cudaFuncSetAttribute(chain<E, ILP>, cudaFuncAttributeMaxDynamicSharedMemorySize, 230000);
    const int steps = 1 << 14;
    chain<E, ILP><<<148, threads, smem>>>(d, entries, ncls, nstates, steps, mode, cyc, sink);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    chain<E, ILP><<<148, threads, smem>>>(d, entries, ncls, nstates, steps, mode, cyc, sink);
    cudaEventRecord(b); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, a, b);
    unsigned long long hc[148]; cudaMemcpy(hc, cyc, sizeof(hc), cudaMemcpyDeviceToHost);
    double avg = 0; for (auto v : hc) avg += v; avg /= 148;
    double cyc_per_step = avg / steps;
    double steps_per_cyc_sm = (double)threads * ILP / cyc_per_step;     // thread-steps per cycle per SM
    double tbs = 148.0 * threads * ILP * steps / (ms * 1e-3) / 1e12;     // bytes/s if 1 step = 1 byte
    printf("ILP=%d E=%zu states=%6d ncls=%2d smem=%7zu threads=%4d mode=%d : %.1f cyc/step/warp, %.2f steps/cyc/SM, %.2f Tsteps/s, err=%s\n",
           ILP, sizeof(E), nstates, ncls, smem, threads, mode, cyc_per_step, steps_per_cyc_sm, tbs, cudaGetErrorString(cudaGetLastError()));
    cudaFree(d); cudaFree(cyc); cudaFree(sink);
}

int main()
{
    for (int threads : {512, 1024}) {
        run<uint16_t, 1>(16000, 7, threads, 0);
        run<uint16_t, 2>(16000, 7, threads, 0);
        run<uint16_t, 4>(16000, 7, threads, 0);
        run<uint16_t, 8>(16000, 7, threads, 0);
        run<uint32_t, 4>(8000, 7, threads, 0);
        run<uint16_t, 4>(16000, 7, threads, 1);
    }
    return 0;
}
