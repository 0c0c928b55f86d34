// Probe: issue rate of the legacy (mma.sync) tensor-core instructions on sm_100a, per SM sub-partition.
//   kind 0: IMMA  m16n8k32 u8 x u8 -> s32        kind 1: QMMA  m16n8k32 e4m3 x e4m3 -> f32
//   kind 2: HMMA  m16n8k16 f16 x f16 -> f32      kind 3: IMMA with per-instruction operand changes (LOP3 between)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate_probe mma_rate_probe.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

template <int kKind>
__global__ void __launch_bounds__(1024) rate(int iters, uint32_t seed, uint32_t *out)
{
    constexpr int kChains = 6;
    uint32_t a[4], b[2];
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = (threadIdx.x * 2654435761u + seed * (i + 1)) & 0x04040404u;
    b[0] = (threadIdx.x * 40503u + seed) & 0x04040404u; b[1] = b[0] >> 1;
    int ci[kChains][4]; float cf[kChains][4];
#pragma unroll
    for (int c = 0; c < kChains; ++c)
#pragma unroll
        for (int i = 0; i < 4; ++i) { ci[c][i] = 0; cf[c][i] = 0.f; }
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int c = 0; c < kChains; ++c) {
            uint32_t a0 = a[0], a1 = a[1], a2 = a[2], a3 = a[3], b0 = b[0], b1 = b[1];
            if (kKind == 3) { const uint32_t m = 0x01010101u << (c & 3); a0 &= m; a1 &= m; a2 &= m; a3 &= m; b0 &= m; b1 &= m; a[c & 3] ^= it; }
            if (kKind == 0 || kKind == 3)
                asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                             : "+r"(ci[c][0]), "+r"(ci[c][1]), "+r"(ci[c][2]), "+r"(ci[c][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
            else if (kKind == 1)
                asm volatile("mma.sync.aligned.m16n8k32.row.col.f32.e4m3.e4m3.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                             : "+f"(cf[c][0]), "+f"(cf[c][1]), "+f"(cf[c][2]), "+f"(cf[c][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
            else
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                             : "+f"(cf[c][0]), "+f"(cf[c][1]), "+f"(cf[c][2]), "+f"(cf[c][3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
        }
    }
    uint32_t s = 0;
#pragma unroll
    for (int c = 0; c < kChains; ++c)
#pragma unroll
        for (int i = 0; i < 4; ++i) s += (uint32_t)ci[c][i] + __float_as_uint(cf[c][i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// exactness of the FP8 path: counts of 0/1 data carried as e4m3 powers of two in f32 accumulators over many MMAs
__global__ void fp8_exact(int nmma, uint32_t *out)
{
    // every byte of A and B = 0x04 (2^-7): each MMA adds 32 * 2^-14 to every accumulator
    float c[4] = {0.f, 0.f, 0.f, 0.f};
    const uint32_t v = 0x04040404u;
    for (int i = 0; i < nmma; ++i)
        asm volatile("mma.sync.aligned.m16n8k32.row.col.f32.e4m3.e4m3.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                     : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(v), "r"(v), "r"(v), "r"(v), "r"(v), "r"(v));
    if (threadIdx.x == 0) { out[0] = (uint32_t)(c[0] * 16384.f); out[1] = 32u * (uint32_t)nmma; }
}

int main()
{
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount; int khz; CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0));
    uint32_t *d; CK(cudaMalloc(&d, (size_t)sms * 1024 * 4));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const char *names[4] = {"IMMA.16832.U8 s32", "QMMA.16832 e4m3 f32", "HMMA.16816 f16 f32", "IMMA + 6 LOP3 each"};
    const int iters = 20000;
    for (int kind = 0; kind < 4; ++kind)
        for (int wps : {4, 8, 16, 32}) {
            float best = 1e30f;
            for (int rep = 0; rep < 4; ++rep) {
                CK(cudaEventRecord(e0));
                if (kind == 0) rate<0><<<sms, wps * 32>>>(iters, 12345u + rep, d);
                if (kind == 1) rate<1><<<sms, wps * 32>>>(iters, 12345u + rep, d);
                if (kind == 2) rate<2><<<sms, wps * 32>>>(iters, 12345u + rep, d);
                if (kind == 3) rate<3><<<sms, wps * 32>>>(iters, 12345u + rep, d);
                CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
                float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (rep && ms < best) best = ms;
            }
            const double per_sm = (double)wps * 6 * iters / (best * 1e-3);
            printf("%-22s %2d warps/SM: %.3f ms, %.2f clk per MMA per SMSP (at %d MHz)\n", names[kind], wps, best, khz * 1e3 / per_sm * 4, khz / 1000);
        }
    for (int n : {16, 512, 4096, 65536, 600000}) {
        fp8_exact<<<1, 32>>>(n, d); CK(cudaDeviceSynchronize());
        uint32_t h[2]; CK(cudaMemcpy(h, d, 8, cudaMemcpyDeviceToHost));
        printf("fp8 accumulate %d MMAs: got %u expected %u %s\n", n, h[0], h[1], h[0] == h[1] ? "exact" : "INEXACT");
    }
    return 0;
}
