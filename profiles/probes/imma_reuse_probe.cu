// Probe: does the issue rate of IMMA.16832.U8.U8 on sm_100a depend on operand-register reuse or on the data?
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o imma_reuse_probe imma_reuse_probe.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)
__device__ __forceinline__ void imma(int (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1)
{
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
// kMode 0: every IMMA reads the same A and B registers      1: A shared by two consecutive IMMAs, B distinct
//       2: A and B distinct for each of 8 IMMAs              3: mode 0 with dense operands   4: mode 2 with dense operands
template <int kMode>
__global__ void __launch_bounds__(512) rate(int iters, uint32_t seed, uint32_t *out)
{
    const uint32_t dm = (kMode >= 3) ? 0xFFFFFFFFu : 0x04040404u;
    uint32_t a[8][4], b[8][2];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
#pragma unroll
        for (int i = 0; i < 4; ++i) a[k][i] = (threadIdx.x * 2654435761u + seed * (4 * k + i + 1)) & dm;
        b[k][0] = (threadIdx.x * 40503u + seed * (k + 3)) & dm; b[k][1] = (b[k][0] >> 1) & dm;
    }
    int c[8][4];
#pragma unroll
    for (int k = 0; k < 8; ++k)
#pragma unroll
        for (int i = 0; i < 4; ++i) c[k][i] = 0;
#pragma unroll 1
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const int ka = (kMode == 0 || kMode == 3) ? 0 : (kMode == 1) ? (k >> 1) : k;
            const int kb = (kMode == 0 || kMode == 3) ? 0 : k;
            imma(c[k], a[ka][0], a[ka][1], a[ka][2], a[ka][3], b[kb][0], b[kb][1]);
        }
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += (uint32_t)(c[k][0] + c[k][1] + c[k][2] + c[k][3]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main()
{
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount; int khz; CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0));
    uint32_t *d; CK(cudaMalloc(&d, (size_t)sms * 512 * 4));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const char *names[5] = {"same A,B (sparse)", "A shared by pairs, B distinct (sparse)", "A,B distinct (sparse)", "same A,B (dense)", "A,B distinct (dense)"};
    const int iters = 20000;
    for (int mode = 0; mode < 5; ++mode)
        for (int wps : {8, 16}) {
            float best = 1e30f;
            for (int rep = 0; rep < 4; ++rep) {
                CK(cudaEventRecord(e0));
                if (mode == 0) rate<0><<<sms, wps * 32>>>(iters, 12345u + rep, d);
                if (mode == 1) rate<1><<<sms, wps * 32>>>(iters, 12345u + rep, d);
                if (mode == 2) rate<2><<<sms, wps * 32>>>(iters, 12345u + rep, d);
                if (mode == 3) rate<3><<<sms, wps * 32>>>(iters, 12345u + rep, d);
                if (mode == 4) rate<4><<<sms, wps * 32>>>(iters, 12345u + rep, d);
                CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
                float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (rep && ms < best) best = ms;
            }
            const double per_sm = (double)wps * 8 * iters / (best * 1e-3);
            printf("%-40s %2d warps/SM: %.2f clk per IMMA per SMSP\n", names[mode], wps, khz * 1e3 / per_sm * 4);
        }
    return 0;
}
