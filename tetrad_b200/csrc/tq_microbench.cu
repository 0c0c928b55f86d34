// Pipe micro-benchmarks: roofline denominators measured on the box (tq_microbench in the C ABI; bench.py reports
// fractions of these).  POPC / LOP3+SHF / DFMA issue rates, POPC mixed with other MIO-queue work at the resolve
// kernel's occupancy, the legacy u8 tensor-core instruction's issue rate, a plain device copy.
#include "tq_kernels.cuh"

namespace tq {

__device__ __forceinline__ void imma_u8(int (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1)
{
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// ------------------------------------------------------------------------------------------
// pipe micro-benchmarks (roofline denominators measured on the box)
// ------------------------------------------------------------------------------------------
template <int kKind>
__global__ void __launch_bounds__(256) pipe_kernel(uint32_t *out, int iters, uint32_t seed)
{
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    if (kKind == 2) {
        double v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = 1.0 + 1e-9 * (double)(tid + u);
        const double a = 1.0 + 1e-12 * (double)seed, b = 1e-15;
        for (int it = 0; it < iters; ++it)
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = fma(v[u], a, b);
        double s = 0.0;
#pragma unroll
        for (int u = 0; u < 8; ++u) s += v[u];
        if (s == 123.456) out[0] = 1u;
    } else {
        uint32_t v[8], acc[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) { v[u] = tid * 2654435761u + seed + u; acc[u] = u; }
        for (int it = 0; it < iters; ++it)
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                if (kKind == 0) acc[u] += __popc(v[u] ^ acc[u]);
                else acc[u] = (acc[u] & v[u]) | (acc[u] >> 1);      // LOP3 + SHF
            }
        uint32_t s = 0;
#pragma unroll
        for (int u = 0; u < 8; ++u) s += acc[u];
        if (s == 0x12345u) out[0] = s;
    }
}

// POPC mixed with other MIO-queue instructions: kMix 0 = POPC only, 1 = + 4 SHFL per 8 POPC,
// 2 = + 2 LDS.128 per 8 POPC, 3 = + 8 independent DFMA per 8 POPC, 4 = + 16 LOP3/IADD per 8 POPC
template <int kMix>
__global__ void __launch_bounds__(128) mix_kernel(uint32_t *out, int iters, uint32_t seed)
{
    __shared__ uint4 sm[128];
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    sm[threadIdx.x] = make_uint4(tid, seed, tid ^ seed, 7u);
    __syncthreads();
    uint32_t v[8], acc[8], extra = seed;
    double d[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) { v[u] = tid * 2654435761u + seed + u; acc[u] = u; d[u] = 1.0 + u; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) acc[u] += __popc(v[u] ^ acc[u]);
        if (kMix == 1) {
#pragma unroll
            for (int u = 0; u < 4; ++u) v[u] ^= __shfl_xor_sync(0xFFFFFFFFu, v[u + 4], 16);
        } else if (kMix == 2) {
            const uint4 a = sm[(threadIdx.x + it) & 127], b = sm[(threadIdx.x + 2 * it + 1) & 127];
            extra += a.x ^ a.w ^ b.y ^ b.z;
        } else if (kMix == 3) {
#pragma unroll
            for (int u = 0; u < 8; ++u) d[u] = fma(d[u], 1.0000001, 1e-9);
        } else if (kMix == 4) {
#pragma unroll
            for (int u = 0; u < 8; ++u) { v[u] = (v[u] & 0x7FFFFFFFu) + (v[u] >> 3); extra ^= v[u] | acc[u]; }
        }
    }
    uint32_t s = extra;
    double ds = 0.0;
#pragma unroll
    for (int u = 0; u < 8; ++u) { s += acc[u] + v[u]; ds += d[u]; }
    if (s == 0x12345u || ds == 1.5) out[0] = s;
}

__global__ void __launch_bounds__(256) copy_kernel(const uint4 *__restrict__ src, uint4 *__restrict__ dst,
                                                   size_t n)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        dst[i] = src[i];
}

// kind 4: issue rate of the legacy u8 tensor-core instruction (IMMA.16832.U8.U8), 6 independent accumulator
// fragments per warp, operands one bit per byte as in count_tile_mma
__global__ void __launch_bounds__(512) imma_rate_kernel(uint32_t *out, int iters, uint32_t seed)
{
    uint32_t a[4], b[2];
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = (threadIdx.x * 2654435761u + seed * (i + 1)) & 0x04040404u;
    b[0] = (threadIdx.x * 40503u + seed) & 0x04040404u; b[1] = b[0] >> 1;
    int c[6][4];
#pragma unroll
    for (int k = 0; k < 6; ++k)
#pragma unroll
        for (int i = 0; i < 4; ++i) c[k][i] = 0;
#pragma unroll 1
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int k = 0; k < 6; ++k) imma_u8(c[k], a[0], a[1], a[2], a[3], b[0], b[1]);
    uint32_t s = 0;
#pragma unroll
    for (int k = 0; k < 6; ++k) s += (uint32_t)(c[k][0] + c[k][1] + c[k][2] + c[k][3]);
    if (s == 0x12345u) out[0] = s;
}

cudaError_t run_microbench(int kind, double *result)
{
    cudaError_t err;
    cudaEvent_t e0, e1;
    if ((err = cudaEventCreate(&e0)) != cudaSuccess) return err;
    if ((err = cudaEventCreate(&e1)) != cudaSuccess) return err;
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    float best_ms = 1e30f;
    double work = 0.0;
    if (kind >= 10) {   // POPC + other instruction mixes at the resolve kernel's occupancy (16 warps/SM)
        uint32_t *out = nullptr;
        if ((err = cudaMalloc(&out, 64)) != cudaSuccess) return err;
        const int iters = 4096, blocks = sms * 4, threads = 128;
        for (int rep = 0; rep < 5; ++rep) {
            cudaEventRecord(e0);
            switch (kind) {
            case 10: mix_kernel<0><<<blocks, threads>>>(out, iters, rep); break;
            case 11: mix_kernel<1><<<blocks, threads>>>(out, iters, rep); break;
            case 12: mix_kernel<2><<<blocks, threads>>>(out, iters, rep); break;
            case 13: mix_kernel<3><<<blocks, threads>>>(out, iters, rep); break;
            default: mix_kernel<4><<<blocks, threads>>>(out, iters, rep); break;
            }
            ++g_launches;
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms = 0.f;
            cudaEventElapsedTime(&ms, e0, e1);
            if (rep > 0 && ms < best_ms) best_ms = ms;
        }
        cudaFree(out);
        work = (double)blocks * threads * (double)iters * 8.0;
    } else if (kind == 4) {
        uint32_t *out = nullptr;
        if ((err = cudaMalloc(&out, 64)) != cudaSuccess) return err;
        const int iters = 8192, blocks = sms, threads = 512;
        for (int rep = 0; rep < 5; ++rep) {
            cudaEventRecord(e0);
            imma_rate_kernel<<<blocks, threads>>>(out, iters, rep);
            ++g_launches;
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms = 0.f;
            cudaEventElapsedTime(&ms, e0, e1);
            if (rep > 0 && ms < best_ms) best_ms = ms;
        }
        cudaFree(out);
        work = (double)blocks * (threads / 32) * (double)iters * 6.0;   // warp-level IMMA instructions
    } else if (kind == 3) {
        const size_t bytes = (size_t)1 << 30;
        uint4 *a = nullptr, *b = nullptr;
        if ((err = cudaMalloc(&a, bytes)) != cudaSuccess) return err;
        if ((err = cudaMalloc(&b, bytes)) != cudaSuccess) { cudaFree(a); return err; }
        cudaMemset(a, 1, bytes);
        for (int rep = 0; rep < 6; ++rep) {
            cudaEventRecord(e0);
            copy_kernel<<<sms * 16, 256>>>(a, b, bytes / 16);
            ++g_launches;
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms = 0.f;
            cudaEventElapsedTime(&ms, e0, e1);
            if (rep > 0 && ms < best_ms) best_ms = ms;
        }
        cudaFree(a); cudaFree(b);
        work = 2.0 * (double)bytes;
    } else {
        uint32_t *out = nullptr;
        if ((err = cudaMalloc(&out, 64)) != cudaSuccess) return err;
        const int iters = 4096, blocks = sms * 8, threads = 256;
        for (int rep = 0; rep < 6; ++rep) {
            cudaEventRecord(e0);
            if (kind == 0) pipe_kernel<0><<<blocks, threads>>>(out, iters, rep);
            else if (kind == 1) pipe_kernel<1><<<blocks, threads>>>(out, iters, rep);
            else pipe_kernel<2><<<blocks, threads>>>(out, iters, rep);
            ++g_launches;
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms = 0.f;
            cudaEventElapsedTime(&ms, e0, e1);
            if (rep > 0 && ms < best_ms) best_ms = ms;
        }
        cudaFree(out);
        work = (double)blocks * threads * (double)iters * 8.0;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if ((err = cudaGetLastError()) != cudaSuccess) return err;
    *result = work / ((double)best_ms * 1e-3);
    return cudaSuccess;
}

}  // namespace tq
