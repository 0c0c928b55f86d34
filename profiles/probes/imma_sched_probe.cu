// Probe: schedules of the tensor-core tile count (count_tile_mma in tq_kernels.cu) in isolation.
//   variant 0: as shipped (4 scales x {word, word >> 4}, compiler-scheduled)
//   variant 1: same arithmetic, LOP3s as volatile asm in an order that puts 6 ALU ops between IMMAs
//   variant 2: 8 scales, no shifts, three-input LOP3 (a & b & mask) straight from the planes (64 accumulators)
//   variant 3: variant 2 with the interleaved volatile order
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o imma_sched_probe imma_sched_probe.cu
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)
static constexpr int kStride = 34;

__device__ __forceinline__ void imma(int (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1)
{
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t vand(uint32_t a, uint32_t b)
{
    uint32_t r; asm volatile("and.b32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
}
__device__ __forceinline__ uint32_t vand3(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t r; asm volatile("lop3.b32 %0, %1, %2, %3, 0x80;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r;
}

template <int V> struct Acc { int v[(V >= 2) ? 8 : 4][2][4]; };

template <int V>
__device__ __forceinline__ void tile(const uint32_t *pa, const uint32_t *pb, Acc<V> &A)
{
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const uint2 a0 = *reinterpret_cast<const uint2 *>(pa + 2 * i);
        const uint2 a1 = *reinterpret_cast<const uint2 *>(pa + 2 * kStride + 2 * i);
        const uint2 b = *reinterpret_cast<const uint2 *>(pb + 2 * i);
        const uint2 c0 = *reinterpret_cast<const uint2 *>(pa + 8 * kStride + 2 * i);
        const uint2 c1 = *reinterpret_cast<const uint2 *>(pa + 10 * kStride + 2 * i);
        const uint2 d = *reinterpret_cast<const uint2 *>(pb + 8 * kStride + 2 * i);
        if (V < 2) {
            uint32_t x[2][4], y[2][4];
            x[0][0] = a0.x & b.x; x[0][1] = a1.x & b.x; x[0][2] = a0.y & b.y; x[0][3] = a1.y & b.y;
            y[0][0] = c0.x & d.x; y[0][1] = c0.y & d.y; y[0][2] = c1.x & d.x; y[0][3] = c1.y & d.y;
#pragma unroll
            for (int q = 0; q < 4; ++q) { x[1][q] = x[0][q] >> 4; y[1][q] = y[0][q] >> 4; }
#pragma unroll
            for (int s = 0; s < 2; ++s)
#pragma unroll
                for (int v = 0; v < 4; ++v) {
                    const uint32_t m = 0x01010101u << v;
                    if (V == 0) {
                        const uint32_t f0 = x[s][0] & m, f1 = x[s][1] & m, f2 = x[s][2] & m, f3 = x[s][3] & m;
                        imma(A.v[v][0], f0, f1, f2, f3, y[s][0] & m, y[s][1] & m);
                        imma(A.v[v][1], f0, f1, f2, f3, y[s][2] & m, y[s][3] & m);
                    } else {
                        const uint32_t f0 = vand(x[s][0], m), f1 = vand(x[s][1], m), f2 = vand(x[s][2], m), f3 = vand(x[s][3], m);
                        const uint32_t g0 = vand(y[s][0], m), g1 = vand(y[s][1], m);
                        imma(A.v[v][0], f0, f1, f2, f3, g0, g1);
                        const uint32_t g2 = vand(y[s][2], m), g3 = vand(y[s][3], m);
                        imma(A.v[v][1], f0, f1, f2, f3, g2, g3);
                    }
                }
        } else {
#pragma unroll
            for (int v = 0; v < 8; ++v) {
                const uint32_t m = 0x01010101u << v;
                if (V == 2) {
                    const uint32_t f0 = a0.x & b.x & m, f1 = a1.x & b.x & m, f2 = a0.y & b.y & m, f3 = a1.y & b.y & m;
                    imma(A.v[v][0], f0, f1, f2, f3, c0.x & d.x & m, c0.y & d.y & m);
                    imma(A.v[v][1], f0, f1, f2, f3, c1.x & d.x & m, c1.y & d.y & m);
                } else {
                    const uint32_t f0 = vand3(a0.x, b.x, m), f1 = vand3(a1.x, b.x, m), f2 = vand3(a0.y, b.y, m), f3 = vand3(a1.y, b.y, m);
                    const uint32_t g0 = vand3(c0.x, d.x, m), g1 = vand3(c0.y, d.y, m);
                    imma(A.v[v][0], f0, f1, f2, f3, g0, g1);
                    const uint32_t g2 = vand3(c1.x, d.x, m), g3 = vand3(c1.y, d.y, m);
                    imma(A.v[v][1], f0, f1, f2, f3, g2, g3);
                }
            }
        }
    }
}

template <int V>
__global__ void __launch_bounds__(256) kern(const uint32_t *__restrict__ planes, int ntile_src, int tiles_per_warp, uint32_t *__restrict__ out)
{
    extern __shared__ uint32_t smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gw = blockIdx.x * (blockDim.x >> 5) + warp;
    uint32_t *P = smem + warp * 16 * kStride;
    const int g = lane >> 2, j = lane & 3;
    const uint32_t *pa = P + (g >> 2) * kStride + 8 * j, *pb = P + (4 + (g & 3)) * kStride + 8 * j;
    Acc<V> A;
    constexpr int NV = (V >= 2) ? 8 : 4;
#pragma unroll
    for (int v = 0; v < NV; ++v)
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int i = 0; i < 4; ++i) A.v[v][h][i] = 0;
#pragma unroll 1
    for (int t = 0; t < tiles_per_warp; ++t) {
        const int src = (gw + t) % ntile_src;
        uint32_t w[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) w[k] = planes[((size_t)src * 16 + k) * 32 + lane];
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 16; ++k) P[k * kStride + lane] = w[k];
        __syncwarp();
        tile<V>(pa, pb, A);
    }
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            uint32_t cnt = 0;
#pragma unroll
            for (int v = 0; v < NV; ++v) cnt += (uint32_t)A.v[v][h][i] >> (2 * (V >= 2 ? v : v));
            const int row = g + 8 * (i >> 1), col = 2 * j + (i & 1) + 8 * h;
            out[(size_t)gw * 256 + row * 16 + col] = cnt;
        }
}

__global__ void ref_count(const uint32_t *__restrict__ planes, int ntile_src, int tiles_per_warp, uint32_t *__restrict__ out)
{
    const int gw = blockIdx.x, pat = threadIdx.x;
    const int a = pat >> 6, b = (pat >> 4) & 3, c = (pat >> 2) & 3, d = pat & 3;
    uint32_t cnt = 0;
    for (int t = 0; t < tiles_per_warp; ++t) {
        const uint32_t *src = planes + (size_t)((gw + t) % ntile_src) * 16 * 32;
        for (int wd = 0; wd < 32; ++wd)
            cnt += __popc(src[(0 + a) * 32 + wd] & src[(4 + b) * 32 + wd] & src[(8 + c) * 32 + wd] & src[(12 + d) * 32 + wd]);
    }
    out[(size_t)gw * 256 + pat] = cnt;
}

template <int V>
static void run(const uint32_t *d_planes, int ntile_src, int sms, int khz, uint32_t *d_out, uint32_t *d_ref)
{
    {   // correctness
        const int blocks = 8, wpb = 4, warps = blocks * wpb, tiles = 49;
        kern<V><<<blocks, wpb * 32, wpb * 16 * kStride * 4>>>(d_planes, ntile_src, tiles, d_out);
        ref_count<<<warps, 256>>>(d_planes, ntile_src, tiles, d_ref);
        CK(cudaDeviceSynchronize());
        std::vector<uint32_t> o(warps * 256), r(warps * 256);
        CK(cudaMemcpy(o.data(), d_out, o.size() * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(r.data(), d_ref, r.size() * 4, cudaMemcpyDeviceToHost));
        size_t bad = 0; for (size_t i = 0; i < o.size(); ++i) bad += o[i] != r[i];
        printf("variant %d: %zu mismatches;", V, bad);
    }
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const int tiles = 2000;
    for (int wps : {4, 8, 12, 16}) {
        const int wpb = 4, bps = wps / wpb;
        float best = 1e30f;
        for (int rep = 0; rep < 4; ++rep) {
            CK(cudaEventRecord(e0));
            kern<V><<<sms * bps, wpb * 32, wpb * 16 * kStride * 4>>>(d_planes, ntile_src, tiles, d_out);
            CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (rep && ms < best) best = ms;
        }
        const double tiles_sm = (double)wps * tiles / (best * 1e-3);
        printf("  %2d w/SM: %6.1f clk/tile/SMSP", wps, khz * 1e3 / tiles_sm * 4);
    }
    printf("\n");
}

int main()
{
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount; int khz; CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0));
    const int ntile_src = 64;
    std::vector<uint32_t> h((size_t)ntile_src * 16 * 32, 0u);
    uint64_t s = 88172645463325252ull;
    auto rnd = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return (uint32_t)(s >> 11); };
    for (int t = 0; t < ntile_src; ++t)
        for (int wd = 0; wd < 32; ++wd)
            for (int bit = 0; bit < 32; ++bit) {
                if (rnd() % 2 == 0) continue;
                for (int tip = 0; tip < 4; ++tip) {
                    const int code = (rnd() % 3 == 0) ? rnd() % 4 : (t + wd + bit) % 4;
                    h[((size_t)t * 16 + 4 * tip + code) * 32 + wd] |= 1u << bit;
                }
            }
    uint32_t *d_planes, *d_out, *d_ref;
    CK(cudaMalloc(&d_planes, h.size() * 4)); CK(cudaMemcpy(d_planes, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMalloc(&d_out, (size_t)sms * 64 * 256 * 4)); CK(cudaMalloc(&d_ref, 32 * 256 * 4));
    run<0>(d_planes, ntile_src, sms, khz, d_out, d_ref);
    run<1>(d_planes, ntile_src, sms, khz, d_out, d_ref);
    run<2>(d_planes, ntile_src, sms, khz, d_out, d_ref);
    run<3>(d_planes, ntile_src, sms, khz, d_out, d_ref);
    return 0;
}
