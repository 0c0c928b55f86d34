// Probe: 256 site-pattern counts of a 1024-locus tile through the legacy integer tensor-core path
// (mma.sync m16n8k32 u8 = IMMA.16832.U8.U8 on sm_100a) instead of 256 x (AND + POPC).
//
//   T[ab][cd] = sum over loci  X_ab[l] * Y_cd[l],   X_ab = A_a & B_b,  Y_cd = C_c & D_d   (bit planes)
//
// A bit plane becomes an u8 operand with ONE LOP3 per four loci: (word & (0x01010101 << v)) gives the
// bytes {0, 2^v}; both operands use the same mask, so the products are {0, 4^v} and each v (0..3, on the
// word and on the word >> 4) accumulates into its own s32 fragment; count = sum_v acc_v >> 2v.
// The K order of an MMA is free as long as A and B agree, so lane (g = lane/4, j = lane%4) feeds
// the four loci of byte-column j of ITS OWN word pair: no transposition of the data is needed.
//
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -o imma_count_probe imma_count_probe.cu
// Run:    ./imma_count_probe            (prints clocks per tile per SM sub-partition for several shapes)
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)

static constexpr int kStride = 34;            // words per plane row in shared memory (32 + 2: conflict-free LDS.64)

__device__ __forceinline__ void imma(int (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1)
{
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// mode bit 0: do the unpack LOP3s, bit 1: do the IMMAs, bit 2: load the planes from global every tile
template <int kMode>
__global__ void __launch_bounds__(256) imma_count(const uint32_t *__restrict__ planes /*[ntile_src][16][32]*/, int ntile_src,
                                                  int tiles_per_warp, uint32_t *__restrict__ out /*[warps][256]*/)
{
    extern __shared__ uint32_t smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int gw = blockIdx.x * (blockDim.x >> 5) + warp;
    uint32_t *P = smem + warp * 16 * kStride;
    const int g = lane >> 2, j = lane & 3, p = g >> 2, r = g & 3;
    // plane order in P: A0..3, B0..3, C0..3, D0..3
    const uint32_t *pA0 = P + (0 + p) * kStride + 8 * j, *pA1 = P + (2 + p) * kStride + 8 * j;
    const uint32_t *pB = P + (4 + r) * kStride + 8 * j;
    const uint32_t *pC0 = P + (8 + p) * kStride + 8 * j, *pC1 = P + (10 + p) * kStride + 8 * j;
    const uint32_t *pD = P + (12 + r) * kStride + 8 * j;

    int acc[4][2][4];
#pragma unroll
    for (int v = 0; v < 4; ++v)
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[v][h][i] = 0;

    uint32_t w[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) w[k] = planes[((size_t)(gw % ntile_src) * 16 + k) * 32 + lane];

#pragma unroll 1
    for (int t = 0; t < tiles_per_warp; ++t) {
        if (kMode & 4) {
            const int src = (gw + t) % ntile_src;
#pragma unroll
            for (int k = 0; k < 16; ++k) w[k] = planes[((size_t)src * 16 + k) * 32 + lane];
        }
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 16; ++k) P[k * kStride + lane] = w[k];
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const uint2 a0 = *reinterpret_cast<const uint2 *>(pA0 + 2 * i), a1 = *reinterpret_cast<const uint2 *>(pA1 + 2 * i);
            const uint2 b = *reinterpret_cast<const uint2 *>(pB + 2 * i);
            const uint2 c0 = *reinterpret_cast<const uint2 *>(pC0 + 2 * i), c1 = *reinterpret_cast<const uint2 *>(pC1 + 2 * i);
            const uint2 d = *reinterpret_cast<const uint2 *>(pD + 2 * i);
            uint32_t x[2][4], y[2][4];     // [shift][row lo W, row hi W, row lo W', row hi W'] / [shift][half0 W, half0 W', half1 W, half1 W']
            x[0][0] = a0.x & b.x; x[0][1] = a1.x & b.x; x[0][2] = a0.y & b.y; x[0][3] = a1.y & b.y;
            y[0][0] = c0.x & d.x; y[0][1] = c0.y & d.y; y[0][2] = c1.x & d.x; y[0][3] = c1.y & d.y;
#pragma unroll
            for (int q = 0; q < 4; ++q) { x[1][q] = x[0][q] >> 4; y[1][q] = y[0][q] >> 4; }
#pragma unroll
            for (int s = 0; s < 2; ++s)
#pragma unroll
                for (int v = 0; v < 4; ++v) {
                    const uint32_t m = 0x01010101u << v;
                    uint32_t fa[4], fb[4];
                    if (kMode & 1) {
#pragma unroll
                        for (int q = 0; q < 4; ++q) { fa[q] = x[s][q] & m; fb[q] = y[s][q] & m; }
                    } else {
#pragma unroll
                        for (int q = 0; q < 4; ++q) { fa[q] = x[s][q]; fb[q] = y[s][q]; }
                    }
                    if (kMode & 2) {
                        imma(acc[v][0], fa[0], fa[1], fa[2], fa[3], fb[0], fb[1]);
                        imma(acc[v][1], fa[0], fa[1], fa[2], fa[3], fb[2], fb[3]);
                    } else {
#pragma unroll
                        for (int q = 0; q < 4; ++q) { acc[v][0][q] += fa[q] ^ fb[q]; acc[v][1][q] ^= fa[q] + fb[(q + 2) & 3]; }
                    }
                }
        }
    }
    // T[m][n]: lane (g, j) holds rows g, g+8 and columns 2j, 2j+1 (+8 for the second half)
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            uint32_t cnt = 0;
#pragma unroll
            for (int v = 0; v < 4; ++v) cnt += (uint32_t)acc[v][h][i] >> (2 * v);
            const int row = g + 8 * (i >> 1), col = 2 * j + (i & 1) + 8 * h;
            out[(size_t)gw * 256 + row * 16 + col] = cnt;
        }
}

// brute-force check: one block per warp-of-the-probe, thread = pattern
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

// raw IMMA throughput: kChains independent accumulators per warp
template <int kChains>
__global__ void __launch_bounds__(1024) imma_peak(int iters, int *out)
{
    int acc[kChains][4];
#pragma unroll
    for (int c = 0; c < kChains; ++c)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[c][i] = 0;
    uint32_t a = threadIdx.x * 0x01010101u, b = threadIdx.x + 7;
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int c = 0; c < kChains; ++c) imma(acc[c], a, a + c, b, a ^ b, b, a + it);
    }
    int s = 0;
#pragma unroll
    for (int c = 0; c < kChains; ++c) s += acc[c][0] + acc[c][1] + acc[c][2] + acc[c][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static float time_ms(F launch, int reps = 5)
{
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; ++r) {
        CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    return best;
}

int main()
{
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount; int khz; CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0));
    printf("device %s, %d SMs, %.0f MHz max\n", prop.name, sms, khz / 1e3);
    const int ntile_src = 64;
    std::vector<uint32_t> h((size_t)ntile_src * 16 * 32, 0u);
    uint64_t s = 88172645463325252ull;
    auto rnd = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return (uint32_t)(s >> 11); };
    for (int t = 0; t < ntile_src; ++t)
        for (int wd = 0; wd < 32; ++wd)
            for (int bit = 0; bit < 32; ++bit) {
                if (rnd() % 10 == 0) continue;                         // locus without a selected column
                for (int tip = 0; tip < 4; ++tip) {
                    const int code = (rnd() % 3 == 0) ? rnd() % 4 : (t + wd + bit) % 4;
                    h[((size_t)t * 16 + 4 * tip + code) * 32 + wd] |= 1u << bit;
                }
            }
    uint32_t *d_planes; CK(cudaMalloc(&d_planes, h.size() * 4)); CK(cudaMemcpy(d_planes, h.data(), h.size() * 4, cudaMemcpyHostToDevice));

    // correctness
    {
        const int blocks = 8, wpb = 8, warps = blocks * wpb, tiles = 49;
        uint32_t *d_out, *d_ref; CK(cudaMalloc(&d_out, warps * 256 * 4)); CK(cudaMalloc(&d_ref, warps * 256 * 4));
        imma_count<7><<<blocks, wpb * 32, wpb * 16 * kStride * 4>>>(d_planes, ntile_src, tiles, d_out);
        ref_count<<<warps, 256>>>(d_planes, ntile_src, tiles, d_ref);
        CK(cudaDeviceSynchronize());
        std::vector<uint32_t> o(warps * 256), r(warps * 256);
        CK(cudaMemcpy(o.data(), d_out, o.size() * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(r.data(), d_ref, r.size() * 4, cudaMemcpyDeviceToHost));
        size_t bad = 0, tot = 0; for (size_t i = 0; i < o.size(); ++i) { bad += o[i] != r[i]; tot += r[i]; }
        printf("IMMA counts vs AND+POPC counts: %zu mismatches of %zu (sum of counts %zu)\n", bad, o.size(), tot);
        if (bad) for (size_t i = 0, shown = 0; i < o.size() && shown < 8; ++i) if (o[i] != r[i]) { printf("  [%zu] imma %u ref %u\n", i, o[i], r[i]); ++shown; }
    }
    // throughput of the tile loop
    const int tiles = 2000;
    uint32_t *d_out; CK(cudaMalloc(&d_out, (size_t)sms * 64 * 256 * 4));
    printf("%-34s %8s %10s %14s %16s\n", "kernel", "warps/SM", "ms", "tiles/s/SM", "clk/tile/SMSP@max");
    auto report = [&](const char *name, int wps, float ms) {
        const double tiles_sm = (double)wps * tiles / (ms * 1e-3);
        printf("%-34s %8d %10.3f %14.4g %16.1f\n", name, wps, ms, tiles_sm, khz * 1e3 / tiles_sm * 4);
    };
    for (int wps : {4, 8, 12, 16, 24, 32}) {
        const int wpb = 4, bps = wps / wpb; const size_t sh = wpb * 16 * kStride * 4;
        report("imma_count full (ld+unpack+imma)", wps, time_ms([&] { imma_count<7><<<sms * bps, wpb * 32, sh>>>(d_planes, ntile_src, tiles, d_out); }));
        report("  no global loads", wps, time_ms([&] { imma_count<3><<<sms * bps, wpb * 32, sh>>>(d_planes, ntile_src, tiles, d_out); }));
        report("  unpack only (no IMMA)", wps, time_ms([&] { imma_count<1><<<sms * bps, wpb * 32, sh>>>(d_planes, ntile_src, tiles, d_out); }));
        report("  IMMA only (no unpack masks)", wps, time_ms([&] { imma_count<2><<<sms * bps, wpb * 32, sh>>>(d_planes, ntile_src, tiles, d_out); }));
    }
    CK(cudaGetLastError());
    // raw IMMA rate
    int *d_i; CK(cudaMalloc(&d_i, (size_t)sms * 1024 * 4));
    for (int wps : {4, 8, 16, 32}) {
        const int iters = 20000;
        float ms = time_ms([&] { imma_peak<8><<<sms, wps * 32>>>(iters, d_i); });
        const double per_sm = (double)wps * 8 * iters / (ms * 1e-3);
        printf("IMMA.16832.U8 peak: %2d warps/SM, 8 chains: %.4g IMMA/s/SM = %.2f clk per IMMA per SMSP, %.1f TOP/s chip\n", wps, per_sm,
               khz * 1e3 / per_sm * 4, per_sm * sms * 2.0 * 16 * 8 * 32 / 1e12);
    }
    CK(cudaGetLastError());
    return 0;
}
