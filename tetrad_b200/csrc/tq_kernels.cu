// Hand-written sm_100a kernels of the quartet-resolution hot path.
//
//   pack_kernel     K1  u8 tmparr -> locus-tiled one-hot A/C/G/T bit-planes
//   resolve_kernel  K3+K4  per quartet: select one SNP per locus, 256 site-pattern counts by
//                   AND + POPC, three 16x16 one-sided Jacobi SVDs in FP64, rank / minrank /
//                   tail score / argmin -- one warp per quartet, nothing but the result leaves
//                   the SM.
//
// Reference semantics restated here (paths relative to the reference repo):
//   tetrad/src/resolve_quartets.py:212-221  gather, missing mask, invariant mask
//   tetrad/src/resolve_quartets.py:42-73    subsample_chunk_to_matrices (first usable SNP/locus)
//   tetrad/src/resolve_quartets.py:76-104   full_chunk_to_matrices
//   tetrad/src/resolve_quartets.py:226-264  nsnps, svd, matrix_rank, minrank, score, argmin
//
// Data layout: locus runs are sorted by length (longest first) and cut into
// tiles of 1024 loci; a tile with K ranks stores, for every tip, K x 32 uint4 {A,C,G,T}:
// word `lane` of rank k holds, at bit b, the k-th column of sorted locus 1024*tile+32*lane+b.
// "First usable column of a locus" (the reference's last_loc state machine) is then
// "lowest rank k whose bit survives the quartet's mask" -- a running OR per lane.
#include "tq_kernels.cuh"

namespace tq {

int64_t g_launches = 0;

static constexpr unsigned kFull = 0xFFFFFFFFu;
static constexpr int kWarps = 4;              // warps (= quartets in flight) per CTA
static constexpr int kLdm = 18;               // padded column length of the smem matrices (doubles)
static constexpr double kEps = 2.220446049250313e-16;
static constexpr int kMaxSweeps = 30;
static constexpr uint32_t kFlushEvery = 2047; // 2047 * 32 < 65536: 16-bit packed counters stay exact

// ------------------------------------------------------------------------------------------
// K1: pack
// grid (ntiles, ntips), 128 threads.  warp w packs ranks w, w+4, ...; lane = word (32 loci).
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) pack_kernel(PlanDev plan, const uint8_t *__restrict__ tmparr,
                                                   int64_t row_stride, uint4 *__restrict__ planes)
{
    const int tile = blockIdx.x;
    const int tip = blockIdx.y;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const uint32_t K = plan.tile_k[tile];
    const uint32_t off = plan.tile_off[tile];
    const uint8_t *row = tmparr + (int64_t)tip * row_stride;
    uint4 *out = planes + (size_t)tip * plan.row_u4 + off;
    const uint32_t l0 = (uint32_t)tile * kTileLoci + lane * 32;
    for (uint32_t k = warp; k < K; k += 4) {
        uint4 w = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll 4
        for (int b = 0; b < 32; ++b) {
            const uint32_t len = __ldg(plan.loc_len + l0 + b);
            if (k < len) {
                const uint32_t code = row[(size_t)__ldg(plan.loc_start + l0 + b) + k];
                const uint32_t bit = 1u << b;
                w.x |= (code == 0u) ? bit : 0u;
                w.y |= (code == 1u) ? bit : 0u;
                w.z |= (code == 2u) ? bit : 0u;
                w.w |= (code == 3u) ? bit : 0u;
            }
        }
        out[k * kTileWords + lane] = w;
    }
}

cudaError_t launch_pack(const PlanDev &plan, const uint8_t *d_tmparr, int64_t ntips,
                        int64_t row_stride, uint32_t *d_planes, cudaStream_t stream)
{
    if (plan.ntiles == 0 || ntips == 0) return cudaSuccess;
    dim3 grid((unsigned)plan.ntiles, (unsigned)ntips);
    pack_kernel<<<grid, 128, 0, stream>>>(plan, d_tmparr, row_stride,
                                          reinterpret_cast<uint4 *>(d_planes));
    ++g_launches;
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// K3: counts
// ------------------------------------------------------------------------------------------
// acc[i] holds two 16-bit counters: pattern 2i (low half) and 2i+1 (high half), where
// pattern p = 64a + 16b + 4c + d = 16*(4a+b) + (4c+d).
__device__ __forceinline__ void accumulate256(const uint32_t (&A)[4], const uint32_t (&B)[4],
                                              const uint32_t (&C)[4], const uint32_t (&D)[4],
                                              uint32_t (&acc)[128])
{
    uint32_t CD[16];
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int d = 0; d < 4; ++d) CD[4 * c + d] = C[c] & D[d];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const uint32_t ab = A[a] & B[b];
#pragma unroll
            for (int h = 0; h < 8; ++h) {
                const uint32_t p0 = __popc(ab & CD[2 * h]);
                const uint32_t p1 = __popc(ab & CD[2 * h + 1]);
                acc[(4 * a + b) * 8 + h] += p0 + (p1 << 16);
            }
        }
}

// Warp-reduce the packed per-lane counters into the warp's T[256] (shared memory), then clear.
__device__ __forceinline__ void flush_acc(uint32_t (&acc)[128], uint32_t *T, int lane)
{
#pragma unroll
    for (int i = 0; i < 128; ++i) {
        const uint32_t lo = __reduce_add_sync(kFull, acc[i] & 0xFFFFu);
        const uint32_t hi = __reduce_add_sync(kFull, acc[i] >> 16);
        if (lane == 0) {
            uint2 *t = reinterpret_cast<uint2 *>(T + 2 * i);
            uint2 v = *t;
            v.x += lo;
            v.y += hi;
            *t = v;
        }
        acc[i] = 0u;
    }
}

template <bool kSubsample>
__device__ __forceinline__ void count_quartet(const PlanDev &plan, const uint4 *__restrict__ r0,
                                              const uint4 *__restrict__ r1,
                                              const uint4 *__restrict__ r2,
                                              const uint4 *__restrict__ r3, uint32_t *T, int lane)
{
    uint32_t acc[128];
#pragma unroll
    for (int i = 0; i < 128; ++i) acc[i] = 0u;
    uint32_t pending = 0;   // accumulate256 calls since the last flush (each adds <= 32 per counter)

    for (int tile = 0; tile < plan.ntiles; ++tile) {
        const uint32_t off = __ldg(plan.tile_off + tile) + lane;
        const uint32_t K = __ldg(plan.tile_k + tile);
        if (kSubsample) {
            // stage 1: lowest rank whose column is complete and segregating for this quartet
            uint32_t taken = 0u;
            uint32_t MA[4] = {0u, 0u, 0u, 0u}, MB[4] = {0u, 0u, 0u, 0u};
            uint32_t MC[4] = {0u, 0u, 0u, 0u}, MD[4] = {0u, 0u, 0u, 0u};
            uint4 a = __ldg(r0 + off), b = __ldg(r1 + off), c = __ldg(r2 + off), d = __ldg(r3 + off);
            for (uint32_t k = 0; k < K; ++k) {
                uint4 na = a, nb = b, nc = c, nd = d;
                if (k + 1 < K) {
                    const uint32_t o2 = off + (k + 1) * kTileWords;
                    na = __ldg(r0 + o2); nb = __ldg(r1 + o2); nc = __ldg(r2 + o2); nd = __ldg(r3 + o2);
                }
                const uint32_t valid = (a.x | a.y | a.z | a.w) & (b.x | b.y | b.z | b.w) &
                                       (c.x | c.y | c.z | c.w) & (d.x | d.y | d.z | d.w);
                const uint32_t same = (a.x & b.x & c.x & d.x) | (a.y & b.y & c.y & d.y) |
                                      (a.z & b.z & c.z & d.z) | (a.w & b.w & c.w & d.w);
                const uint32_t sel = valid & ~same & ~taken;
                taken |= sel;
                MA[0] |= a.x & sel; MA[1] |= a.y & sel; MA[2] |= a.z & sel; MA[3] |= a.w & sel;
                MB[0] |= b.x & sel; MB[1] |= b.y & sel; MB[2] |= b.z & sel; MB[3] |= b.w & sel;
                MC[0] |= c.x & sel; MC[1] |= c.y & sel; MC[2] |= c.z & sel; MC[3] |= c.w & sel;
                MD[0] |= d.x & sel; MD[1] |= d.y & sel; MD[2] |= d.z & sel; MD[3] |= d.w & sel;
                a = na; b = nb; c = nc; d = nd;
            }
            // stage 2: 256 pattern counts over the tile's 1024 loci
            accumulate256(MA, MB, MC, MD, acc);
            if (++pending == kFlushEvery) { flush_acc(acc, T, lane); pending = 0; }
        } else {
            for (uint32_t k = 0; k < K; ++k) {
                const uint32_t o = off + k * kTileWords;
                const uint4 a = __ldg(r0 + o), b = __ldg(r1 + o), c = __ldg(r2 + o), d = __ldg(r3 + o);
                const uint32_t A[4] = {a.x, a.y, a.z, a.w}, B[4] = {b.x, b.y, b.z, b.w};
                const uint32_t C[4] = {c.x, c.y, c.z, c.w}, D[4] = {d.x, d.y, d.z, d.w};
                accumulate256(A, B, C, D, acc);
                if (++pending == kFlushEvery) { flush_acc(acc, T, lane); pending = 0; }
            }
        }
    }
    flush_acc(acc, T, lane);
}

// ------------------------------------------------------------------------------------------
// K4: scores.  One-sided (Hestenes) Jacobi on the three 16x16 flattenings, columns in shared
// memory, lane (m, s) = (lane / 8, lane % 8) rotates pair s of matrix m in each of the 15
// rounds of a round-robin sweep.
// ------------------------------------------------------------------------------------------
struct __align__(16) WarpScratch {
    uint32_t T[256];
    double M[3][16][kLdm];   // M[m][col][row]
    double sv[3][16];
    double srt[3][16];
};

__device__ __forceinline__ uint64_t splitmix64(uint64_t x)
{
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

__device__ __forceinline__ void svd_scores(WarpScratch &ws, int lane, double (&score)[3],
                                           uint32_t &best, uint32_t &minrank)
{
    // the three flattenings (resolve_quartets.py:63,67-72), stored column-major
    double fro2 = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int p = lane + 32 * i;
        const int a = p >> 6, b = (p >> 4) & 3, c = (p >> 2) & 3, d = p & 3;
        const double v = (double)ws.T[p];
        fro2 += v * v;
        ws.M[0][4 * c + d][4 * a + b] = v;   // rows (a,b), cols (c,d)
        ws.M[1][4 * b + d][4 * a + c] = v;   // rows (a,c), cols (b,d)
        ws.M[2][4 * b + c][4 * a + d] = v;   // rows (a,d), cols (b,c)
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) fro2 += __shfl_xor_sync(kFull, fro2, o);
    const double tiny = (kEps * kEps) * fro2;   // columns this small are null directions
    __syncwarp();

    const int m = lane >> 3, s = lane & 7;
    const bool active = lane < 24;
    for (int sweep = 0; sweep < kMaxSweeps; ++sweep) {
        bool rotated = false;
        for (int rnd = 0; rnd < 15; ++rnd) {
            if (active) {
                int i, j;
                if (s == 0) { i = 15; j = rnd; }
                else { i = (rnd + s) % 15; j = (rnd + 15 - s) % 15; }
                double2 *ci = reinterpret_cast<double2 *>(&ws.M[m][i][0]);
                double2 *cj = reinterpret_cast<double2 *>(&ws.M[m][j][0]);
                double2 x[8], y[8];
#pragma unroll
                for (int r = 0; r < 8; ++r) { x[r] = ci[r]; y[r] = cj[r]; }
                double al = 0.0, be = 0.0, ga = 0.0;
#pragma unroll
                for (int r = 0; r < 8; ++r) {
                    al = fma(x[r].x, x[r].x, al); al = fma(x[r].y, x[r].y, al);
                    be = fma(y[r].x, y[r].x, be); be = fma(y[r].y, y[r].y, be);
                    ga = fma(x[r].x, y[r].x, ga); ga = fma(x[r].y, y[r].y, ga);
                }
                if (al > tiny && be > tiny && fabs(ga) > (2.0 * kEps) * sqrt(al * be)) {
                    const double zeta = (be - al) / (2.0 * ga);
                    const double az = fabs(zeta);
                    double t = (az < 1e100) ? 1.0 / (az + sqrt(1.0 + zeta * zeta)) : 0.5 / az;
                    t = (zeta < 0.0) ? -t : t;
                    const double cs = 1.0 / sqrt(1.0 + t * t);
                    const double sn = cs * t;
#pragma unroll
                    for (int r = 0; r < 8; ++r) {
                        double2 nx, ny;
                        nx.x = cs * x[r].x - sn * y[r].x; nx.y = cs * x[r].y - sn * y[r].y;
                        ny.x = sn * x[r].x + cs * y[r].x; ny.y = sn * x[r].y + cs * y[r].y;
                        ci[r] = nx; cj[r] = ny;
                    }
                    rotated = true;
                }
            }
            __syncwarp();
        }
        if (!__any_sync(kFull, rotated)) break;
    }

    // singular values = column norms; sort descending per matrix by rank counting
    for (int c = lane; c < 48; c += 32) {
        const double *col = &ws.M[c >> 4][c & 15][0];
        double n2 = 0.0;
#pragma unroll
        for (int r = 0; r < 16; ++r) n2 = fma(col[r], col[r], n2);
        ws.sv[c >> 4][c & 15] = sqrt(n2);
    }
    __syncwarp();
    for (int c = lane; c < 48; c += 32) {
        const int mm = c >> 4, col = c & 15;
        const double v = ws.sv[mm][col];
        int pos = 0;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            const double u = ws.sv[mm][j];
            pos += (u > v || (u == v && j < col)) ? 1 : 0;
        }
        ws.srt[mm][pos] = v;
    }
    __syncwarp();
    // rank = #{sigma > sigma_max * 16 * eps} (numpy.linalg.matrix_rank), :243
    int rank = 16;
    if (lane < 3) {
        const double tol = ws.srt[lane][0] * (16.0 * kEps);
        rank = 0;
#pragma unroll
        for (int j = 0; j < 16; ++j) rank += (ws.srt[lane][j] > tol) ? 1 : 0;
    }
    int mr = min(__shfl_sync(kFull, rank, 0), min(__shfl_sync(kFull, rank, 1), __shfl_sync(kFull, rank, 2)));
    mr = min(mr, 10);                                                  // :246
    double sc = 0.0;
    if (lane < 3) {
        for (int j = mr; j < 16; ++j) sc = fma(ws.srt[lane][j], ws.srt[lane][j], sc);
        sc = sqrt(sc);                                                 // :248
    }
    score[0] = __shfl_sync(kFull, sc, 0);
    score[1] = __shfl_sync(kFull, sc, 1);
    score[2] = __shfl_sync(kFull, sc, 2);
    best = 0;                                                          // first argmin, :251
    if (score[1] < score[best]) best = 1;
    if (score[2] < score[best]) best = 2;
    minrank = (uint32_t)mr;
}

// ------------------------------------------------------------------------------------------
// fused resolve kernel: one warp per quartet, persistent grid-stride over quartets
// ------------------------------------------------------------------------------------------
template <bool kSubsample>
__global__ void __launch_bounds__(kWarps * 32, 2)
resolve_kernel(PlanDev plan, const uint4 *__restrict__ planes, uint32_t ntips,
               const uint4 *__restrict__ quartets, int64_t m, uint64_t seed,
               double *__restrict__ rscor, uint2 *__restrict__ rstat,
               uint32_t *__restrict__ counts, uint32_t *__restrict__ minrank_out,
               uint32_t *__restrict__ status)
{
    __shared__ WarpScratch scratch[kWarps];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    WarpScratch &ws = scratch[warp];
    const int64_t nwarps = (int64_t)gridDim.x * kWarps;

    for (int64_t q = (int64_t)blockIdx.x * kWarps + warp; q < m; q += nwarps) {
        const uint4 t = __ldg(quartets + q);
        if (t.x >= ntips || t.y >= ntips || t.z >= ntips || t.w >= ntips) {
            if (lane == 0) {
                rstat[q] = make_uint2(0xFFFFFFFFu, 0u);
                rscor[3 * q] = rscor[3 * q + 1] = rscor[3 * q + 2] = 0.0;
                if (status) atomicOr(status, 1u);
            }
            continue;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) ws.T[lane + 32 * i] = 0u;
        __syncwarp();
        count_quartet<kSubsample>(plan, planes + (size_t)t.x * plan.row_u4,
                                  planes + (size_t)t.y * plan.row_u4,
                                  planes + (size_t)t.z * plan.row_u4,
                                  planes + (size_t)t.w * plan.row_u4, ws.T, lane);
        __syncwarp();
        if (!kSubsample && lane < 4) ws.T[85u * lane] = 0u;   // invariant patterns aaaa, :218
        __syncwarp();
        uint32_t nsnps = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const uint32_t v = ws.T[lane + 32 * i];
            nsnps += v;
            if (counts) counts[q * 256 + lane + 32 * i] = v;
        }
        nsnps = __reduce_add_sync(kFull, nsnps);                        // :226

        double score[3] = {0.001, 0.001, 0.001};                        // :232
        uint32_t best, minrank = 0;
        if (nsnps == 0) {
            uint64_t h = seed;
            h = splitmix64(h ^ t.x); h = splitmix64(h ^ t.y);
            h = splitmix64(h ^ t.z); h = splitmix64(h ^ t.w);
            best = (uint32_t)(h % 3ull);                                // stands in for :231
        } else {
            svd_scores(ws, lane, score, best, minrank);
        }
        if (lane == 0) {
            rscor[3 * q] = score[0]; rscor[3 * q + 1] = score[1]; rscor[3 * q + 2] = score[2];
            rstat[q] = make_uint2(best, nsnps);                         // :251, :264
            if (minrank_out) minrank_out[q] = minrank;
        }
        __syncwarp();
    }
}

cudaError_t launch_resolve(const PlanDev &plan, const uint32_t *d_planes, int64_t ntips,
                           const uint32_t *d_quartets, int64_t m, bool subsample, uint64_t seed,
                           double *d_rscor, uint32_t *d_rstat, uint32_t *d_counts,
                           uint32_t *d_minrank, uint32_t *d_status, int num_sms,
                           cudaStream_t stream)
{
    if (m == 0) return cudaSuccess;
    const int64_t want = (m + kWarps - 1) / kWarps;
    const int64_t cap = (int64_t)num_sms * 2;          // 2 resident CTAs per SM (launch bounds)
    const unsigned grid = (unsigned)(want < cap ? want : cap);
    const uint4 *planes = reinterpret_cast<const uint4 *>(d_planes);
    const uint4 *quartets = reinterpret_cast<const uint4 *>(d_quartets);
    uint2 *rstat = reinterpret_cast<uint2 *>(d_rstat);
    if (subsample)
        resolve_kernel<true><<<grid, kWarps * 32, 0, stream>>>(plan, planes, (uint32_t)ntips, quartets,
                                                               m, seed, d_rscor, rstat, d_counts,
                                                               d_minrank, d_status);
    else
        resolve_kernel<false><<<grid, kWarps * 32, 0, stream>>>(plan, planes, (uint32_t)ntips, quartets,
                                                                m, seed, d_rscor, rstat, d_counts,
                                                                d_minrank, d_status);
    ++g_launches;
    return cudaGetLastError();
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

__global__ void __launch_bounds__(256) copy_kernel(const uint4 *__restrict__ src, uint4 *__restrict__ dst,
                                                   size_t n)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        dst[i] = src[i];
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
    if (kind == 3) {
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
