// Per-replicate prepare kernels of the quartet-resolution hot path (sm_100a).
//
//   pack_planes_kernel<false>   K1     u8 tmparr -> locus-tiled one-hot A/C/G/T bit-planes
//   boot_keys_kernel,           K2/K5  bootstrap replicate on the device: one Philox4x32-10 key per column instance,
//   boot_scatter_kernel                one radix sort of (sorted locus, key) over all columns (cub::DeviceRadixSort) gives
//   pack_planes_kernel<true>           every instance its random column order whatever its length; planes are packed
//                                      straight from the resident ASCII seqarr, IUPAC cells resolved by Philox coins
//   unrank_kernel                      lexicographic rank -> quartet (combinations.py:94-106)
//
// The pack kernels give one warp one rank chunk (1024 loci x one rank) of a group of tips: lane b reads the byte of
// locus 32 w + b for each of the 32 words w -- 32 independent byte gathers per warp instruction instead of 32 dependent
// ones per lane -- and four warp ballots per word and tip assemble the plane words; a lane keeps word `lane` of every
// tip and the chunk leaves as one 512-byte store per tip.  Column indices and locus lengths are read once per word for
// all tips of the group.
//
// Reference semantics restated here (paths relative to the reference repo):
//   tetrad/src/resolve_quartets.py:212-216       row gather + `seqs >= 78` missing mask (done once per replicate here)
//   tetrad/src/run_inference.py:99-143           resample_tmp_database
//   tetrad/jit/resample.py:20-64                 jit_resample (loci with replacement, columns shuffled inside a locus)
//   tetrad/jit/resolve_ambigs.py:12-36, tetrad/src/utils.py:14-21   ambiguity codes -> one of their two bases
//   tetrad/src/combinations.py:94-106            _index_to_combination
#include "tq_kernels.cuh"

#include <cub/device/device_radix_sort.cuh>

namespace tq {

static constexpr unsigned kFull = 0xFFFFFFFFu;
static constexpr int kPackTips = 8;              // tips per warp pass (their 8 x 4 plane words live in registers)

__host__ __device__ __forceinline__ uint32_t philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                                           uint32_t k0, uint32_t k1)
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        c1 = (uint32_t)p1; c3 = (uint32_t)p0; c0 = n0; c2 = n2;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return c0;
}

// largest tile t with tile_off[t] <= first uint4 of rank chunk `chunk`
__device__ __forceinline__ int tile_of_chunk(const PlanDev &plan, uint32_t chunk)
{
    const uint32_t x = chunk * kTileWords;
    int lo = 0, hi = plan.ntiles - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (__ldg(plan.tile_off + mid) <= x) lo = mid; else hi = mid - 1;
    }
    return lo;
}

// Cell byte -> 0..3 (base), 4 (missing), 8 + i (two-base ambiguity class i of GETCONS, tetrad/src/utils.py:14-21:
// R=G|A K=G|T S=G|C Y=T|C W=T|A M=C|A; coin set -> first base).  kAscii: the matrix holds ASCII (seqarr), else codes.
template <bool kAscii>
__device__ __forceinline__ void fill_lut(uint8_t *lut, int tid, int nthreads)
{
    for (int ch = tid; ch < 256; ch += nthreads) {
        uint8_t v = 4;
        if (ch < 4) v = (uint8_t)ch;
        if (kAscii) {
            switch (ch) {
            case 65: v = 0; break; case 67: v = 1; break; case 71: v = 2; break; case 84: v = 3; break;
            case 82: v = 8; break; case 75: v = 9; break; case 83: v = 10; break;
            case 89: v = 11; break; case 87: v = 12; break; case 77: v = 13; break;
            default: break;
            }
        }
        lut[ch] = v;
    }
}

__device__ __forceinline__ uint32_t resolve_ambiguity(uint32_t cls, uint32_t coin)
{
    // first / second base of R K S Y W M
    const uint32_t first = (0x222331u >> (4 * (5 - cls))) & 0xFu;     // G G G T T C
    const uint32_t second = (0x031100u >> (4 * (5 - cls))) & 0xFu;    // A T C C A A
    return coin ? first : second;
}

// grid.x = ceil(nchunks / 4), grid.y = ceil(ntips / kPackTips); 128 threads, warp = one rank chunk.
// kBoot: columns come from srccol (the replicate's shuffled instance columns) and ambiguity cells are resolved.
template <bool kBoot>
__global__ void __launch_bounds__(128) pack_planes_kernel(PlanDev plan, const uint8_t *__restrict__ arr, int64_t row_stride,
                                                          uint32_t tip_lo, uint32_t tip_hi,
                                                          const uint32_t *__restrict__ srccol, uint32_t seed_ambig,
                                                          uint4 *__restrict__ planes)
{
    __shared__ uint8_t lut[256];
    fill_lut<kBoot>(lut, threadIdx.x, 128);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t nchunks = plan.row_u4 / kTileWords;
    const uint32_t chunk = blockIdx.x * 4 + warp;
    if (chunk >= nchunks) return;
    const int tile = tile_of_chunk(plan, chunk);
    const uint32_t k = chunk - __ldg(plan.tile_off + tile) / kTileWords;
    const uint32_t l0 = (uint32_t)tile * kTileLoci;
    const uint32_t tip0 = tip_lo + blockIdx.y * kPackTips;
    const int nt = (int)min((uint32_t)kPackTips, tip_hi - tip0);
    const uint8_t *rows = arr + (int64_t)tip0 * row_stride;
    uint4 mine[kPackTips];
#pragma unroll
    for (int t = 0; t < kPackTips; ++t) mine[t] = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll 2
    for (int w = 0; w < 32; ++w) {
        const uint32_t l = l0 + 32 * w + lane;
        const bool valid = k < __ldg(plan.loc_len + l);
        uint32_t col = 0, inst = 0;
        if (valid) {
            if (kBoot) { col = __ldg(srccol + (size_t)chunk * kTileLoci + 32 * w + lane); inst = __ldg(plan.loc_run + l); }
            else col = __ldg(plan.loc_start + l) + k;
        }
#pragma unroll
        for (int t = 0; t < kPackTips; ++t) {
            uint32_t code = 4u;
            if (valid && t < nt) {
                code = lut[rows[(int64_t)t * row_stride + col]];
                if (kBoot && code >= 8u) {
                    const uint32_t coin = philox4x32_10(inst, k, tip0 + t, 1u, seed_ambig, 0xA3B1u) & 1u;
                    code = resolve_ambiguity(code - 8u, coin);
                }
            }
            const uint32_t bA = __ballot_sync(kFull, code == 0u), bC = __ballot_sync(kFull, code == 1u);
            const uint32_t bG = __ballot_sync(kFull, code == 2u), bT = __ballot_sync(kFull, code == 3u);
            if (lane == w) mine[t] = make_uint4(bA, bC, bG, bT);
        }
    }
#pragma unroll
    for (int t = 0; t < kPackTips; ++t)
        if (t < nt) planes[(size_t)(tip0 + t) * plan.row_u4 + (size_t)chunk * kTileWords + lane] = mine[t];
}

cudaError_t launch_pack(const PlanDev &plan, const uint8_t *d_tmparr, int64_t tip_lo, int64_t tip_hi,
                        int64_t row_stride, uint32_t *d_planes, cudaStream_t stream)
{
    if (plan.ntiles == 0 || tip_hi <= tip_lo) return cudaSuccess;
    const uint32_t nchunks = plan.row_u4 / kTileWords;
    dim3 grid((nchunks + 3) / 4, (unsigned)((tip_hi - tip_lo + kPackTips - 1) / kPackTips));
    pack_planes_kernel<false><<<grid, 128, 0, stream>>>(plan, d_tmparr, row_stride, (uint32_t)tip_lo, (uint32_t)tip_hi,
                                                        nullptr, 0u, reinterpret_cast<uint4 *>(d_planes));
    ++g_launches;
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// Bootstrap column orders.  Column instance t = loc_pos[sl] + j (j-th column of sorted locus sl, instance
// inst = loc_run[sl]) gets the 64-bit sort key (sl << 32) | philox(inst, j; seed_shuffle) and the value
// src_start[inst] + j.  After a stable radix sort the columns of sl sit at loc_pos[sl] .. in the order (key, j):
// position p holds rank p - loc_pos[sl].  One sort whatever the locus lengths (a scaffold-sized locus costs what its
// columns cost); oracle/boot_oracle.py restates the order in numpy.
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) boot_keys_kernel(PlanDev plan, const uint32_t *__restrict__ src_start,
                                                        uint32_t seed_shuffle, uint64_t *__restrict__ keys,
                                                        uint32_t *__restrict__ vals)
{
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= plan.ncols) return;
    // sorted locus of column instance t: largest sl with loc_pos[sl] <= t
    uint32_t lo = 0, hi = plan.nloci - 1;
    while (lo < hi) {
        const uint32_t mid = (lo + hi + 1) >> 1;
        if (__ldg(plan.loc_pos + mid) <= t) lo = mid; else hi = mid - 1;
    }
    const uint32_t sl = lo, j = t - __ldg(plan.loc_pos + sl);
    const uint32_t inst = __ldg(plan.loc_run + sl);
    keys[t] = ((uint64_t)sl << 32) | philox4x32_10(inst, j, 0u, 0u, seed_shuffle, 0x5EEDu);
    vals[t] = __ldg(src_start + inst) + j;
}

// srccol[chunk * 1024 + locus_in_tile] = source column of rank k of that instance, chunk = tile_off[tile] / 32 + k
__global__ void __launch_bounds__(256) boot_scatter_kernel(PlanDev plan, const uint64_t *__restrict__ keys,
                                                           const uint32_t *__restrict__ vals, uint32_t *__restrict__ srccol)
{
    const uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= plan.ncols) return;
    const uint32_t sl = (uint32_t)(keys[p] >> 32);
    const uint32_t rank = p - __ldg(plan.loc_pos + sl);
    const uint32_t tile = sl / kTileLoci, within = sl % kTileLoci;
    srccol[((size_t)__ldg(plan.tile_off + tile) + (size_t)rank * kTileWords) * kTileWords + within] = vals[p];
}

cudaError_t launch_boot_pack(const PlanDev &plan, const uint8_t *d_seqarr, int64_t tip_lo, int64_t tip_hi, int64_t row_stride,
                             const uint32_t *d_src_start, uint32_t seed_shuffle, uint32_t seed_ambig,
                             uint32_t *d_srccol, uint32_t *d_planes, cudaStream_t stream)
{
    if (plan.ntiles == 0 || plan.ncols == 0) return cudaSuccess;
    const uint32_t n = plan.ncols;
    int sl_bits = 1;
    while ((1ull << sl_bits) < (uint64_t)plan.nloci) ++sl_bits;
    // keys in | keys out | vals in | vals out | cub temp: stream-ordered scratch, returned to the pool right after
    size_t temp_bytes = 0;
    cudaError_t e = cub::DeviceRadixSort::SortPairs(nullptr, temp_bytes, (const uint64_t *)nullptr, (uint64_t *)nullptr,
                                                    (const uint32_t *)nullptr, (uint32_t *)nullptr, (int)n, 0, 32 + sl_bits, stream);
    if (e != cudaSuccess) return e;
    const size_t kb = ((size_t)n * 8 + 255) & ~(size_t)255, vb = ((size_t)n * 4 + 255) & ~(size_t)255;
    char *scratch = nullptr;
    if ((e = cudaMallocAsync(&scratch, 2 * kb + 2 * vb + temp_bytes, stream)) != cudaSuccess) return e;
    uint64_t *keys_in = reinterpret_cast<uint64_t *>(scratch), *keys_out = reinterpret_cast<uint64_t *>(scratch + kb);
    uint32_t *vals_in = reinterpret_cast<uint32_t *>(scratch + 2 * kb), *vals_out = reinterpret_cast<uint32_t *>(scratch + 2 * kb + vb);
    boot_keys_kernel<<<(n + 255) / 256, 256, 0, stream>>>(plan, d_src_start, seed_shuffle, keys_in, vals_in);
    ++g_launches;
    e = cub::DeviceRadixSort::SortPairs(scratch + 2 * kb + 2 * vb, temp_bytes, keys_in, keys_out, vals_in, vals_out, (int)n, 0,
                                        32 + sl_bits, stream);
    g_launches += 8;                                                  // cub: histogram + one pass per 8 key bits (library kernels)
    if (e == cudaSuccess) {
        boot_scatter_kernel<<<(n + 255) / 256, 256, 0, stream>>>(plan, keys_out, vals_out, d_srccol);
        ++g_launches;
        e = cudaGetLastError();
    }
    cudaFreeAsync(scratch, stream);
    if (e != cudaSuccess || tip_hi <= tip_lo) return e;
    const uint32_t nchunks = plan.row_u4 / kTileWords;
    dim3 grid((nchunks + 3) / 4, (unsigned)((tip_hi - tip_lo + kPackTips - 1) / kPackTips));
    pack_planes_kernel<true><<<grid, 128, 0, stream>>>(plan, d_seqarr, row_stride, (uint32_t)tip_lo, (uint32_t)tip_hi, d_srccol,
                                                       seed_ambig, reinterpret_cast<uint4 *>(d_planes));
    ++g_launches;
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------
// Quartet unranking (reference tetrad/src/combinations.py:94-106, _index_to_combination):
// lexicographic rank in itertools.combinations(range(n), 4) -> (a, b, c, d).  The number of
// quartets whose first id is < a is C(n,4) - C(n-a,4), and so on down the levels; each level is
// a binary search on a closed-form binomial (C(n,4) < 2^32 is asserted by the reference).
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t choose2(uint64_t m) { return m < 2 ? 0 : m * (m - 1) / 2; }
__device__ __forceinline__ uint64_t choose3(uint64_t m) { return m < 3 ? 0 : m * (m - 1) / 2 * (m - 2) / 3; }
__device__ __forceinline__ uint64_t choose4(uint64_t m) { return m < 4 ? 0 : m * (m - 1) / 2 * (m - 2) / 3 * (m - 3) / 4; }

// smallest m in [lo, hi] with f(m) >= x  (f non-decreasing)
template <typename F>
__device__ __forceinline__ uint32_t lower_bound_fn(uint32_t lo, uint32_t hi, uint64_t x, F f)
{
    while (lo < hi) {
        const uint32_t mid = (lo + hi) >> 1;
        if (f(mid) >= x) hi = mid; else lo = mid + 1;
    }
    return lo;
}

__global__ void __launch_bounds__(256) unrank_kernel(const int64_t *__restrict__ ranks, int64_t start, int64_t m,
                                                     uint32_t n, uint4 *__restrict__ quartets)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    uint64_t r = (uint64_t)(ranks ? ranks[i] : start + i);
    const uint64_t total = choose4(n);
    // first id a: largest a with total - C(n-a,4) <= r  <=>  smallest m1 = n-a with C(m1,4) >= total - r
    const uint32_t m1 = lower_bound_fn(4u, n, total - r, [](uint32_t mm) { return choose4(mm); });
    const uint32_t a = n - m1;
    r -= total - choose4(m1);
    const uint64_t x = choose3(m1 - 1) - r;                         // >= 1
    const uint32_t m2 = lower_bound_fn(3u, m1 - 1, x, [](uint32_t mm) { return choose3(mm); });
    const uint32_t b = n - m2;
    r -= choose3(m1 - 1) - choose3(m2);
    const uint64_t y = choose2(m2 - 1) - r;
    const uint32_t m3 = lower_bound_fn(2u, m2 - 1, y, [](uint32_t mm) { return choose2(mm); });
    const uint32_t c = n - m3;
    r -= choose2(m2 - 1) - choose2(m3);
    quartets[i] = make_uint4(a, b, c, c + 1 + (uint32_t)r);
}

// ------------------------------------------------------------------------------------------
// Random quartet sample without a host draw.  The reference draws rng.choice(C(n,4), Q, replace=False) on the host
// (tetrad/src/combinations.py:109-114) -- 0.13 s per million quartets of 64.7 M, serial, which is what an 8-GPU
// replicate loop would wait for.  Here quartet i of a replicate's sample is unrank(pi_key(i)), pi_key a keyed
// pseudo-random PERMUTATION of [0, C(n,4)): a 6-round Feistel network on ceil(log2 C(n,4)) (rounded up to even) bits,
// cycle-walked into the range.  Distinct i give distinct quartets by construction, the first Q images of a random
// permutation are a uniform random Q-subset in random order, any rank computes any index range on its own, and nothing
// is uploaded.  oracle/sample_oracle.py restates pi in numpy (bit-exact test + uniformity test).
// ------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint32_t fmix32(uint32_t h)
{
    h ^= h >> 16; h *= 0x85EBCA6Bu; h ^= h >> 13; h *= 0xC2B2AE35u; h ^= h >> 16;
    return h;
}

__host__ __device__ __forceinline__ uint64_t feistel_permute(uint64_t x, uint64_t total, uint32_t half, uint32_t k0, uint32_t k1)
{
    const uint32_t mask = (uint32_t)((1ull << half) - 1ull);
    do {
        uint32_t L = (uint32_t)(x >> half), R = (uint32_t)x & mask;
#pragma unroll
        for (uint32_t r = 0; r < 6; ++r) {
            const uint32_t f = fmix32(fmix32(R ^ k0 ^ (r * 0x9E3779B9u)) + k1) & mask;
            const uint32_t nr = L ^ f;
            L = R; R = nr;
        }
        x = ((uint64_t)L << half) | R;
    } while (x >= total);
    return x;
}

__device__ __forceinline__ uint4 unrank_quartet(uint64_t r, uint32_t n)
{
    const uint64_t total = choose4(n);
    const uint32_t m1 = lower_bound_fn(4u, n, total - r, [](uint32_t mm) { return choose4(mm); });
    const uint32_t a = n - m1;
    r -= total - choose4(m1);
    const uint64_t x = choose3(m1 - 1) - r;
    const uint32_t m2 = lower_bound_fn(3u, m1 - 1, x, [](uint32_t mm) { return choose3(mm); });
    const uint32_t b = n - m2;
    r -= choose3(m1 - 1) - choose3(m2);
    const uint64_t y = choose2(m2 - 1) - r;
    const uint32_t m3 = lower_bound_fn(2u, m2 - 1, y, [](uint32_t mm) { return choose2(mm); });
    const uint32_t c = n - m3;
    r -= choose2(m2 - 1) - choose2(m3);
    return make_uint4(a, b, c, c + 1 + (uint32_t)r);
}

__global__ void __launch_bounds__(256) sample_unrank_kernel(int64_t start, int64_t m, uint32_t n, uint32_t half, uint32_t k0,
                                                            uint32_t k1, uint4 *__restrict__ quartets)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    quartets[i] = unrank_quartet(feistel_permute((uint64_t)(start + i), choose4(n), half, k0, k1), n);
}

// host twin of the permutation (CPU tests of the sampler; no kernel)
void sample_ranks_host(int64_t start, int64_t m, uint32_t ntips, uint64_t key, int64_t *out)
{
    const uint64_t n = ntips, total = n * (n - 1) / 2 * (n - 2) / 3 * (n - 3) / 4;
    uint32_t bits = 2;
    while ((1ull << bits) < total) ++bits;
    bits += bits & 1u;
    for (int64_t i = 0; i < m; ++i)
        out[i] = (int64_t)feistel_permute((uint64_t)(start + i), total, bits / 2, (uint32_t)key, (uint32_t)(key >> 32));
}

cudaError_t launch_sample_unrank(int64_t start, int64_t m, uint32_t ntips, uint64_t key, uint32_t *d_quartets, cudaStream_t stream)
{
    if (m == 0) return cudaSuccess;
    const uint64_t n = ntips, total = n * (n - 1) / 2 * (n - 2) / 3 * (n - 3) / 4;
    uint32_t bits = 2;
    while ((1ull << bits) < total) ++bits;
    bits += bits & 1u;
    sample_unrank_kernel<<<(unsigned)((m + 255) / 256), 256, 0, stream>>>(start, m, ntips, bits / 2, (uint32_t)key, (uint32_t)(key >> 32),
                                                                          reinterpret_cast<uint4 *>(d_quartets));
    ++g_launches;
    return cudaGetLastError();
}

cudaError_t launch_unrank(const int64_t *d_ranks, int64_t start, int64_t m, uint32_t ntips, uint32_t *d_quartets,
                          cudaStream_t stream)
{
    if (m == 0) return cudaSuccess;
    unrank_kernel<<<(unsigned)((m + 255) / 256), 256, 0, stream>>>(d_ranks, start, m, ntips,
                                                                   reinterpret_cast<uint4 *>(d_quartets));
    ++g_launches;
    return cudaGetLastError();
}

}  // namespace tq
