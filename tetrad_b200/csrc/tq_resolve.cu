// Resolve kernels of the quartet-resolution hot path (sm_100a): fused count + score, only the result leaves the SM.
//
//   resolve_mma_kernel     the kernel large launches run.  16 warps per CTA, one CTA per SM: 8 count + 8 score warps or
//                          12 + 4, chosen from the plan (pick_count_warps).  Count warps: stage 1 selects one SNP per
//                          locus (running OR over the rank chunks of the four tip rows; in subsample mode on
//                          {hi, lo, valid} code planes converted per launch from the one-hot matrix: 18 LOP3 per chunk),
//                          stacks the merged code planes of consecutive sparse tiles in registers (a locus whose bit
//                          slot is taken is counted by a shared-memory atomic), and stage 2 forms the 256 site-pattern
//                          counts of a stack of tiles on the integer tensor cores (IMMA.16832.U8.U8 on bit planes
//                          unpacked by one LOP3 per four loci); the tip rows reach shared memory through a per-warp
//                          cp.async ring.  Score warps take four finished quartets at a time through the FP64 QR +
//                          Jacobi of tq_svd.cuh.  (A CTA-wide ring of TMA bulk copies that
//                          holds every distinct row of eight consecutive quartets once was built and measured in round 2:
//                          correct, but half as fast -- sharing one ring lock-steps the count warps, so their IMMA and load
//                          phases stop overlapping, and four stages do not cover a bulk copy's latency;
//                          profiles/dropped/r02_shared_tma_ring.md.)
//   resolve_kernel         one warp per quartet, counts by AND + POPC with the packed counters in tensor memory
//                          (tcgen05.alloc / ld / st), scores by the same tq_svd.cuh code: what small launches run (a
//                          count -> score hand-off is pure latency when there are fewer quartets than warps).
//   (full mode, dense quartet lists: tq_pairgemm.cu -- tcgen05.mma over tip pairs)
//
// launch_resolve picks by launch size; tq_set_option("kernel", ..) / TQ_KERNEL force one.  All give bit-identical results.
// The measurements behind the constants (warp mixes, ring depth 4, L1-bypassing cp.async, four mask scales, stack_ovf, ...)
// are in profiles/r01_v6_resolve_ncu.md, profiles/r01_v7_resolve_ncu.md and profiles/r02_warp_mix.md; the build switches
// that produced the round-1 ones are in the history of this file's predecessor (tq_kernels.cu, commit 425353e), not in
// the product.
//
// Reference semantics restated here (paths relative to the reference repo):
//   tetrad/src/resolve_quartets.py:212-221  gather, missing mask, invariant mask
//   tetrad/src/resolve_quartets.py:42-73    subsample_chunk_to_matrices (first usable SNP/locus)
//   tetrad/src/resolve_quartets.py:76-104   full_chunk_to_matrices
//   tetrad/src/resolve_quartets.py:226-264  nsnps, svd, matrix_rank, minrank, score, argmin
//
// Data layout: locus runs are sorted by length (longest first) and cut into tiles of 1024 loci; a tile
// with K ranks stores, for every tip, K x 32 uint4 {A,C,G,T}: word `lane` of rank k holds, at bit b, the k-th column of
// sorted locus 1024*tile+32*lane+b.  "First usable column of a locus" (the reference's last_loc state machine) is then
// "lowest rank k whose bit survives the quartet's mask" -- a running OR per lane.
#include "tq_kernels.cuh"
#include "tq_device.cuh"
#include "tq_svd.cuh"

#include <cstdlib>
#include <cstring>
#include <mutex>

namespace tq {

std::atomic<int64_t> g_launches{0};
std::atomic<const char *> g_last_kernel{"none"};
std::atomic<int> g_last_count_warps{0};          // count warps per CTA of the last tensor-core launch
// Kernel choice and test hooks (tq_set_option in the C ABI; the environment gives the initial values).
//   g_kernel_choice: 0 = by launch size and mode, 1 = fused (one warp per quartet, POPC), 3 = tensor-core count,
//                    5 = tip-pair GEMM (full mode only; subsample launches run 3)
static int initial_kernel_choice()
{
    const char *e = getenv("TQ_KERNEL");
    if (!e) return 0;
    return !strcmp(e, "fused") ? 1 : !strcmp(e, "mma") ? 3 : !strcmp(e, "pairgemm") ? 5 : 0;
}
static int64_t initial_mma_flush()
{
    const char *e = getenv("TQ_MMA_FLUSH");
    return e ? atol(e) : 0;
}
static int initial_count_warps()
{
    const char *e = getenv("TQ_COUNT_WARPS");
    return e ? atoi(e) : 0;
}
std::atomic<int> g_kernel_choice{initial_kernel_choice()};
std::atomic<int64_t> g_mma_flush{initial_mma_flush()};
static int initial_stack_fill()
{
    const char *e = getenv("TQ_STACK_FILL");
    return e ? atoi(e) : -1;
}
std::atomic<int> g_stack_fill{initial_stack_fill()};     // < 0: default; else percent of a tile's 1024 bit slots (0: never stack)
static int initial_stack_ovf()
{
    const char *e = getenv("TQ_STACK_OVF");
    return e ? atoi(e) : 8;
}
std::atomic<int> g_stack_ovf{initial_stack_ovf()};
std::atomic<int> g_stack_stats{0};                       // != 0: the kernel adds its stacking diagnostics to status[1..3]       // most overflow loci a lane may have for a tile to be stacked
std::atomic<int> g_count_warps{initial_count_warps()};

// Bit slots of P that must be filled before the stacked tiles' product is taken (count_quartet_mma).
static constexpr int kStackFillPercent = 100;
static uint32_t stack_fill_slots()
{
    const int pct = g_stack_fill.load();
    return (uint32_t)((pct < 0 ? kStackFillPercent : pct > 100 ? 100 : pct) * kTileLoci / 100) |
           (g_stack_stats.load() ? 0x80000000u : 0u);
}
   // 0: by the plan; else the count warps per CTA of the tensor-core kernel

static constexpr unsigned kFull = 0xFFFFFFFFu;
static constexpr int kWarps = 4;                 // warps (= quartets in flight) per CTA of the fused kernel
static constexpr int kStages = 4;                // cp.async ring depth (rank chunks in flight per warp)
static constexpr uint32_t kFlushEvery = 63;      // POPC kernel: tile words between flushes of the packed 16-bit counters
static constexpr uint32_t kTmemCols = 64;        // POPC kernel: tensor-memory columns per CTA (64 packed counters per thread)

// cudaFuncSetAttribute is per device and must have RETURNED before any thread launches with > 48 KB of dynamic shared
// memory: one mutex-guarded bit per device and kernel, set only after success.
struct SmemAttr {
    std::mutex mu;
    uint64_t done = 0;
    template <typename K>
    cudaError_t ensure(K kern, size_t bytes)
    {
        int dev = 0;
        cudaGetDevice(&dev);
        const uint64_t bit = 1ull << (dev & 63);
        std::lock_guard<std::mutex> lk(mu);
        if (done & bit) return cudaSuccess;
        const cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
        if (e == cudaSuccess) done |= bit;
        return e;
    }
};

// ==========================================================================================
// Small launches: one warp per quartet, AND + POPC counts with the counters in tensor memory
// ==========================================================================================
// Lane l owns locus word l of every tile (32 loci).  stage 1 as described above; stage 2: the 256 pattern counts of a
// tile are popc(A_a & B_b & C_c & D_d) on the merged planes.  To halve the counters the two half-warps split the
// patterns: lanes 0-15 count a in {0,1}, lanes 16-31 a in {2,3}, each for its own word and for the word of its partner
// lane (l ^ 16), whose planes arrive by 14 warp shuffles.  Counter i packs two 16-bit counts: patterns (a, b, cd=2h) and
// (a, b, cd=2h+1) with i = ((a & 1) * 4 + b) * 8 + h.  A lane adds at most 64 per call, the 16-lane butterfly of the flush
// multiplies by 16, so flushing every 63 calls keeps every field below 65536.
// Tensor memory as the counter file: a CTA (4 warps) allocates 64 columns; warp w owns TMEM lanes 32w..32w+31 of them,
// i.e. every thread has 64 private 32-bit cells, which therefore do not occupy registers across tiles.
struct __align__(16) WarpScratch {
    union {
        uint4 ring[kStages][4][32];             // stage-1 staging: [stage][tip][lane]      8 KB
        SvdScratch4 svd;                        // the Jacobi reuses the ring once counting is over
    };
    uint32_t T[256];
};

// One (a,b) row = 16 patterns.  The ANDs and the POPCs are volatile asm so that ptxas keeps them in the order written:
// a software pipeline in which every POPC has ~16 independent POPCs and the next row's ANDs between its issue and the add
// that consumes it (left alone, ptxas puts the consumer 1-3 instructions behind the POPC and the warp sits on the XU
// latency).
__device__ __forceinline__ void row_and(uint32_t ab, const uint32_t (&C)[4], const uint32_t (&D)[4], uint32_t (&v)[16])
{
#pragma unroll
    for (int cd = 0; cd < 16; ++cd)
        asm volatile("lop3.b32 %0, %1, %2, %3, 0x80;" : "=r"(v[cd]) : "r"(ab), "r"(C[cd >> 2]), "r"(D[cd & 3]));
}
__device__ __forceinline__ void row_popc(uint32_t (&v)[16])
{
#pragma unroll
    for (int cd = 0; cd < 16; ++cd) asm volatile("popc.b32 %0, %0;" : "+r"(v[cd]));
}
__device__ __forceinline__ void row_add(const uint32_t (&v)[16], uint32_t *acc8)
{
#pragma unroll
    for (int h = 0; h < 8; ++h) acc8[h] += v[2 * h] + (v[2 * h + 1] << 16);
}

// 256 pattern counts of one tile word pair (own word + partner lane's word), split by half-warp.
__device__ __forceinline__ void count_tile(const uint32_t (&MA)[4], const uint32_t (&MB)[4],
                                           const uint32_t (&MC)[4], const uint32_t (&MD)[4],
                                           bool upper, uint32_t tacc)
{
    // planes the partner needs from me: the a-planes of ITS half, and all of B, C, D
    const uint32_t sendA0 = upper ? MA[0] : MA[2], sendA1 = upper ? MA[1] : MA[3];
    const uint32_t pA0 = __shfl_xor_sync(kFull, sendA0, 16), pA1 = __shfl_xor_sync(kFull, sendA1, 16);
    uint32_t pB[4], pC[4], pD[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        pB[i] = __shfl_xor_sync(kFull, MB[i], 16);
        pC[i] = __shfl_xor_sync(kFull, MC[i], 16);
        pD[i] = __shfl_xor_sync(kFull, MD[i], 16);
    }
    const uint32_t A0 = upper ? MA[2] : MA[0], A1 = upper ? MA[3] : MA[1];
    tmem_wait_st();                                  // the previous tile's stores have landed
    // counters 16g .. 16g+15 hold rows (ai, b) = (g / 2, 2 (g % 2) + {0, 1}).  A runtime loop: the counters are addressed
    // in tensor memory, so nothing forces the 256 patterns to be unrolled, and the whole body (64 AND + 64 POPC + 64 add)
    // stays resident in the instruction cache.
#pragma unroll 1
    for (int g = 0; g < 4; ++g) {
        uint32_t acc[16];
        tmem_ld16(tacc + 16 * g, acc);
        const bool hi = (g & 1) != 0;
        const uint32_t a_own = (g >> 1) ? A1 : A0, a_par = (g >> 1) ? pA1 : pA0;
        const uint32_t o0 = a_own & (hi ? MB[2] : MB[0]), o1 = a_own & (hi ? MB[3] : MB[1]);
        const uint32_t q0 = a_par & (hi ? pB[2] : pB[0]), q1 = a_par & (hi ? pB[3] : pB[1]);
        uint32_t va[16], vb[16];
        row_and(o0, MC, MD, va);
        row_popc(va);
        row_and(o1, MC, MD, vb);
        tmem_wait_ld();
        row_add(va, acc);
        row_popc(vb);
        row_and(q0, pC, pD, va);
        row_add(vb, acc + 8);
        row_popc(va);
        row_and(q1, pC, pD, vb);
        row_add(va, acc);
        row_popc(vb);
        row_add(vb, acc + 8);
        tmem_st16(tacc + 16 * g, acc);
    }
}

__device__ __forceinline__ void zero_acc(uint32_t tacc)
{
    uint32_t z[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) z[i] = 0u;
    tmem_wait_st();
#pragma unroll
    for (int g = 0; g < 4; ++g) tmem_st16(tacc + 16 * g, z);
}

// Read the 64 packed counters back from tensor memory, butterfly-add them over the 16 lanes of each half-warp, add into
// T, and clear them.  Runtime loop over the four groups (code size).
__device__ __noinline__ void flush_tmem(uint32_t tacc, uint32_t *T, int lane)
{
    const int abase = (lane >> 4) * 2;           // first a-value of this half-warp
    tmem_wait_st();
#pragma unroll 1
    for (int g = 0; g < 4; ++g) {
        uint32_t part[16];
        tmem_ld16(tacc + 16 * g, part);
        tmem_wait_ld();
        uint32_t *row = T + ((abase + (g >> 1)) * 4 + 2 * (g & 1)) * 16;      // pattern (a, b = 2(g%2), cd = 0)
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            uint32_t v = part[k];
            v += __shfl_xor_sync(kFull, v, 8);
            v += __shfl_xor_sync(kFull, v, 4);
            v += __shfl_xor_sync(kFull, v, 2);
            v += __shfl_xor_sync(kFull, v, 1);
            if ((lane & 15) == k) {                                            // k = 8 (b & 1) + h
                uint2 *t = reinterpret_cast<uint2 *>(row + (k >> 3) * 16 + 2 * (k & 7));
                uint2 cur = *t;
                cur.x += v & 0xFFFFu;
                cur.y += v >> 16;
                *t = cur;
            }
            part[k] = 0u;
        }
        tmem_st16(tacc + 16 * g, part);
    }
}

// Issue the four 16-byte cp.async of rank chunk `nu` (one per tip) into ring stage nu % kStages.
// All global addressing is a 32-bit uint4 index off `planes` (the host checks ntips*row_u4 < 2^32).
__device__ __forceinline__ void issue_chunk(uint32_t ring_s, const uint4 *__restrict__ planes, uint32_t lane_nu,
                                            uint32_t stage, uint32_t t0, uint32_t t1, uint32_t t2, uint32_t t3)
{
    const uint32_t s = ring_s + stage * (4 * 32 * 16);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(planes + (t0 + lane_nu)) : "memory");
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n" ::"r"(s + 512), "l"(planes + (t1 + lane_nu)) : "memory");
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n" ::"r"(s + 1024), "l"(planes + (t2 + lane_nu)) : "memory");
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n" ::"r"(s + 1536), "l"(planes + (t3 + lane_nu)) : "memory");
}

template <bool kSubsample>
__device__ __forceinline__ void count_quartet(const PlanDev &plan, const uint4 *__restrict__ planes,
                                              uint4 tips, const uint4 *ring, uint32_t *T, int lane, uint32_t tacc)
{
    const bool upper = lane >= 16;
    uint32_t pending = 0;
    const uint32_t nchunks = plan.row_u4 / kTileWords;
    const uint32_t t0 = tips.x * plan.row_u4, t1 = tips.y * plan.row_u4;    // row starts, uint4 units
    const uint32_t t2 = tips.z * plan.row_u4, t3 = tips.w * plan.row_u4;
    const uint32_t ring_s = smem_u32(ring + lane);

    // prologue: chunks 0 .. kStages-2 in flight
#pragma unroll
    for (int s = 0; s < kStages - 1; ++s) {
        if ((uint32_t)s < nchunks) issue_chunk(ring_s, planes, s * kTileWords + lane, s, t0, t1, t2, t3);
        cp_async_commit();
    }

    uint32_t u = 0;                                   // rank-chunk index along the rows
    for (int tile = 0; tile < plan.ntiles; ++tile) {
        const uint32_t K = __ldg(plan.tile_k + tile);
        uint32_t taken = 0u;
        uint32_t MA[4] = {0u, 0u, 0u, 0u}, MB[4] = {0u, 0u, 0u, 0u};
        uint32_t MC[4] = {0u, 0u, 0u, 0u}, MD[4] = {0u, 0u, 0u, 0u};
        for (uint32_t k = 0; k < K; ++k, ++u) {
            {   // keep kStages-1 chunks in flight
                const uint32_t nu = u + kStages - 1;
                if (nu < nchunks) issue_chunk(ring_s, planes, nu * kTileWords + lane, nu % kStages, t0, t1, t2, t3);
                cp_async_commit();
            }
            cp_async_wait<kStages - 1>();
            const int s = u % kStages;
            const uint4 *st = ring + s * (4 * 32) + lane;
            const uint4 a = st[0], b = st[32], c = st[64], d = st[96];
            if (kSubsample) {
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
            } else {
                // full mode (resolve_quartets.py:92-95): every complete column counts; the four
                // invariant patterns aaaa are zeroed by the caller
                const uint32_t A[4] = {a.x, a.y, a.z, a.w}, B[4] = {b.x, b.y, b.z, b.w};
                const uint32_t C[4] = {c.x, c.y, c.z, c.w}, D[4] = {d.x, d.y, d.z, d.w};
                count_tile(A, B, C, D, upper, tacc);
                if (++pending == kFlushEvery) { flush_tmem(tacc, T, lane); pending = 0; }
            }
        }
        if (kSubsample) {
            count_tile(MA, MB, MC, MD, upper, tacc);
            if (++pending == kFlushEvery) { flush_tmem(tacc, T, lane); pending = 0; }
        }
    }
    cp_async_wait<0>();
    flush_tmem(tacc, T, lane);
}

// What every kernel does with a finished count tensor T[256] of a quartet whose tips are in range: zero the invariant
// patterns in full mode (:218), nsnps (:226), optional export; returns nsnps (warp-uniform).
template <bool kSubsample>
__device__ __forceinline__ uint32_t finish_counts(uint32_t *T, int lane, uint32_t *__restrict__ counts, long long q)
{
    if (!kSubsample && lane < 4) T[85u * lane] = 0u;   // invariant patterns aaaa, :218
    __syncwarp();
    uint32_t nsnps = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const uint32_t v = T[lane + 32 * i];
        nsnps += v;
        if (counts) counts[q * 256 + lane + 32 * i] = v;
    }
    return __reduce_add_sync(kFull, nsnps);                             // :226
}

template <bool kSubsample>
__global__ void __launch_bounds__(kWarps * 32, 4)
resolve_kernel(PlanDev plan, const uint4 *__restrict__ planes, uint32_t ntips,
               const uint4 *__restrict__ quartets, int64_t m, uint64_t seed,
               double *__restrict__ rscor, uint2 *__restrict__ rstat,
               uint32_t *__restrict__ counts, uint32_t *__restrict__ minrank_out,
               uint32_t *__restrict__ status)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    WarpScratch *scratch = reinterpret_cast<WarpScratch *>(smem_raw);
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    WarpScratch &ws = scratch[warp];
    const int64_t nwarps = (int64_t)gridDim.x * kWarps;
    __shared__ uint32_t tmem_base_slot;
    if (warp == 0) tmem_alloc(&tmem_base_slot, kTmemCols);   // one warp allocates (and later frees) the CTA's columns
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = tmem_base_slot;
    const uint32_t tacc = tmem_base + ((uint32_t)(warp * 32) << 16);   // this warp's lane quadrant
    zero_acc(tacc);

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
        count_quartet<kSubsample>(plan, planes, t, &ws.ring[0][0][0], ws.T, lane, tacc);
        __syncwarp();
        const uint32_t nsnps = finish_counts<kSubsample>(ws.T, lane, counts, q);
        double score[3] = {0.001, 0.001, 0.001};                        // :232
        uint32_t best = 0, minrank = 0;
        if (nsnps == 0) best = tiebreak_split(seed, t);                 // stands in for :231
        else svd_scores4(ws.T, 0u, 1u, ws.svd, lane, score, best, minrank);   // one quartet: pass 0, matrices 0..2
        if (lane == 0) {
            rscor[3 * q] = score[0]; rscor[3 * q + 1] = score[1]; rscor[3 * q + 2] = score[2];
            rstat[q] = make_uint2(best, nsnps);                         // :251, :264
            if (minrank_out) minrank_out[q] = minrank;
        }
        __syncwarp();
    }
    tmem_wait_st();
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, kTmemCols);
}

// ==========================================================================================
// Large launches: stage 2 on the integer tensor cores
// ==========================================================================================
//   T[ab][cd] = sum over the tile's 1024 loci of  X_ab[l] * Y_cd[l],   X_ab = A_a & B_b,  Y_cd = C_c & D_d
//
// is a 16 x 16 x 1024 product of 0/1 matrices whose rows ARE the merged bit planes.  sm_100a has no native 1-bit MMA
// (ptxas expands mma.b1 into IMMA + ~100 unpack instructions), but the legacy u8 instruction mma.sync.m16n8k32 (SASS
// IMMA.16832.U8.U8) is native, and a bit plane becomes an u8 operand with ONE LOP3 per four loci: word & (0x01010101 << v)
// holds the bytes {0, 2^v}.  Both operands use the same mask, so all products of one instruction are {0, 4^v}; v = 0..3 on
// the word and on the word >> 4 cover the eight bits of every byte, and each v has its own s32 accumulator fragment:
// count = sum_v acc_v >> 2v.  The K order of an MMA is free as long as A and B agree, so lane (g = lane / 4, j = lane % 4)
// simply feeds bytes j of ITS OWN word pair -- the data is never transposed: the lane reads planes A_p, A_p+2, B_r (rows g
// and g + 8 of the A operand, p = g / 4, r = g % 4) and C_p, C_p+2, D_r (column g of the two B operands) of words 8 j ..
// 8 j + 7 from a per-warp shared-memory copy of the tile's 16 merged planes (row stride 34 words: conflict-free 8-byte
// reads).  Per tile and warp: 16 STS + 24 LDS.64 + ~320 LOP3/SHF + 64 IMMA; nothing runs on the XU pipe, the MMA sums over
// the whole tile (no cross-lane butterfly), and the counters are 32 registers.  (Why not tcgen05 here: the operands are
// per quartet -- the selected column of a locus depends on all four tips -- and 16 x 16 x 1024 is far below a tcgen05 tile;
// batching eight quartets into one 128 x 128 tile computes 64 blocks to use 8.  Full mode, where the operands belong to
// tip PAIRS, does use it: tq_pairgemm.cu.)
static constexpr int kMStages = 4;                  // ring depth (per-warp cp.async ring and CTA-shared TMA ring)
static constexpr int kScales = 4;
static constexpr int kPStride = 34;                 // words per plane row of the per-warp tile copy
static constexpr uint32_t kMmaFlushEvery = 16384;   // tile products between accumulator flushes: 16384 * 1024 * 4^3 < 2^31
                                                    // (tq_set_option("mma_flush_every") lowers it so that tests reach the flush)
static constexpr int kMmaWarps = 16;                // warps per CTA: 16 x 32 threads x 128 registers = the SM's register file
// Count / score warps per CTA.  The FP64 scores cost the same for every quartet while the count work grows with the tiles
// and rank chunks of a tip row, so the mix follows the plan (pick_count_warps): 8 + 8 at the C2 shape (49 tiles: the two
// sides balance, profiles/r01_v7_resolve_ncu.md), 12 + 4 where a row has a hundred tiles or more and eight score warps
// would idle on half of the registers (C3 / C4 / C5: +7 .. +10 %).  Warp w runs on scheduler w % 4, so only multiples of
// four keep the four schedulers evenly loaded: 10 + 6, 11 + 5, 13 + 3 and 14 + 2 all measured slower than both
// neighbours, and 16 homogeneous warps (count, then score the same quartet) slower than 12 + 4 (profiles/r02_warp_mix.md).
static constexpr int kMixes[] = {8, 12};            // count warps of the instantiated kernels (score warps: the rest)

__device__ __forceinline__ void imma_u8(int (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1)
{
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.u8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// acc[v][h][i]: scale 4^v, column half h (cd = 8h ..), fragment register i
struct MmaLanePtrs {              // per-lane views of the warp's tile copy P, set up once per quartet
    uint32_t *st;                 // P + lane: where this lane stores its word of every plane
    const uint32_t *pa;           // A_p words 8j..: A_p+2 two rows on, C_p eight rows on, C_p+2 ten rows on
    const uint32_t *pb;           // B_r words 8j..: D_r eight rows on
};
__device__ __forceinline__ MmaLanePtrs mma_lane_ptrs(uint32_t *P, int lane)
{
    const int g = lane >> 2, j = lane & 3;
    MmaLanePtrs lp;
    lp.st = P + lane;
    lp.pa = P + (g >> 2) * kPStride + 8 * j;
    lp.pb = P + (4 + (g & 3)) * kPStride + 8 * j;
    return lp;
}

// This lane's word of the 16 planes -> the warp's tile copy P (plain store: whatever P held is replaced).
__device__ __forceinline__ void store_tile_planes(const uint32_t (&MA)[4], const uint32_t (&MB)[4],
                                                  const uint32_t (&MC)[4], const uint32_t (&MD)[4], const MmaLanePtrs &lp)
{
    __syncwarp();                                    // the previous tile product's reads are over
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        lp.st[(0 + k) * kPStride] = MA[k];
        lp.st[(4 + k) * kPStride] = MB[k];
        lp.st[(8 + k) * kPStride] = MC[k];
        lp.st[(12 + k) * kPStride] = MD[k];
    }
}

// The 16 x 16 x 1024 product of the planes in P, added to the accumulators.
__device__ __forceinline__ void tile_product_mma(const MmaLanePtrs &lp, int (&acc)[kScales][2][4])
{
    __syncwarp();                                    // every lane's stores to P are visible
    const uint32_t *pa = lp.pa, *pb = lp.pb;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const uint2 a0 = *reinterpret_cast<const uint2 *>(pa + 2 * i);
        const uint2 a1 = *reinterpret_cast<const uint2 *>(pa + 2 * kPStride + 2 * i);
        const uint2 b = *reinterpret_cast<const uint2 *>(pb + 2 * i);
        const uint2 c0 = *reinterpret_cast<const uint2 *>(pa + 8 * kPStride + 2 * i);
        const uint2 c1 = *reinterpret_cast<const uint2 *>(pa + 10 * kPStride + 2 * i);
        const uint2 d = *reinterpret_cast<const uint2 *>(pb + 8 * kPStride + 2 * i);
        uint32_t x[2][4], y[2][4];    // x: rows g / g+8 of word W, then of word W';  y: column g of half 0 (W, W'), of half 1 (W, W')
        x[0][0] = a0.x & b.x; x[0][1] = a1.x & b.x; x[0][2] = a0.y & b.y; x[0][3] = a1.y & b.y;
        y[0][0] = c0.x & d.x; y[0][1] = c0.y & d.y; y[0][2] = c1.x & d.x; y[0][3] = c1.y & d.y;
#pragma unroll
        for (int q = 0; q < 4; ++q) { x[1][q] = x[0][q] >> 4; y[1][q] = y[0][q] >> 4; }
#pragma unroll
        for (int s = 0; s < 2; ++s)
#pragma unroll
            for (int v = 0; v < 4; ++v) {
                const uint32_t msk = 0x01010101u << v;
                const uint32_t f0 = x[s][0] & msk, f1 = x[s][1] & msk, f2 = x[s][2] & msk, f3 = x[s][3] & msk;
                imma_u8(acc[v][0], f0, f1, f2, f3, y[s][0] & msk, y[s][1] & msk);
                imma_u8(acc[v][1], f0, f1, f2, f3, y[s][2] & msk, y[s][3] & msk);
            }
    }
}

__device__ __forceinline__ void count_tile_mma(const uint32_t (&MA)[4], const uint32_t (&MB)[4],
                                               const uint32_t (&MC)[4], const uint32_t (&MD)[4],
                                               const MmaLanePtrs &lp, int (&acc)[kScales][2][4])
{
    store_tile_planes(MA, MB, MC, MD, lp);
    tile_product_mma(lp, acc);
}

// ---- stacking sparse tiles (subsample mode) ------------------------------------------------------------------------
// A tile product costs the same whether 1 024 or 100 of the tile's loci selected a column, and on real and synthetic
// data most loci select none for a given quartet (a column must be complete in all four tips and segregating: 24 - 34 % of
// the loci at the BASELINE shapes).  Bit positions carry no meaning once the planes are merged, so the merged code planes
// of SEVERAL tiles are stacked -- in registers -- before one product is taken: a selected locus whose bit slot is still free
// (`put`) is OR-ed into the stack; one whose slot is taken (`ovf`) is counted on the spot by a shared-memory atomic on its
// pattern index.  The price of stacking a tile is known before it is paid (the busiest lane's number of such loci), so
// dense data behaves as before (one product per tile, no overflow) and sparse data takes a fraction of the products.
// Every selected column is counted exactly once, by the product or by the atomic: the counts stay bit-exact.
//
// Count the selected loci in `ovf` (bit slots already taken in P) one by one: pattern index ((a*4+b)*4+c)*4+d from the
// tile's merged code planes (A, C, G, T = 0..3: H = high bit, L = low bit, one word per tip), one shared-memory atomic each.
// The count warps pay for every instruction of their own streams, so the eight code bits of a locus are gathered with two
// instructions each: the plane word shifted right by the locus' position, then one funnel shift that pushes its lowest bit
// into the top of the index under construction (first plane pushed ends up lowest).  Two loci per trip -- the highest and
// the lowest set bit -- halve the trips of the lane with the most overflow.
__device__ __forceinline__ uint32_t code_index(uint32_t pos, uint32_t ha, uint32_t la, uint32_t hb, uint32_t lb,
                                               uint32_t hc, uint32_t lc, uint32_t hd, uint32_t ld)
{
    uint32_t idx = 0u;                               // ((a*4+b)*4+c)*4+d: ld is bit 0, ha bit 7
    idx = __funnelshift_r(idx, ld >> pos, 1); idx = __funnelshift_r(idx, hd >> pos, 1);
    idx = __funnelshift_r(idx, lc >> pos, 1); idx = __funnelshift_r(idx, hc >> pos, 1);
    idx = __funnelshift_r(idx, lb >> pos, 1); idx = __funnelshift_r(idx, hb >> pos, 1);
    idx = __funnelshift_r(idx, la >> pos, 1); idx = __funnelshift_r(idx, ha >> pos, 1);
    return idx >> 24;
}
__device__ __forceinline__ void scatter_overflow(uint32_t ovf, const uint32_t (&H)[4], const uint32_t (&L)[4], uint32_t *T)
{
    const uint32_t ha = H[0], la = L[0], hb = H[1], lb = L[1], hc = H[2], lc = L[2], hd = H[3], ld = L[3];
    while (ovf) {
        const uint32_t pos = 31u - __clz(ovf);
        ovf ^= 1u << pos;
        const bool second = ovf != 0u;
        const uint32_t pos2 = second ? (uint32_t)__ffs(ovf) - 1u : pos;
        ovf &= ovf - 1u;
        const uint32_t idx = code_index(pos, ha, la, hb, lb, hc, lc, hd, ld);
        const uint32_t idx2 = code_index(pos2, ha, la, hb, lb, hc, lc, hd, ld);
        atomicAdd(T + idx, 1u);
        if (second) atomicAdd(T + idx2, 1u);
    }
}

// Add the accumulated counts into T and clear the accumulators.  Lane (g, j) holds rows g and g + 8,
// columns 2j and 2j + 1 of each column half: no cross-lane step.
__device__ __forceinline__ void flush_mma(int (&acc)[kScales][2][4], uint32_t *T, int lane)
{
    const int g = lane >> 2, j = lane & 3;
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
            uint32_t c0 = 0u, c1 = 0u;
#pragma unroll
            for (int v = 0; v < kScales; ++v) {
                c0 += (uint32_t)acc[v][h][2 * rr] >> (2 * v);
                c1 += (uint32_t)acc[v][h][2 * rr + 1] >> (2 * v);
                acc[v][h][2 * rr] = 0; acc[v][h][2 * rr + 1] = 0;
            }
            uint2 *t = reinterpret_cast<uint2 *>(T + (g + 8 * rr) * 16 + 8 * h + 2 * j);
            uint2 cur = *t;
            cur.x += c0; cur.y += c1;
            *t = cur;
        }
}

// ---- subsample mode on CODE planes --------------------------------------------------------------------------------
// The packed matrix is one-hot ({A, C, G, T} words: what full mode, the tip-pair GEMM and the POPC kernel want).  For the
// per-locus selection a 2-bit code per base plus a validity bit is cheaper: launch_resolve_mma<true> first converts the
// matrix into {hi, lo, valid, 0} words (A, C, G, T = 0..3: hi = G | T, lo = C | T; one streaming pass, 0.1 ms at C4, into
// stream-ordered scratch), and the rank-chunk loop then needs 17 LOP3 instead of 30: complete = AND of four validity words
// (2), some tip differs from the first = (a^b)|(a^c)|(a^d) on the hi and on the lo words (4), sel (2), taken (1), eight
// merged code planes (8).  The sixteen one-hot planes the tile product wants are expanded once per tile.
__global__ void __launch_bounds__(256) planes_to_codes_kernel(const uint4 *__restrict__ planes, uint4 *__restrict__ codes, size_t n)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const uint4 p = __ldg(planes + i);
        codes[i] = make_uint4(p.z | p.w, p.y | p.w, p.x | p.y | p.z | p.w, 0u);
    }
}
// a.x = hi, a.y = lo, a.z = valid of 32 columns of the first tip, ...
__device__ __forceinline__ void select_chunk_codes(const uint4 &a, const uint4 &b, const uint4 &c, const uint4 &d, uint32_t &taken,
                                                   uint32_t (&H)[4], uint32_t (&L)[4])
{
    const uint32_t valid = a.z & b.z & c.z & d.z;
    const uint32_t dh = (a.x ^ b.x) | (a.x ^ c.x) | (a.x ^ d.x), dl = (a.y ^ b.y) | (a.y ^ c.y) | (a.y ^ d.y);
    const uint32_t sel = valid & ~taken & (dh | dl);
    taken |= sel;
    H[0] |= a.x & sel; H[1] |= b.x & sel; H[2] |= c.x & sel; H[3] |= d.x & sel;
    L[0] |= a.y & sel; L[1] |= b.y & sel; L[2] |= c.y & sel; L[3] |= d.y & sel;
}
__device__ __forceinline__ void expand_tile_planes(uint32_t taken, uint32_t h, uint32_t l, uint32_t (&M)[4])
{
    M[0] = taken & ~h & ~l; M[1] = taken & ~h & l; M[2] = taken & h & ~l; M[3] = taken & h & l;
}

// Per-warp ring: stage 1 by cp.async (L1-bypassing .cg: consecutive quartets' shared rows hitting in L1 was measured 3 %
// slower), stage 2 by count_tile_mma.  The tile's rank count is fetched one tile ahead, the ring reads are issued before
// the next chunk's address arithmetic and cp.async so that their latency is covered, and the global addresses are four
// 64-bit row pointers plus one running 32-bit offset.
template <bool kSubsample>
__device__ __forceinline__ void count_quartet_mma(const PlanDev &plan, const uint4 *__restrict__ planes,
                                                  uint4 tips, const uint4 *ring, uint32_t *P, uint32_t *T, int lane,
                                                  uint32_t flush_every, uint32_t stack_fill, uint32_t stack_ovf, uint32_t *stats)
{
    stack_fill &= 0x7FFFFFFFu;                        // the top bit asked for the diagnostics (`stats` is then non-null)
    static_assert((kMStages & (kMStages - 1)) == 0, "ring depth must be a power of two");
    int acc[kScales][2][4];
#pragma unroll
    for (int v = 0; v < kScales; ++v)
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[v][h][i] = 0;
    const MmaLanePtrs lp = mma_lane_ptrs(P, lane);
    uint32_t pending = 0;
    const uint32_t nchunks = plan.row_u4 / kTileWords;
    const uint4 *g0 = planes + (size_t)tips.x * plan.row_u4 + lane, *g1 = planes + (size_t)tips.y * plan.row_u4 + lane;
    const uint4 *g2 = planes + (size_t)tips.z * plan.row_u4 + lane, *g3 = planes + (size_t)tips.w * plan.row_u4 + lane;
    const uint32_t ring_s = smem_u32(ring + lane);
    constexpr uint32_t kStageBytes = 4 * 32 * 16, kRingMask = (kMStages - 1) * kStageBytes;
    // chunk at uint4 offset `off` of the rows -> ring stage at byte offset `sb`
    auto issue = [&](uint32_t sb, uint32_t off) {
        const uint32_t s = ring_s + sb;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(g0 + off) : "memory");
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s + 512), "l"(g1 + off) : "memory");
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s + 1024), "l"(g2 + off) : "memory");
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s + 1536), "l"(g3 + off) : "memory");
    };
#pragma unroll
    for (int s = 0; s < kMStages - 1; ++s) {
        if ((uint32_t)s < nchunks) issue(s * kStageBytes, s * kTileWords);
        cp_async_commit();
    }
    // The loop's bookkeeping is two running values: the ring stage of the chunk being read (byte offset `rd`; the stage to
    // refill is the one before it) and the row offset of the chunk to fetch (`gnext`, uint4 units).
    uint32_t rd = 0u, gnext = (kMStages - 1) * kTileWords;
    const uint32_t gend = nchunks * kTileWords;
    const int ntiles = plan.ntiles;
    uint32_t Knext = ntiles > 0 ? __ldg(plan.tile_k) : 0u;
    uint32_t occ = 0u;                                // bit slots in use (subsample mode: stacked tiles)
    uint32_t SH[4] = {0u, 0u, 0u, 0u}, SL[4] = {0u, 0u, 0u, 0u};   // the stack: merged code planes of the tiles not multiplied out yet
    bool stacked = false;                             // the stack holds tiles that no product has taken yet (warp-uniform)
    uint32_t n_products = 0u, n_stacked = 0u, n_walk = 0u;   // diagnostics: status[1..3]
    // Subsample mode runs one extra, empty pass of the tile loop so that the product of the last stacked tiles is taken at
    // the same (single) call site as every other product: the IMMA body is in the instruction stream once.
    const int npasses = kSubsample ? ntiles + 1 : ntiles;
    for (int tile = 0; tile < npasses; ++tile) {
        const bool last = kSubsample && tile == ntiles;
        const uint32_t K = last ? 0u : Knext;
        if (tile + 1 < ntiles) Knext = __ldg(plan.tile_k + tile + 1);
        uint32_t taken = 0u;
        uint32_t H[4] = {0u, 0u, 0u, 0u}, L[4] = {0u, 0u, 0u, 0u};      // subsample mode: merged code planes of the four tips
        for (uint32_t k = K; k != 0u; --k) {
            cp_async_wait<kMStages - 2>();            // this chunk has landed (the next kMStages-2 may be in flight)
            const uint4 *st = reinterpret_cast<const uint4 *>(reinterpret_cast<const char *>(ring + lane) + rd);
            const uint4 a = st[0], b = st[32], c = st[64], d = st[96];
            // refill the stage read in the previous iteration; keeps kMStages-1 chunks in flight
            if (gnext < gend) issue((rd + kRingMask) & kRingMask, gnext);
            cp_async_commit();
            gnext += kTileWords;
            rd = (rd + kStageBytes) & kRingMask;
            if (kSubsample) {
                select_chunk_codes(a, b, c, d, taken, H, L);
            } else {
                // full mode (resolve_quartets.py:92-95): every complete column counts; the four
                // invariant patterns aaaa are zeroed by the caller
                const uint32_t A[4] = {a.x, a.y, a.z, a.w}, B[4] = {b.x, b.y, b.z, b.w};
                const uint32_t C[4] = {c.x, c.y, c.z, c.w}, D[4] = {d.x, d.y, d.z, d.w};
                count_tile_mma(A, B, C, D, lp, acc);
                if (++pending == flush_every) { flush_mma(acc, T, lane); pending = 0; }
            }
        }
        if (kSubsample) {
            // The stack of tiles lives in REGISTERS as eight code planes (SH, SL) + `occ`; stacking a tile is nine LOP3.  The
            // price of stacking this tile is known exactly: the lane with the most taken slots walks its overflow bit by
            // bit.  Past `stack_ovf` loci, or once `stack_fill` slots are in use, the stack is multiplied out first -- only
            // then are its sixteen one-hot planes expanded and stored to the tile copy P -- and the tile starts a new one.
            const uint32_t ovf = taken & occ;
            const uint32_t worst = __reduce_max_sync(kFull, (uint32_t)__popc(ovf));
            // (the fill rule is off by default -- 100 % -- and then costs nothing: the overflow rule alone decides)
            const uint32_t filled = stack_fill < (uint32_t)kTileLoci ? __reduce_add_sync(kFull, (uint32_t)__popc(occ)) : 0u;
            if (stacked && (last || worst > stack_ovf || filled >= stack_fill)) {
                uint32_t MA[4], MB[4], MC[4], MD[4];
                expand_tile_planes(occ, SH[0], SL[0], MA); expand_tile_planes(occ, SH[1], SL[1], MB);
                expand_tile_planes(occ, SH[2], SL[2], MC); expand_tile_planes(occ, SH[3], SL[3], MD);
                store_tile_planes(MA, MB, MC, MD, lp);
                tile_product_mma(lp, acc);
                ++n_products;
                stacked = false;
                if (++pending == flush_every) { __syncwarp(); flush_mma(acc, T, lane); pending = 0; }
            }
            if (last) break;
            if (!stacked) {
#pragma unroll
                for (int i = 0; i < 4; ++i) { SH[i] = H[i]; SL[i] = L[i]; }
                occ = taken;
                stacked = true;
            } else {
                const uint32_t put = taken & ~occ;
#pragma unroll
                for (int i = 0; i < 4; ++i) { SH[i] |= H[i] & put; SL[i] |= L[i] & put; }
                occ |= taken;
                ++n_stacked;
                n_walk += worst;
                if (worst) scatter_overflow(ovf, H, L, T);
            }
        }
    }
    cp_async_wait<0>();
    if (kSubsample && stats && lane == 0) {
        atomicAdd(stats + 1, n_products); atomicAdd(stats + 2, n_stacked); atomicAdd(stats + 3, n_walk);
    }
    __syncwarp();                                     // the overflow atomics on T precede the accumulators' flush
    flush_mma(acc, T, lane);
}

struct __align__(16) WsSlotMeta { long long q; uint4 tips; };   // hand-off slot: quartet index (or -1 - q: bad tips) and tips

// ------------------------------------------------------------------------------------------
// Warp-specialised kernel around the tensor-core count path.  A CTA has NC count warps and NS score warps.  The count
// path has no XU bottleneck and plenty of independent work per warp, so few count warps keep the SM's integer pipes busy;
// the FP64 Jacobi is bound by its own dependency chains, so half of the register file goes to score warps.  Hand-off
// through a CTA-wide ring of NC + 4 NS slots of 256 counts: producers take one ticket, consumers four (one shared-memory
// atomic each), slot = ticket % slots, one full / empty mbarrier pair per slot, so ANY score warp takes the next finished
// quartets -- no fixed count -> score pairing (every wait traps after seconds instead of hanging).
// ------------------------------------------------------------------------------------------
template <int NC, int NS>
struct __align__(16) MmaShared {
    static constexpr int kSlots = NC + 4 * NS;      // every count warp fills one slot, every score warp holds up to four
    uint4 ring[NC][kMStages][4][32];                // per-warp cp.async rings
    uint32_t P[NC][16 * kPStride];
    uint32_t T[kSlots][256];
    WsSlotMeta meta[kSlots];
    SvdScratch4 svd[NS];
    unsigned long long full[kSlots], empty[kSlots];
    uint32_t produce, consume;
};

template <bool kSubsample, int NC, int NS>
__global__ void __launch_bounds__((NC + NS) * 32, 1)
resolve_mma_kernel(PlanDev plan, const uint4 *__restrict__ planes, uint32_t ntips,
                   const uint4 *__restrict__ quartets, int64_t m, uint64_t seed,
                   double *__restrict__ rscor, uint2 *__restrict__ rstat,
                   uint32_t *__restrict__ counts, uint32_t *__restrict__ minrank_out,
                   uint32_t *__restrict__ status, uint32_t flush_every, uint32_t stack_fill, uint32_t stack_ovf,
                   const uint32_t *__restrict__ skip)
{
    if (skip && skip[0]) return;                     // full mode: the tip-pair GEMM took this chunk (tq_pairgemm.cu)
    using Shared = MmaShared<NC, NS>;
    constexpr int kSlots = Shared::kSlots;
    static_assert(kSlots <= 256, "slot indices travel in one byte each");
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Shared &sh = *reinterpret_cast<Shared *>(smem_raw);
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;

    if (threadIdx.x == 0) {
        for (int s = 0; s < kSlots; ++s) { mbar_init(&sh.full[s], 1); mbar_init(&sh.empty[s], 1); }
        sh.produce = 0u; sh.consume = 0u;
        mbar_init_fence();
    }
    __syncthreads();

    // Work split: batch b = blockIdx.x + it * gridDim.x holds quartets b * NC .. b * NC + NC - 1, one per count warp.
    const int64_t nbatches = (m + NC - 1) / NC;
    const int64_t my_batches = (int64_t)blockIdx.x < nbatches ? (nbatches - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    // quartets this CTA resolves (the last batch may be short)
    int64_t total = my_batches * NC;
    {
        const int64_t last_b = (int64_t)blockIdx.x + (my_batches - 1) * gridDim.x;
        if (my_batches > 0 && last_b == nbatches - 1) total -= nbatches * NC - m;
    }

    if (warp < NC) {
        // ------------------------------ count warps ------------------------------
        const int c = warp;
        for (int64_t it = 0; it < my_batches; ++it) {
            const int64_t q = ((int64_t)blockIdx.x + it * gridDim.x) * NC + c;
            if (q >= m) break;                                               // the last batch may be short
            uint32_t ticket = 0;
            if (lane == 0) ticket = atomicAdd(&sh.produce, 1u);
            ticket = __shfl_sync(kFull, ticket, 0);
            const uint32_t slot = ticket % kSlots;
            mbar_wait(&sh.empty[slot], ((ticket / kSlots) & 1u) ^ 1u);       // slot free (passes at once on its first use)
            uint32_t *T = sh.T[slot];
            const uint4 t = __ldg(quartets + q);
            const bool bad = t.x >= ntips || t.y >= ntips || t.z >= ntips || t.w >= ntips;
#pragma unroll
            for (int i = 0; i < 8; ++i) T[lane + 32 * i] = 0u;
            __syncwarp();
            if (!bad) count_quartet_mma<kSubsample>(plan, planes, t, &sh.ring[c][0][0][0], sh.P[c], T, lane, flush_every, stack_fill, stack_ovf, (stack_fill & 0x80000000u) ? status : nullptr);
            if (lane == 0) { sh.meta[slot].q = bad ? -1 - q : q; sh.meta[slot].tips = t; }
            __syncwarp();
            if (lane == 0) mbar_arrive(&sh.full[slot]);
        }
    } else {
        // ------------------------------ score warps ------------------------------
        // A score warp takes up to four finished quartets at a time (svd_scores4: 12 matrices in three passes of four).
        SvdScratch4 &sv = sh.svd[warp - NC];
        while (true) {
            uint32_t ticket0 = 0;
            if (lane == 0) ticket0 = atomicAdd(&sh.consume, 4u);
            ticket0 = __shfl_sync(kFull, ticket0, 0);
            if ((int64_t)ticket0 >= total) break;
            const int nq = (int)((total - (int64_t)ticket0) < 4 ? (total - (int64_t)ticket0) : 4);
            uint32_t slots = 0u, actmask = 0u, my_nsnps = 0u, my_best = 0u;
            long long my_q = -1;                                           // lane j < 4: quartet j of this batch
#pragma unroll 1
            for (int j = 0; j < nq; ++j) {
                const uint32_t ticket = ticket0 + (uint32_t)j;
                const uint32_t slot = ticket % kSlots;
                mbar_wait(&sh.full[slot], (ticket / kSlots) & 1u);
                slots |= slot << (8 * j);
                uint32_t *T = sh.T[slot];
                const long long mq = sh.meta[slot].q;
                const uint4 t = sh.meta[slot].tips;
                if (mq < 0) {                                              // tip index out of range
                    const long long q = -1 - mq;
                    if (lane == 0) {
                        rstat[q] = make_uint2(0xFFFFFFFFu, 0u);
                        rscor[3 * q] = rscor[3 * q + 1] = rscor[3 * q + 2] = 0.0;
                        if (status) atomicOr(status, 1u);
                    }
                    continue;
                }
                const uint32_t nsnps = finish_counts<kSubsample>(T, lane, counts, mq);
                uint32_t best = 0;
                if (nsnps == 0) best = tiebreak_split(seed, t);            // stands in for :231
                else actmask |= 1u << j;
                if (lane == j) { my_q = mq; my_nsnps = nsnps; my_best = best; }
            }
            double score[3] = {0.001, 0.001, 0.001};                        // :232
            uint32_t minrank = 0;
            if (actmask) svd_scores4(&sh.T[0][0], slots, actmask, sv, lane, score, my_best, minrank);
            if (lane < 4 && my_q >= 0) {
                rscor[3 * my_q] = score[0]; rscor[3 * my_q + 1] = score[1]; rscor[3 * my_q + 2] = score[2];
                rstat[my_q] = make_uint2(my_best, my_nsnps);                // :251, :264
                if (minrank_out) minrank_out[my_q] = minrank;
            }
            __syncwarp();
            if (lane == 0)
                for (int j = 0; j < nq; ++j) mbar_arrive(&sh.empty[(ticket0 + (uint32_t)j) % kSlots]);
        }
    }
}

// Count warps for a plan.  Per quartet a count warp spends ~1 980 clocks per tile product (stage 2) and ~440 per rank
// chunk (stage 1), a score warp ~190 000 clocks (phase clocks at C2, profiles/r01_v7_resolve_ncu.md); the count share of
// the 16 warps is rounded to the nearest instantiated mix.  Full mode runs one tile product per rank chunk.
static int pick_count_warps(const PlanDev &plan, bool subsample)
{
    const int forced = g_count_warps.load();
    if (forced > 0) return forced >= 10 ? 12 : 8;
    const double nchunks = (double)(plan.row_u4 / kTileWords);
    const double products = subsample ? (double)plan.ntiles : nchunks;
    const double count = 1980.0 * products + 440.0 * nchunks, score = 190000.0;
    return count >= 1.5 * score ? 12 : 8;
}

template <bool kSubsample, int NC>
static cudaError_t launch_resolve_mma_mix(const PlanDev &plan, const uint4 *planes, uint32_t ntips, const uint4 *quartets,
                                          int64_t m, uint64_t seed, double *d_rscor, uint2 *rstat, uint32_t *d_counts,
                                          uint32_t *d_minrank, uint32_t *d_status, int num_sms, cudaStream_t stream,
                                          const uint32_t *d_skip)
{
    constexpr int NS = kMmaWarps - NC;
    auto kern = resolve_mma_kernel<kSubsample, NC, NS>;
    constexpr size_t smem = sizeof(MmaShared<NC, NS>);
    static_assert(smem <= 227 * 1024, "warp mix does not fit in shared memory");
    static SmemAttr attr;
    cudaError_t e = attr.ensure(kern, smem);
    if (e != cudaSuccess) return e;
    const int64_t want = (m + NC - 1) / NC;
    const unsigned grid = (unsigned)(want < num_sms ? want : num_sms);
    const int64_t fv = g_mma_flush.load();
    const uint32_t flush_every = (fv >= 1 && fv < (int64_t)kMmaFlushEvery) ? (uint32_t)fv : kMmaFlushEvery;
    uint4 *codes = nullptr;
    if (kSubsample) {                                // one-hot planes -> {hi, lo, valid, 0} words (see select_chunk_codes)
        const size_t n = (size_t)ntips * plan.row_u4;
        if (n) {
            e = cudaMallocAsync(&codes, n * sizeof(uint4), stream);
            if (e != cudaSuccess) return e;
            const size_t blocks = (n + 255) / 256, cap = (size_t)num_sms * 16;
            planes_to_codes_kernel<<<(unsigned)(blocks < cap ? blocks : cap), 256, 0, stream>>>(planes, codes, n);
            ++g_launches;
            planes = codes;
        }
    }
    kern<<<grid, kMmaWarps * 32, smem, stream>>>(plan, planes, ntips, quartets, m, seed, d_rscor, rstat, d_counts, d_minrank,
                                                 d_status, flush_every, stack_fill_slots(), (uint32_t)g_stack_ovf.load(), d_skip);
    ++g_launches;
    g_last_kernel = kSubsample ? "tq::resolve_mma_kernel<true>" : "tq::resolve_mma_kernel<false>";
    g_last_count_warps = NC;
    e = cudaGetLastError();
    if (codes) cudaFreeAsync(codes, stream);
    return e;
}

template <bool kSubsample>
static cudaError_t launch_resolve_mma(const PlanDev &plan, const uint4 *planes, uint32_t ntips, const uint4 *quartets,
                                      int64_t m, uint64_t seed, double *d_rscor, uint2 *rstat, uint32_t *d_counts,
                                      uint32_t *d_minrank, uint32_t *d_status, int num_sms, cudaStream_t stream,
                                      const uint32_t *d_skip = nullptr)
{
#define TQ_MIX(NC)                                                                                                          \
    case NC:                                                                                                                \
        return launch_resolve_mma_mix<kSubsample, NC>(plan, planes, ntips, quartets, m, seed, d_rscor, rstat, d_counts,    \
                                                      d_minrank, d_status, num_sms, stream, d_skip)
    switch (pick_count_warps(plan, kSubsample)) {
        TQ_MIX(12);
    default:
        TQ_MIX(8);
    }
#undef TQ_MIX
}

cudaError_t launch_resolve_mma_full(const PlanDev &plan, const uint32_t *d_planes, int64_t ntips, const uint32_t *d_quartets,
                                    int64_t m, uint64_t seed, double *d_rscor, uint32_t *d_rstat, uint32_t *d_counts,
                                    uint32_t *d_minrank, uint32_t *d_status, int num_sms, const uint32_t *d_skip,
                                    cudaStream_t stream)
{
    return launch_resolve_mma<false>(plan, reinterpret_cast<const uint4 *>(d_planes), (uint32_t)ntips,
                                     reinterpret_cast<const uint4 *>(d_quartets), m, seed, d_rscor, reinterpret_cast<uint2 *>(d_rstat),
                                     d_counts, d_minrank, d_status, num_sms, stream, d_skip);
}

cudaError_t launch_resolve(const PlanDev &plan, const uint32_t *d_planes, int64_t ntips,
                           const uint32_t *d_quartets, int64_t m, bool subsample, uint64_t seed,
                           double *d_rscor, uint32_t *d_rstat, uint32_t *d_counts,
                           uint32_t *d_minrank, uint32_t *d_status, int num_sms,
                           cudaStream_t stream)
{
    if (m == 0) return cudaSuccess;
    const uint4 *planes = reinterpret_cast<const uint4 *>(d_planes);
    const uint4 *quartets = reinterpret_cast<const uint4 *>(d_quartets);
    uint2 *rstat = reinterpret_cast<uint2 *>(d_rstat);
    // Variants of the same device code.  On tiny launches (C1: 495 quartets) a count -> score hand-off is pure latency, so
    // the one-warp-per-quartet kernel runs; from 16 quartets per SM the tensor-core count kernel (profiles/tools/
    // time_small.py, C2 matrix, M quartets/s fused / mma: 1 184 quartets 4.5 / 4.1, 2 368: 6.0 / 6.2, 4 736: 7.1 / 8.0,
    // 37 888: 7.4 / 10.7).  Full mode with a dense quartet list goes to the tip-pair GEMM (tq_pairgemm.cu).
    const int force = g_kernel_choice.load();
    if (!subsample && (force == 5 || (force == 0 && pairgemm_applies(ntips, m, num_sms))))
        return launch_resolve_pairgemm(plan, d_planes, ntips, d_quartets, m, seed, d_rscor, d_rstat, d_counts, d_minrank,
                                       d_status, num_sms, force, stream);
    const bool use_mma = force ? (force == 3 || force == 5) : (m >= (int64_t)num_sms * 16);
    if (use_mma)
        return subsample ? launch_resolve_mma<true>(plan, planes, (uint32_t)ntips, quartets, m, seed, d_rscor, rstat, d_counts,
                                                    d_minrank, d_status, num_sms, stream)
                         : launch_resolve_mma<false>(plan, planes, (uint32_t)ntips, quartets, m, seed, d_rscor, rstat, d_counts,
                                                     d_minrank, d_status, num_sms, stream);
    g_last_kernel = subsample ? "tq::resolve_kernel<true>" : "tq::resolve_kernel<false>";
    const size_t smem = sizeof(WarpScratch) * kWarps;
    static SmemAttr attr_sub, attr_full;
    cudaError_t e = subsample ? attr_sub.ensure(resolve_kernel<true>, smem) : attr_full.ensure(resolve_kernel<false>, smem);
    if (e != cudaSuccess) return e;
    const int64_t want = (m + kWarps - 1) / kWarps;
    const int64_t cap = (int64_t)num_sms * 4;
    const unsigned grid = (unsigned)(want < cap ? want : cap);
    if (subsample)
        resolve_kernel<true><<<grid, kWarps * 32, smem, stream>>>(plan, planes, (uint32_t)ntips, quartets, m, seed, d_rscor, rstat,
                                                                  d_counts, d_minrank, d_status);
    else
        resolve_kernel<false><<<grid, kWarps * 32, smem, stream>>>(plan, planes, (uint32_t)ntips, quartets, m, seed, d_rscor, rstat,
                                                                   d_counts, d_minrank, d_status);
    ++g_launches;
    return cudaGetLastError();
}

}  // namespace tq
