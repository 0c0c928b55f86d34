// Internal declarations shared by tq_kernels.cu (device code + launchers) and tq_api.cu (C ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

namespace tq {

constexpr int kTileLoci  = 1024;  // loci per tile = 32 lanes x 32 bits
constexpr int kTileWords = 32;    // uint4 {A,C,G,T} words per rank of a tile (one per lane)

// Device-side view of a locus plan (see include/tetrad_b200.h, tq_plan_create).
struct PlanDev {
    const uint32_t *tile_off;   // [ntiles+1] offset of tile in a tip row, in uint4 units
    const uint32_t *tile_k;     // [ntiles]   ranks (max locus length) of the tile
    const uint32_t *loc_start;  // [ntiles*1024] first column of sorted locus l (pad: 0)
    const uint32_t *loc_len;    // [ntiles*1024] columns of sorted locus l      (pad: 0)
    const uint32_t *loc_run;    // [ntiles*1024] index of the locus run before the length sort
    const uint32_t *loc_pos;    // [ntiles*1024+1] columns of the sorted loci before l (prefix sum of loc_len)
    int      ntiles;
    uint32_t row_u4;            // uint4 per tip row
    uint32_t ncols;             // columns of the replicate (sum of loc_len)
    uint32_t nloci;             // locus runs (the tables are padded to ntiles*1024)
};

// tip rows [tip_lo, tip_hi) of the matrix at d_tmparr / d_seqarr -> the same rows of d_planes (both full-matrix bases)
cudaError_t launch_pack(const PlanDev &plan, const uint8_t *d_tmparr, int64_t tip_lo, int64_t tip_hi,
                        int64_t row_stride, uint32_t *d_planes, cudaStream_t stream);

cudaError_t launch_boot_pack(const PlanDev &plan, const uint8_t *d_seqarr, int64_t tip_lo, int64_t tip_hi, int64_t row_stride,
                             const uint32_t *d_src_start, uint32_t seed_shuffle, uint32_t seed_ambig,
                             uint32_t *d_srccol, uint32_t *d_planes, cudaStream_t stream);

cudaError_t launch_resolve(const PlanDev &plan, const uint32_t *d_planes, int64_t ntips,
                           const uint32_t *d_quartets, int64_t m, bool subsample, uint64_t seed,
                           double *d_rscor, uint32_t *d_rstat, uint32_t *d_counts,
                           uint32_t *d_minrank, uint32_t *d_status, int num_sms,
                           cudaStream_t stream);

// full mode over tip pairs on tcgen05 (tq_pairgemm.cu): whether a launch is worth planning; the planner then routes every
// chunk of the list, on the device, to the GEMM (dense) or to the per-quartet kernel (launch_resolve_mma_full with the
// route flag as `d_skip`); force = 5 always takes the GEMM
bool pairgemm_applies(int64_t ntips, int64_t m, int num_sms);
cudaError_t launch_resolve_pairgemm(const PlanDev &plan, const uint32_t *d_planes, int64_t ntips, const uint32_t *d_quartets,
                                    int64_t m, uint64_t seed, double *d_rscor, uint32_t *d_rstat, uint32_t *d_counts,
                                    uint32_t *d_minrank, uint32_t *d_status, int num_sms, int force, cudaStream_t stream);
cudaError_t launch_resolve_mma_full(const PlanDev &plan, const uint32_t *d_planes, int64_t ntips, const uint32_t *d_quartets,
                                    int64_t m, uint64_t seed, double *d_rscor, uint32_t *d_rstat, uint32_t *d_counts,
                                    uint32_t *d_minrank, uint32_t *d_status, int num_sms, const uint32_t *d_skip,
                                    cudaStream_t stream);
void pairgemm_stats(int64_t out[2]);     // chunks routed to the GEMM / to the per-quartet kernel since load

cudaError_t launch_unrank(const int64_t *d_ranks, int64_t start, int64_t m, uint32_t ntips, uint32_t *d_quartets,
                          cudaStream_t stream);

cudaError_t launch_sample_unrank(int64_t start, int64_t m, uint32_t ntips, uint64_t key, uint32_t *d_quartets,
                                 cudaStream_t stream);

void sample_ranks_host(int64_t start, int64_t m, uint32_t ntips, uint64_t key, int64_t *out);

cudaError_t run_microbench(int kind, double *result);

extern std::atomic<int64_t> g_launches;      // kernels launched by this library since load
extern std::atomic<int> g_kernel_choice;          // 0 by launch size and mode, 1 fused POPC, 3 tensor-core count, 5 tip-pair GEMM (full mode)
extern std::atomic<int64_t> g_mma_flush;          // > 0: tile products between accumulator flushes of the tensor-core kernel (tests)
extern std::atomic<const char *> g_last_kernel;   // name of the resolve kernel variant launched last
extern std::atomic<int> g_count_warps;            // 0 by the plan; else count warps per CTA of the tensor-core kernel (of 16)
extern std::atomic<int> g_stack_fill;             // < 0 default; else percent of a tile's bit slots filled before a stacked product (0: never stack)
extern std::atomic<int> g_stack_ovf;              // most overflow loci a lane may have for a tile to be stacked on the previous ones
extern std::atomic<int> g_stack_stats;            // != 0: stacking diagnostics are added to status[1..3]
extern std::atomic<int> g_last_count_warps;       // count warps per CTA of the last tensor-core launch

}  // namespace tq
