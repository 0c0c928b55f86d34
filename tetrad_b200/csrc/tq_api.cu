// C ABI of libtetrad_b200.so (declared in include/tetrad_b200.h): locus plan, pack, resolve,
// host-buffer convenience entry, micro-benchmarks.  Host code only; kernels are in tq_kernels.cu.
#include "../../include/tetrad_b200.h"
#include "tq_kernels.cuh"
#include "tq_error.h"

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <new>
#include <unordered_set>
#include <vector>

namespace {

thread_local char g_err[512] = "";

#define fail tq::set_error

#define TQ_CUDA(expr)                                                                      \
    do {                                                                                   \
        cudaError_t e__ = (expr);                                                          \
        if (e__ != cudaSuccess)                                                            \
            return fail(TQ_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), \
                        __FILE__, __LINE__);                                               \
    } while (0)

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) { cudaGetDevice(&prev); if (dev != prev) cudaSetDevice(dev); }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

}  // namespace

int tq::set_error(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

// Pinned host buffers for the plan tables.  A replicate loop creates one plan per replicate while the previous
// replicate's kernels are still running: the upload must neither block the host (pageable memory would, once the
// driver's staging buffers are full behind a 40 ms kernel) nor pay cudaHostAlloc every time.  Buffers go back to a
// free list together with the event that marks the end of their last upload.
namespace {
struct PinnedBuf { uint32_t *ptr = nullptr; size_t words = 0; cudaEvent_t done = nullptr; };
std::mutex g_pinned_mu;
std::vector<PinnedBuf> g_pinned_free;

PinnedBuf pinned_get(size_t words)
{
    {
        std::lock_guard<std::mutex> lk(g_pinned_mu);
        for (size_t i = 0; i < g_pinned_free.size(); ++i)
            if (g_pinned_free[i].words >= words && g_pinned_free[i].words <= 2 * words + 4096) {
                PinnedBuf b = g_pinned_free[i];
                g_pinned_free.erase(g_pinned_free.begin() + (long)i);
                if (b.done) cudaEventSynchronize(b.done);          // its last upload (long over in practice)
                return b;
            }
    }
    PinnedBuf b;
    const size_t want = words + words / 8 + 1024;
    if (cudaHostAlloc(reinterpret_cast<void **>(&b.ptr), want * sizeof(uint32_t), cudaHostAllocDefault) != cudaSuccess) {
        cudaGetLastError();
        return PinnedBuf{};
    }
    b.words = want;
    cudaEventCreateWithFlags(&b.done, cudaEventDisableTiming);
    return b;
}

void pinned_put(PinnedBuf b)
{
    if (!b.ptr) return;
    std::lock_guard<std::mutex> lk(g_pinned_mu);
    if (g_pinned_free.size() >= 8) {                       // keep a handful; the rest goes back to the driver
        if (b.done) { cudaEventSynchronize(b.done); cudaEventDestroy(b.done); }
        cudaFreeHost(b.ptr);
        return;
    }
    g_pinned_free.push_back(b);
}
}  // namespace

struct tq_plan {
    int device = 0;
    int num_sms = 0;
    int64_t ncols = 0, nloci = 0, ntiles = 0, row_u4 = 0, max_len = 0, padded_cols = 0;
    void *d_tables = nullptr;      // one allocation: tile_off | tile_k | loc_start | loc_len
    cudaStream_t stream = nullptr; // the tables are stream-ordered memory of the stream the plan was created on
    std::vector<uint32_t> h_vec;    // host copy of the same tables: a vector for host-only plans,
    PinnedBuf h_pin;                // a pinned buffer (the upload's source) for device plans
    const uint32_t *h_tables() const { return h_pin.ptr ? h_pin.ptr : h_vec.data(); }
    size_t o_off = 0, o_k = 0, o_start = 0, o_len = 0, o_run = 0, o_pos = 0;
    tq::PlanDev dev{};
};

static constexpr int kHostChunks = 4;     // tq_resolve_host: result chunks copied back beside the next chunk's kernel

struct tq_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;    // device -> host copies of finished result chunks
    cudaEvent_t chunk_done[kHostChunks] = {};
    // cached replicate state
    tq_plan *plan = nullptr;
    const void *key_arr = nullptr, *key_loc = nullptr;
    int64_t key_ntips = -1, key_ncols = -1;
    uint64_t key_epoch = 0;
    // growable device buffers
    void *d_tmparr = nullptr;  size_t cap_tmparr = 0;
    void *d_planes = nullptr;  size_t cap_planes = 0;
    void *d_quart = nullptr;   size_t cap_quart = 0;
    void *d_out = nullptr;     size_t cap_out = 0;
    int64_t last_h2d = 0, last_d2h = 0;
};

extern "C" {

int tq_version(void) { return TQ_VERSION; }

const char *tq_last_error(void) { return g_err; }

int64_t tq_launch_count(void) { return tq::g_launches.load(); }
const char *tq_last_kernel(void) { return tq::g_last_kernel.load(); }
int tq_last_count_warps(void) { return tq::g_last_count_warps.load(); }

int tq_pairgemm_stats(int64_t out[2])
{
    if (!out) return fail(TQ_EINVAL, "tq_pairgemm_stats: out is NULL");
    tq::pairgemm_stats(out);
    return TQ_OK;
}

int tq_set_option(const char *name, int64_t value)
{
    if (!name) return fail(TQ_EINVAL, "tq_set_option: null name");
    if (!strcmp(name, "kernel")) {
        if (value != 0 && value != 1 && value != 3 && value != 5)
            return fail(TQ_EINVAL, "tq_set_option: kernel must be 0 (auto), 1 (fused), 3 (mma) or 5 (pairgemm)");
        tq::g_kernel_choice = (int)value;
        return TQ_OK;
    }
    if (!strcmp(name, "mma_flush_every")) {
        if (value < 0) return fail(TQ_EINVAL, "tq_set_option: mma_flush_every must be >= 0");
        tq::g_mma_flush = value;
        return TQ_OK;
    }
    if (!strcmp(name, "stack_fill")) {
        if (value < -1 || value > 100) return fail(TQ_EINVAL, "tq_set_option: stack_fill must be -1 (default) or 0..100 (percent)");
        tq::g_stack_fill = (int)value;
        return TQ_OK;
    }
    if (!strcmp(name, "stack_stats")) {
        tq::g_stack_stats = value != 0;
        return TQ_OK;
    }
    if (!strcmp(name, "stack_ovf")) {
        if (value < 0 || value > 32) return fail(TQ_EINVAL, "tq_set_option: stack_ovf must be 0..32");
        tq::g_stack_ovf = (int)value;
        return TQ_OK;
    }
    if (!strcmp(name, "count_warps")) {
        if (value != 0 && value != 8 && value != 12)
            return fail(TQ_EINVAL, "tq_set_option: count_warps must be 0 (by the plan), 8 or 12");
        tq::g_count_warps = (int)value;
        return TQ_OK;
    }
    return fail(TQ_EINVAL, "tq_set_option: unknown option");
}

// ------------------------------------------------------------------------------------------
// plan
// ------------------------------------------------------------------------------------------
// Let the device's default memory pool keep what is freed into it (its default gives memory back at every
// synchronisation, which would turn cudaMallocAsync into cudaMalloc again).  Once per device and process.
static cudaError_t keep_pool_cached(int device)
{
    static std::atomic<uint64_t> done{0};
    const uint64_t bit = 1ull << (device & 63);
    if (done.fetch_or(bit) & bit) return cudaSuccess;
    cudaMemPool_t pool = nullptr;
    cudaError_t e = cudaDeviceGetDefaultMemPool(&pool, device);
    if (e != cudaSuccess) return e;
    uint64_t keep = 256ull << 20;          // plan tables are a few MB; cap what the pool may hold on to
    return cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
}

// Build a plan from locus runs given as (first column, length) in column order.
static int build_plan(const std::vector<uint32_t> &run_start, const std::vector<uint32_t> &run_len, int64_t ncols,
                      int device, void *stream, tq_plan **out)
{
    const int64_t nloci = (int64_t)run_start.size();
    // stable counting sort of runs by length, longest first
    uint32_t max_len = 0;
    for (uint32_t l : run_len) max_len = std::max(max_len, l);
    std::vector<int64_t> bucket((size_t)max_len + 2, 0);
    for (uint32_t l : run_len) bucket[max_len - l + 1]++;
    for (size_t i = 1; i < bucket.size(); ++i) bucket[i] += bucket[i - 1];
    const int64_t ntiles = (nloci + tq::kTileLoci - 1) / tq::kTileLoci;
    const int64_t lpad = ntiles * tq::kTileLoci;
    std::vector<uint32_t> vec;
    PinnedBuf pin;
    // layout: tile_off[ntiles+1] | tile_k[ntiles] | loc_start[lpad] | loc_len[lpad] | loc_run[lpad] | loc_pos[lpad+1]
    const size_t o_off = 0, o_k = (size_t)ntiles + 1, o_start = o_k + (size_t)ntiles,
                 o_len = o_start + (size_t)lpad, o_run = o_len + (size_t)lpad, o_pos = o_run + (size_t)lpad,
                 total = o_pos + (size_t)lpad + 1;
    if (device >= 0) {
        DeviceGuard guard(device);
        pin = pinned_get(total + 4);
        if (!pin.ptr) return fail(TQ_ENOMEM, "tq_plan_create: no pinned host memory for the plan tables");
    } else {
        vec.resize(total + 4);
    }
    uint32_t *tables = pin.ptr ? pin.ptr : vec.data();
    // every word below is written by the loops that follow except the padding of the three per-locus tables and the tail
    for (int64_t l = nloci; l < lpad; ++l) tables[o_start + l] = tables[o_len + l] = tables[o_run + l] = 0u;
    for (size_t i = total; i < total + 4; ++i) tables[i] = 0u;
    for (int64_t r = 0; r < nloci; ++r) {
        const int64_t pos = bucket[max_len - run_len[r]]++;
        tables[o_start + pos] = run_start[r];
        tables[o_len + pos] = run_len[r];
        tables[o_run + pos] = (uint32_t)r;
    }
    int64_t off = 0, padded = 0;
    for (int64_t t = 0; t < ntiles; ++t) {
        const uint32_t k = tables[o_len + t * tq::kTileLoci];   // longest locus of the tile
        tables[o_off + t] = (uint32_t)off;
        tables[o_k + t] = k;
        off += (int64_t)k * tq::kTileWords;
        padded += (int64_t)k * tq::kTileLoci;
        if (off >= ((int64_t)1 << 31)) { pinned_put(pin); return fail(TQ_EINVAL, "tq_plan_create: packed row too large"); }
    }
    tables[o_off + ntiles] = (uint32_t)off;
    {   // columns of the sorted loci before l: where locus l's columns sit once all columns are grouped by sorted locus
        uint32_t pos = 0;
        for (int64_t l = 0; l < lpad; ++l) { tables[o_pos + l] = pos; pos += tables[o_len + l]; }
        tables[o_pos + lpad] = pos;
    }

    tq_plan *p = new (std::nothrow) tq_plan();
    if (!p) { pinned_put(pin); return fail(TQ_ENOMEM, "tq_plan_create: out of host memory"); }
    p->device = device;
    p->ncols = ncols; p->nloci = nloci; p->ntiles = ntiles; p->row_u4 = off;
    p->max_len = max_len; p->padded_cols = padded;
    p->o_off = o_off; p->o_k = o_k; p->o_start = o_start; p->o_len = o_len; p->o_run = o_run; p->o_pos = o_pos;
    p->dev.ntiles = (int)ntiles;
    p->dev.row_u4 = (uint32_t)off;
    p->dev.ncols = (uint32_t)ncols;
    p->dev.nloci = (uint32_t)nloci;
    if (device < 0) {               // host-only plan: layout queries, no kernels
        p->h_vec.swap(vec);
        *out = p;
        return TQ_OK;
    }
    p->h_pin = pin;

    DeviceGuard guard(device);
    cudaError_t e = cudaDeviceGetAttribute(&p->num_sms, cudaDevAttrMultiProcessorCount, device);
    // Stream-ordered allocation from the device's default pool: a replicate loop creates and destroys one plan per
    // replicate, and cudaMalloc / cudaFree (which synchronises the whole device) were the jitter of the host entry.
    if (e == cudaSuccess) e = keep_pool_cached(device);
    p->stream = (cudaStream_t)stream;
    if (e == cudaSuccess) e = cudaMallocAsync(&p->d_tables, (total + 4) * sizeof(uint32_t), p->stream);
    if (e == cudaSuccess)       // pinned source owned by the plan: asynchronous, the host goes on
        e = cudaMemcpyAsync(p->d_tables, tables, (total + 4) * sizeof(uint32_t), cudaMemcpyHostToDevice, (cudaStream_t)stream);
    if (e == cudaSuccess) e = cudaEventRecord(p->h_pin.done, (cudaStream_t)stream);
    if (e != cudaSuccess) {
        if (p->d_tables) cudaFreeAsync(p->d_tables, p->stream);
        pinned_put(p->h_pin);
        delete p;
        return fail(TQ_ECUDA, "tq_plan_create: %s", cudaGetErrorString(e));
    }
    const uint32_t *base = static_cast<const uint32_t *>(p->d_tables);
    p->dev.tile_off = base + o_off;
    p->dev.tile_k = base + o_k;
    p->dev.loc_start = base + o_start;
    p->dev.loc_len = base + o_len;
    p->dev.loc_run = base + o_run;
    p->dev.loc_pos = base + o_pos;
    *out = p;
    return TQ_OK;
}

int tq_plan_create(const uint32_t *h_locus, int64_t ncols, int64_t stride, int device, void *stream,
                   tq_plan **out)
{
    if (!out) return fail(TQ_EINVAL, "tq_plan_create: out is NULL");
    *out = nullptr;
    if (ncols < 0 || stride < 1 || (ncols > 0 && !h_locus))
        return fail(TQ_EINVAL, "tq_plan_create: bad arguments (ncols=%lld stride=%lld)",
                    (long long)ncols, (long long)stride);
    if (ncols >= ((int64_t)1 << 32)) return fail(TQ_EINVAL, "tq_plan_create: ncols must be < 2^32");

    // locus runs (maximal runs of equal id, resolve_quartets.py:58-64)
    std::vector<uint32_t> run_start, run_len;
    bool monotone = true;
    if (ncols > 0) {
        // one pass, run boundaries only (this loop is on the host entry's critical path: the pack waits for the plan)
        // branch-free: a run boundary every few columns is a mispredicted branch every few columns otherwise (4.4 -> 1 ns
        // per column)
        run_start.resize((size_t)ncols + 1);
        uint32_t *rs = run_start.data();
        uint32_t prev = h_locus[0];
        size_t n = 1;
        rs[0] = 0u;
        for (int64_t c = 1; c < ncols; ++c) {
            const uint32_t id = h_locus[c * stride];
            rs[n] = (uint32_t)c;
            n += id != prev;
            monotone &= id >= prev;
            prev = id;
        }
        run_start.resize(n);
        const size_t nr = n;
        run_len.resize(nr);
        for (size_t r = 0; r + 1 < nr; ++r) run_len[r] = run_start[r + 1] - run_start[r];
        run_len[nr - 1] = (uint32_t)ncols - run_start[nr - 1];
    }
    if (!monotone) {   // ids may come in any order, but every id must own exactly one run
        std::unordered_set<uint32_t> seen;
        seen.reserve(run_start.size() * 2);
        for (size_t r = 0; r < run_start.size(); ++r)
            if (!seen.insert(h_locus[(int64_t)run_start[r] * stride]).second)
                return fail(TQ_ELOCUS,
                            "locus id %u occurs in more than one run of columns; sort tmpmap by locus",
                            h_locus[(int64_t)run_start[r] * stride]);
    }
    return build_plan(run_start, run_len, ncols, device, stream, out);
}

int tq_plan_create_runs(const uint32_t *h_run_len, int64_t nruns, int device, void *stream, tq_plan **out)
{
    if (!out) return fail(TQ_EINVAL, "tq_plan_create_runs: out is NULL");
    *out = nullptr;
    if (nruns < 0 || (nruns > 0 && !h_run_len)) return fail(TQ_EINVAL, "tq_plan_create_runs: bad arguments");
    std::vector<uint32_t> run_start((size_t)nruns), run_len(h_run_len, h_run_len + nruns);
    int64_t col = 0;
    for (int64_t r = 0; r < nruns; ++r) {
        if (run_len[r] == 0) return fail(TQ_EINVAL, "tq_plan_create_runs: run %lld is empty", (long long)r);
        run_start[r] = (uint32_t)col;
        col += run_len[r];
        if (col >= ((int64_t)1 << 32)) return fail(TQ_EINVAL, "tq_plan_create_runs: more than 2^32 columns");
    }
    return build_plan(run_start, run_len, col, device, stream, out);
}

int tq_plan_tables(const tq_plan *plan, uint32_t *tile_off, uint32_t *tile_k, uint32_t *loc_start,
                   uint32_t *loc_len, uint32_t *loc_run)
{
    if (!plan) return fail(TQ_EINVAL, "tq_plan_tables: plan is NULL");
    const uint32_t *t = plan->h_tables();
    const size_t nt = (size_t)plan->ntiles, lp = nt * tq::kTileLoci;
    if (tile_off) memcpy(tile_off, t + plan->o_off, (nt + 1) * sizeof(uint32_t));
    if (tile_k) memcpy(tile_k, t + plan->o_k, nt * sizeof(uint32_t));
    if (loc_start) memcpy(loc_start, t + plan->o_start, lp * sizeof(uint32_t));
    if (loc_len) memcpy(loc_len, t + plan->o_len, lp * sizeof(uint32_t));
    if (loc_run) memcpy(loc_run, t + plan->o_run, lp * sizeof(uint32_t));
    return TQ_OK;
}

void tq_plan_destroy(tq_plan *plan)
{
    if (!plan) return;
    if (plan->d_tables) {                 // freed in stream order: work already enqueued on the plan's stream still sees the tables
        DeviceGuard guard(plan->device);
        cudaFreeAsync(plan->d_tables, plan->stream);
    }
    pinned_put(plan->h_pin);
    delete plan;
}

int tq_plan_info(const tq_plan *plan, int64_t info[8])
{
    if (!plan || !info) return fail(TQ_EINVAL, "tq_plan_info: NULL argument");
    info[0] = plan->ncols; info[1] = plan->nloci; info[2] = plan->ntiles;
    info[3] = plan->row_u4 * 4; info[4] = plan->max_len; info[5] = plan->padded_cols;
    info[6] = plan->device; info[7] = 0;
    return TQ_OK;
}

// ------------------------------------------------------------------------------------------
// device-buffer entries
// ------------------------------------------------------------------------------------------
int tq_pack_rows(const tq_plan *plan, const uint8_t *d_tmparr, int64_t ntips, int64_t row_stride,
                 int64_t tip_lo, int64_t tip_hi, uint32_t *d_planes, void *stream)
{
    if (!plan) return fail(TQ_EINVAL, "tq_pack: plan is NULL");
    if (plan->device < 0) return fail(TQ_EINVAL, "tq_pack: host-only plan (device < 0)");
    if (ntips < 0 || ntips > 65535) return fail(TQ_EINVAL, "tq_pack: ntips=%lld out of range", (long long)ntips);
    if (tip_lo < 0 || tip_hi < tip_lo || tip_hi > ntips)
        return fail(TQ_EINVAL, "tq_pack: tip rows [%lld, %lld) outside [0, %lld)", (long long)tip_lo, (long long)tip_hi, (long long)ntips);
    if (tip_hi == tip_lo || plan->row_u4 == 0) return TQ_OK;
    if (!d_tmparr || !d_planes) return fail(TQ_EINVAL, "tq_pack: NULL device pointer");
    if (row_stride < plan->ncols) return fail(TQ_EINVAL, "tq_pack: row_stride < ncols");
    if (((uintptr_t)d_planes & 15u) != 0) return fail(TQ_EINVAL, "tq_pack: d_planes must be 16-byte aligned");
    DeviceGuard guard(plan->device);
    TQ_CUDA(tq::launch_pack(plan->dev, d_tmparr, tip_lo, tip_hi, row_stride, d_planes, (cudaStream_t)stream));
    return TQ_OK;
}

int tq_pack(const tq_plan *plan, const uint8_t *d_tmparr, int64_t ntips, int64_t row_stride,
            uint32_t *d_planes, void *stream)
{
    return tq_pack_rows(plan, d_tmparr, ntips, row_stride, 0, ntips, d_planes, stream);
}

int tq_unrank(int device, const int64_t *d_ranks, int64_t start, int64_t m, int64_t ntips, uint32_t *d_quartets,
              void *stream)
{
    if (m < 0 || ntips < 4 || ntips > 65535) return fail(TQ_EINVAL, "tq_unrank: bad sizes (m=%lld ntips=%lld)", (long long)m, (long long)ntips);
    if (m == 0) return TQ_OK;
    if (!d_quartets || ((uintptr_t)d_quartets & 15u)) return fail(TQ_EINVAL, "tq_unrank: d_quartets must be a 16-byte aligned device pointer");
    const uint64_t n = (uint64_t)ntips, total = n * (n - 1) / 2 * (n - 2) / 3 * (n - 3) / 4;
    if (!d_ranks && (start < 0 || (uint64_t)start + (uint64_t)m > total))
        return fail(TQ_EINVAL, "tq_unrank: range [%lld, %lld) exceeds C(%lld,4)", (long long)start, (long long)(start + m), (long long)ntips);
    DeviceGuard guard(device);
    TQ_CUDA(tq::launch_unrank(d_ranks, start, m, (uint32_t)ntips, d_quartets, (cudaStream_t)stream));
    return TQ_OK;
}

int tq_sample_quartets(int device, int64_t start, int64_t m, int64_t ntips, uint64_t key, uint32_t *d_quartets, void *stream)
{
    if (m < 0 || start < 0 || ntips < 4 || ntips > 65535)
        return fail(TQ_EINVAL, "tq_sample_quartets: bad sizes (start=%lld m=%lld ntips=%lld)", (long long)start, (long long)m, (long long)ntips);
    if (m == 0) return TQ_OK;
    if (!d_quartets || ((uintptr_t)d_quartets & 15u)) return fail(TQ_EINVAL, "tq_sample_quartets: d_quartets must be a 16-byte aligned device pointer");
    const uint64_t n = (uint64_t)ntips, total = n * (n - 1) / 2 * (n - 2) / 3 * (n - 3) / 4;
    if (total >= ((uint64_t)1 << 32)) return fail(TQ_EINVAL, "tq_sample_quartets: C(%lld,4) >= 2^32 (the reference asserts the same)", (long long)ntips);
    if ((uint64_t)start + (uint64_t)m > total)
        return fail(TQ_EINVAL, "tq_sample_quartets: sample indices [%lld, %lld) exceed C(%lld,4)", (long long)start, (long long)(start + m), (long long)ntips);
    DeviceGuard guard(device);
    TQ_CUDA(tq::launch_sample_unrank(start, m, (uint32_t)ntips, key, d_quartets, (cudaStream_t)stream));
    return TQ_OK;
}

int tq_sample_ranks_host(int64_t start, int64_t m, int64_t ntips, uint64_t key, int64_t *h_ranks)
{
    if (m < 0 || start < 0 || ntips < 4 || ntips > 65535 || (m > 0 && !h_ranks)) return fail(TQ_EINVAL, "tq_sample_ranks_host: bad arguments");
    const uint64_t n = (uint64_t)ntips, total = n * (n - 1) / 2 * (n - 2) / 3 * (n - 3) / 4;
    if (total >= ((uint64_t)1 << 32) || (uint64_t)start + (uint64_t)m > total)
        return fail(TQ_EINVAL, "tq_sample_ranks_host: sample indices [%lld, %lld) exceed C(%lld,4)", (long long)start, (long long)(start + m), (long long)ntips);
    tq::sample_ranks_host(start, m, (uint32_t)ntips, key, h_ranks);
    return TQ_OK;
}

int tq_boot_pack_rows(const tq_plan *plan, const uint8_t *d_seqarr, int64_t ntips, int64_t row_stride,
                      int64_t tip_lo, int64_t tip_hi, const uint32_t *d_src_start, uint32_t seed_shuffle,
                      uint32_t seed_ambig, uint32_t *d_srccol, uint32_t *d_planes, void *stream)
{
    if (!plan) return fail(TQ_EINVAL, "tq_boot_pack: plan is NULL");
    if (plan->device < 0) return fail(TQ_EINVAL, "tq_boot_pack: host-only plan (device < 0)");
    if (ntips < 0 || ntips > 65535) return fail(TQ_EINVAL, "tq_boot_pack: ntips=%lld out of range", (long long)ntips);
    if (tip_lo < 0 || tip_hi < tip_lo || tip_hi > ntips)
        return fail(TQ_EINVAL, "tq_boot_pack: tip rows [%lld, %lld) outside [0, %lld)", (long long)tip_lo, (long long)tip_hi, (long long)ntips);
    if (ntips == 0 || plan->row_u4 == 0) return TQ_OK;
    if (!d_seqarr || !d_src_start || !d_srccol || !d_planes) return fail(TQ_EINVAL, "tq_boot_pack: NULL device pointer");
    if (((uintptr_t)d_planes & 15u) != 0) return fail(TQ_EINVAL, "tq_boot_pack: d_planes must be 16-byte aligned");
    DeviceGuard guard(plan->device);
    TQ_CUDA(tq::launch_boot_pack(plan->dev, d_seqarr, tip_lo, tip_hi, row_stride, d_src_start, seed_shuffle, seed_ambig,
                                 d_srccol, d_planes, (cudaStream_t)stream));
    return TQ_OK;
}

int tq_boot_pack(const tq_plan *plan, const uint8_t *d_seqarr, int64_t ntips, int64_t row_stride,
                 const uint32_t *d_src_start, uint32_t seed_shuffle, uint32_t seed_ambig,
                 uint32_t *d_srccol, uint32_t *d_planes, void *stream)
{
    return tq_boot_pack_rows(plan, d_seqarr, ntips, row_stride, 0, ntips, d_src_start, seed_shuffle, seed_ambig,
                             d_srccol, d_planes, stream);
}

int tq_resolve(const tq_plan *plan, const uint32_t *d_planes, int64_t ntips,
               const uint32_t *d_quartets, int64_t m, int subsample, uint64_t seed,
               double *d_rscor, uint32_t *d_rstat, uint32_t *d_counts, uint32_t *d_minrank,
               uint32_t *d_status, void *stream)
{
    if (!plan) return fail(TQ_EINVAL, "tq_resolve: plan is NULL");
    if (plan->device < 0) return fail(TQ_EINVAL, "tq_resolve: host-only plan (device < 0)");
    if (m < 0 || ntips < 0) return fail(TQ_EINVAL, "tq_resolve: negative size");
    if (m == 0) return TQ_OK;
    if (!d_quartets || !d_rscor || !d_rstat) return fail(TQ_EINVAL, "tq_resolve: NULL device pointer");
    if (plan->row_u4 > 0 && !d_planes) return fail(TQ_EINVAL, "tq_resolve: d_planes is NULL");
    if ((uint64_t)ntips * (uint64_t)plan->row_u4 >= ((uint64_t)1 << 32))
        return fail(TQ_EINVAL, "tq_resolve: packed matrix exceeds 2^32 x 16 bytes (ntips * row_words too large)");
    if (((uintptr_t)d_planes & 15u) || ((uintptr_t)d_quartets & 15u) || ((uintptr_t)d_rstat & 7u))
        return fail(TQ_EINVAL, "tq_resolve: planes/quartets must be 16-byte and rstat 8-byte aligned");
    DeviceGuard guard(plan->device);
    TQ_CUDA(tq::launch_resolve(plan->dev, d_planes, ntips, d_quartets, m, subsample != 0, seed, d_rscor,
                               d_rstat, d_counts, d_minrank, d_status, plan->num_sms,
                               (cudaStream_t)stream));
    return TQ_OK;
}

// ------------------------------------------------------------------------------------------
// host-buffer entry
// ------------------------------------------------------------------------------------------
int tq_ctx_create(int device, tq_ctx **out)
{
    if (!out) return fail(TQ_EINVAL, "tq_ctx_create: out is NULL");
    *out = nullptr;
    int ndev = 0;
    TQ_CUDA(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail(TQ_EINVAL, "tq_ctx_create: no CUDA device %d", device);
    tq_ctx *c = new (std::nothrow) tq_ctx();
    if (!c) return fail(TQ_ENOMEM, "tq_ctx_create: out of host memory");
    c->device = device;
    DeviceGuard guard(device);
    cudaError_t e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking);
    for (int i = 0; i < kHostChunks && e == cudaSuccess; ++i) e = cudaEventCreateWithFlags(&c->chunk_done[i], cudaEventDisableTiming);
    if (e != cudaSuccess) { tq_ctx_destroy(c); return fail(TQ_ECUDA, "tq_ctx_create: %s", cudaGetErrorString(e)); }
    *out = c;
    return TQ_OK;
}

void tq_ctx_destroy(tq_ctx *c)
{
    if (!c) return;
    DeviceGuard guard(c->device);
    if (c->plan) tq_plan_destroy(c->plan);
    cudaFree(c->d_tmparr); cudaFree(c->d_planes); cudaFree(c->d_quart); cudaFree(c->d_out);
    for (int i = 0; i < kHostChunks; ++i)
        if (c->chunk_done[i]) cudaEventDestroy(c->chunk_done[i]);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    if (c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

int tq_ctx_last_traffic(const tq_ctx *ctx, int64_t out[2])
{
    if (!ctx || !out) return fail(TQ_EINVAL, "tq_ctx_last_traffic: NULL argument");
    out[0] = ctx->last_h2d; out[1] = ctx->last_d2h;
    return TQ_OK;
}

static int grow(void **buf, size_t *cap, size_t need)
{
    if (need <= *cap) return TQ_OK;
    if (*buf) cudaFree(*buf);
    *buf = nullptr; *cap = 0;
    const size_t want = need + need / 8 + 256;
    TQ_CUDA(cudaMalloc(buf, want));
    *cap = want;
    return TQ_OK;
}

int tq_resolve_host(tq_ctx *c, const uint8_t *h_tmparr, int64_t ntips, int64_t ncols, int64_t row_stride,
                    const uint32_t *h_locus, int64_t locus_stride, const uint32_t *h_quartets, int64_t m,
                    int subsample, uint64_t seed, uint64_t epoch, double *h_rscor, uint32_t *h_rstat,
                    uint32_t *h_counts, uint32_t *h_minrank)
{
    if (!c) return fail(TQ_EINVAL, "tq_resolve_host: ctx is NULL");
    if (ntips < 0 || ncols < 0 || m < 0 || row_stride < ncols || locus_stride < 1)
        return fail(TQ_EINVAL, "tq_resolve_host: bad sizes");
    if (ntips > 65535) return fail(TQ_EINVAL, "tq_resolve_host: ntips=%lld > 65535", (long long)ntips);
    if ((ncols > 0 && ntips > 0 && !h_tmparr) || (ncols > 0 && !h_locus))
        return fail(TQ_EINVAL, "tq_resolve_host: NULL input");
    if (m > 0 && (!h_quartets || !h_rscor || !h_rstat)) return fail(TQ_EINVAL, "tq_resolve_host: NULL quartets/outputs");
    c->last_h2d = c->last_d2h = 0;
    if (m == 0) return TQ_OK;
    DeviceGuard guard(c->device);
    int rc;

    // (re)build the replicate state unless the caller presents the same arrays again
    const bool same = c->plan && c->key_arr == h_tmparr && c->key_loc == h_locus && c->key_ntips == ntips &&
                      c->key_ncols == ncols && c->key_epoch == epoch && epoch != 0;
    // Host -> device copies that do not depend on the plan go first, so that they run while the host builds it.
    if ((rc = grow(&c->d_quart, &c->cap_quart, (size_t)m * 16)) != TQ_OK) return rc;
    TQ_CUDA(cudaMemcpyAsync(c->d_quart, h_quartets, (size_t)m * 16, cudaMemcpyHostToDevice, c->stream));
    c->last_h2d += m * 16;
    if (!same) {
        c->key_arr = c->key_loc = nullptr;          // nothing is cached until the rebuild below has succeeded
        const size_t arr_bytes = (size_t)ntips * (size_t)ncols;
        if ((rc = grow(&c->d_tmparr, &c->cap_tmparr, arr_bytes + 16)) != TQ_OK) return rc;
        if (arr_bytes) {
            TQ_CUDA(cudaMemcpy2DAsync(c->d_tmparr, (size_t)ncols, h_tmparr, (size_t)row_stride, (size_t)ncols,
                                      (size_t)ntips, cudaMemcpyHostToDevice, c->stream));
            c->last_h2d += (int64_t)arr_bytes;
        }
        if (c->plan) { tq_plan_destroy(c->plan); c->plan = nullptr; }
        if ((rc = tq_plan_create(h_locus, ncols, locus_stride, c->device, c->stream, &c->plan)) != TQ_OK) return rc;
        c->last_h2d += (int64_t)((2 * c->plan->ntiles + 2 + 4 * c->plan->ntiles * tq::kTileLoci) * 4);
        if ((rc = grow(&c->d_planes, &c->cap_planes, (size_t)ntips * (size_t)c->plan->row_u4 * 16 + 16)) != TQ_OK) return rc;
        if ((rc = tq_pack(c->plan, (const uint8_t *)c->d_tmparr, ntips, ncols, (uint32_t *)c->d_planes, c->stream)) != TQ_OK)
            return rc;
        c->key_arr = h_tmparr; c->key_loc = h_locus; c->key_ntips = ntips; c->key_ncols = ncols;
        c->key_epoch = epoch;
    }

    // outputs: rscor f64[m][3] | rstat u32[m][2] | minrank u32[m] | status u32[4] | counts u32[m][256]
    const size_t o_scor = 0, o_stat = o_scor + (size_t)m * 24, o_mr = o_stat + (size_t)m * 8,
                 o_status = o_mr + (((size_t)m * 4 + 15) & ~(size_t)15),
                 o_counts = o_status + 16, out_bytes = o_counts + (h_counts ? (size_t)m * 1024 : 0);
    if ((rc = grow(&c->d_out, &c->cap_out, out_bytes)) != TQ_OK) return rc;
    char *o = static_cast<char *>(c->d_out);
    TQ_CUDA(cudaMemsetAsync(o + o_status, 0, 16, c->stream));
    // Subsample mode, large lists: the list is resolved in kHostChunks launches and every chunk's results travel back
    // on the copy stream while the next chunk is being resolved, so only the last chunk's copy is exposed.  (Full mode
    // keeps one launch: the tip-pair GEMM is chosen by the density of the WHOLE list.)
    const int nch = (subsample && m >= 65536) ? kHostChunks : 1;
    for (int ch = 0; ch < nch; ++ch) {
        const int64_t lo = m * ch / nch, hi = m * (ch + 1) / nch;
        if ((rc = tq_resolve(c->plan, (const uint32_t *)c->d_planes, ntips, (const uint32_t *)c->d_quart + 4 * lo, hi - lo,
                             subsample, seed, (double *)(o + o_scor) + 3 * lo, (uint32_t *)(o + o_stat) + 2 * lo,
                             h_counts ? (uint32_t *)(o + o_counts) + 256 * lo : nullptr, (uint32_t *)(o + o_mr) + lo,
                             (uint32_t *)(o + o_status), c->stream)) != TQ_OK)
            return rc;
        TQ_CUDA(cudaEventRecord(c->chunk_done[ch], c->stream));
    }
    for (int ch = 0; ch < nch; ++ch) {
        const int64_t lo = m * ch / nch, hi = m * (ch + 1) / nch, n = hi - lo;
        TQ_CUDA(cudaStreamWaitEvent(c->copy_stream, c->chunk_done[ch], 0));
        TQ_CUDA(cudaMemcpyAsync(h_rscor + 3 * lo, o + o_scor + lo * 24, (size_t)n * 24, cudaMemcpyDeviceToHost, c->copy_stream));
        TQ_CUDA(cudaMemcpyAsync(h_rstat + 2 * lo, o + o_stat + lo * 8, (size_t)n * 8, cudaMemcpyDeviceToHost, c->copy_stream));
        if (h_minrank) TQ_CUDA(cudaMemcpyAsync(h_minrank + lo, o + o_mr + lo * 4, (size_t)n * 4, cudaMemcpyDeviceToHost, c->copy_stream));
        if (h_counts) TQ_CUDA(cudaMemcpyAsync(h_counts + 256 * lo, o + o_counts + lo * 1024, (size_t)n * 1024, cudaMemcpyDeviceToHost, c->copy_stream));
    }
    c->last_d2h += m * 32 + (h_minrank ? m * 4 : 0) + (h_counts ? m * 1024 : 0);
    // The tip indices are checked on the host WHILE the device works (the kernels flag an index out of range themselves
    // and touch nothing outside the matrix, so launching first is safe); the error, if any, is reported after the queue
    // has drained, and the outputs of a failed call mean nothing.
    int64_t bad_q = -1;
    for (int64_t q = 0; q < m && bad_q < 0; ++q)
        for (int i = 0; i < 4; ++i)
            if ((int64_t)h_quartets[4 * q + i] >= ntips) { bad_q = q; break; }
    TQ_CUDA(cudaStreamSynchronize(c->copy_stream));
    TQ_CUDA(cudaStreamSynchronize(c->stream));
    if (bad_q >= 0) {
        uint32_t t = 0;
        for (int i = 0; i < 4; ++i)
            if ((int64_t)h_quartets[4 * bad_q + i] >= ntips) { t = h_quartets[4 * bad_q + i]; break; }
        return fail(TQ_EINVAL, "tq_resolve_host: quartet %lld has tip index %u >= ntips %lld", (long long)bad_q, t, (long long)ntips);
    }
    return TQ_OK;
}

int tq_microbench(int device, int kind, double *result)
{
    if (!result || kind < 0 || (kind > 4 && (kind < 10 || kind > 14))) return fail(TQ_EINVAL, "tq_microbench: bad arguments");
    DeviceGuard guard(device);
    TQ_CUDA(tq::run_microbench(kind, result));
    return TQ_OK;
}

}  // extern "C"
