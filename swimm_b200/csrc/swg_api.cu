// swg_api.cu -- the C ABI of libswimm_cuda.so (include/swimm_gpu.h): context, resident database,
// query upload, kernel scheduling, top-r, statistics, and the reference-signature entry point.
//
// One context drives one GPU.  swg_gpu_run plans the batch (swg_plan.cu), then enqueues without waiting on the host:
//     single queries      profile build (K0) -> wavefront<Lane16, G, K> (K1) -> 32-bit recomputation of its list (K2)
//     groups of queries   per launch: pair profile -> wavefront_q2<G, K> (K1q); K2 when a query ends
//     long tiles          split off per work item and run beside the main launches on the high-priority stream:
//                         wavefront_xw (K3), or -- one-pass queries -- the 32-thread shape of K1 on column chunks
//     per chunk of queries  top-r selection (K4)
// swg_gpu_submit / swg_gpu_poll run the same with two query-buffer sets, so that two batches can be in flight.
#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>
#include <set>
#include <string>
#include <vector>

#include "../../include/swimm_gpu.h"
#include "swg_internal.h"
#include "swg_plan.h"
#include "wavefront.cuh"
#include "wavefront_q2.cuh"
#include "wavefront_xw.cuh"

using namespace swg;

namespace {

constexpr uint32_t kDefaultLongCols = 0;       // 0: the long-tile threshold is estimated per query / launch
constexpr uint64_t kMaxLongBlocks = 12;

struct DeviceBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes)
    {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        cudaError_t e = cudaMalloc(&p, bytes);
        if (e == cudaSuccess) cap = bytes;
        return e;
    }
    void release()
    {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <class T>
    T *as() const { return reinterpret_cast<T *>(p); }
};

}  // namespace

struct swg_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t side_stream = nullptr;      // carries the main search kernel while long tiles run on `stream`
    cudaEvent_t ev_begin = nullptr, ev_end = nullptr, ev_fork = nullptr, ev_join = nullptr;
    char err[512] = {0};

    // resident database shard
    bool db_ready = false;
    uint64_t n_total = 0;            // sequences of the whole database
    uint32_t shard = 0, num_shards = 1;
    uint32_t ntiles = 0;             // local tiles
    uint32_t first_long_tile = 0;    // local tiles [first_long_tile, ntiles) have more than long_cols columns
    uint64_t local_seqs = 0, local_residues = 0, db_units = 0;
    uint32_t maxcols = 8;
    double avg_cols = 8;
    std::vector<uint32_t> h_tile_cols;
    std::vector<uint64_t> h_cols_prefix;      // [ntiles + 1] running sum of h_tile_cols
    std::vector<uint32_t> h_tile_cols_mono;   // running maximum of h_tile_cols (ascending even when trailing tiles are padding)
    DeviceBuf d_db, d_tile_off, d_tile_cols, d_line_off;
    uint64_t line_units = 0;         // uint2 entries of all per-sequence pass lines (query-pair kernel, multi-pass)

    // queries of the current batch
    bool queries_ready = false;
    uint64_t q_count = 0;
    std::vector<uint16_t> q_len;
    std::vector<uint32_t> q_off;     // offsets into d_queries
    int open_gap = 10, extend_gap = 2;
    // the buffer set of a batch (swapped with a slot's by swg_gpu_submit): queries + matrix on the device, the pinned
    // staging they are uploaded from, the hit lists, and the two events that guard their reuse
    DeviceBuf d_queries, d_submat;
    char *h_stage = nullptr;         // pinned staging of the query upload
    size_t h_stage_cap = 0;
    cudaEvent_t ev_upload = nullptr;     // recorded on copy_stream when the upload from h_stage has completed
    cudaEvent_t ev_bufs_free = nullptr;  // recorded on stream behind the last run that read d_queries / wrote d_top_out
    cudaStream_t copy_stream = nullptr;
    struct QuerySlot {
        DeviceBuf d_queries, d_submat, d_top_out;
        char *h_stage = nullptr;
        size_t h_stage_cap = 0;
        cudaEvent_t ev_upload = nullptr, ev_bufs_free = nullptr, ev_begin = nullptr, ev_done = nullptr;
        uint64_t *h_keys = nullptr;      // pinned: the batch's hit lists, [q_count][top]
        size_t h_keys_cap = 0;
        bool busy = false;
        int ticket = -1;
        uint64_t q_count = 0, top = 0, top_stride = 0;
    } slots[2];
    int next_ticket = 0;
    cudaEvent_t begin_mark = nullptr;    // swg_gpu_submit: the slot's begin event, recorded by swg_gpu_run beside ev_begin

    // work buffers
    DeviceBuf d_scores, d_profile, d_profile32, d_boundary, d_counters, d_resc_list, d_topk_scratch, d_top_out;
    DeviceBuf d_profile_q2, d_lines, d_resc_list2, d_q2_counters;
    DeviceBuf d_profile_xw, d_xw_counters, d_xw_list;
    // column-chunk tables of the long tiles (uint4 per chunk), all queries of a run; two sets, used in turn, so that the
    // next batch's tables can be staged while the running batch still reads its own (swg_gpu_submit)
    DeviceBuf d_vt[2];
    uint4 *h_vt[2] = {nullptr, nullptr};     // pinned staging of the same
    size_t h_vt_cap[2] = {0, 0};
    cudaEvent_t ev_vt[2] = {nullptr, nullptr};   // recorded behind the run that read set k
    int vt_set = 0;
    int submat_max = 0;                      // largest entry of the current substitution matrix
    DeviceBuf d_q_off, d_align_lines, d_coords;                      // coordinate pass (align_ends.cu)     // long-sequence kernel (wavefront_xw.cuh)
    std::vector<WorkItem> items;            // schedule of the last run
    std::vector<cudaEvent_t> item_events;   // items.size() + 1 marks
    std::vector<cudaEvent_t> chunk_events;  // two per chunk of queries, around its top-r selection
    uint64_t run_chunks = 0, run_window = 0;
    long score_budget = 4l << 30;           // bytes of score rows a run may keep when the caller wants hit lists only
    std::vector<uint32_t> h_counters;
    std::vector<cudaEvent_t> q_events;      // q_count + 1 marks around every query's kernels
    std::vector<double> q_seconds;
    std::vector<uint32_t> warmed;           // kernel instantiations already loaded on this device (see warm_kernels)
    uint64_t run_top = 0, run_top_stride = 0;
    bool run_done = false, run_kept_scores = false;

    // options
    long long_cols = kDefaultLongCols;
    long xw_warps = 0, xw_rows = 0;         // forced shape of the long-sequence kernel (0: planner)
    long trace = 0;                         // 1: an event after every launch of a run, timeline printed by swg_gpu_sync
    std::vector<std::pair<std::string, cudaEvent_t>> trace_marks;
    std::vector<cudaEvent_t> trace_pool;
    long chunk_columns = 0;                 // column chunks of long tiles for one-pass queries: 0 planner, 1 off, > 1 this many columns
    long long_kernel = 1;                   // 1: long tiles run the cross-warp wavefront kernel (K3), 0: the 32-thread shape of K1
    long force_group = 0, force_rows = 0;
    long query_pairing = 1;                 // 0: never pair queries, 1: pair when the planner expects a gain, 2: always
    long q2_group = 0, q2_rows = 0;         // forced shape of the query-pair kernel (0: planner)
    long verbose = 0;                       // 1: print the schedule of every run to stderr
    long pass_lines = 1;                    // 0: never allocate pass lines (8 B per database column): single-launch pairs only
    long grid_blocks = 0;                   // CTAs per search launch (0: one per SM); small values make every warp run many tasks

    swg_stats stats;
};

namespace {

const char *status_text(int st)
{
    switch (st) {
        case SWG_OK: return "ok";
        case SWG_ERR_NO_DEVICE: return "no CUDA device";
        case SWG_ERR_CUDA: return "CUDA error";
        case SWG_ERR_ARG: return "bad argument";
        case SWG_ERR_STATE: return "call out of order";
        case SWG_ERR_NOMEM: return "out of memory";
        default: return "unknown";
    }
}

thread_local char g_last_error[512] = "no error";      // per host thread: contexts may be driven concurrently

int fail(swg_ctx *ctx, int st, const char *fmt, ...)
{
    char buf[400];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    snprintf(g_last_error, sizeof(g_last_error), "%s: %s", status_text(st), buf);
    if (ctx) snprintf(ctx->err, sizeof(ctx->err), "%s: %s", status_text(st), buf);
    return st;
}

int cuda_fail(swg_ctx *ctx, cudaError_t e, const char *what)
{
    const int st = (e == cudaErrorMemoryAllocation) ? SWG_ERR_NOMEM
                   : (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) ? SWG_ERR_NO_DEVICE
                                                                                  : SWG_ERR_CUDA;
    return fail(ctx, st, "%s: %s", what, cudaGetErrorString(e));
}

#define SWG_CUDA(ctx, call)                                         \
    do {                                                            \
        cudaError_t e__ = (call);                                   \
        if (e__ != cudaSuccess) return cuda_fail(ctx, e__, #call);  \
    } while (0)

cudaError_t launch_q2(int G, int K, bool cin, bool cout, int grid, cudaStream_t stream, const WfParams &p)
{
    if (!cin && !cout) {
        switch (G) {
            case 8: return launch_q2_g8(K, grid, stream, p);
            case 16: return launch_q2_g16(K, grid, stream, p);
            case 32: return launch_q2_g32(K, grid, stream, p);
            default: return cudaErrorInvalidValue;
        }
    }
    if (G != 32) return cudaErrorInvalidValue;
    if (!cin) return launch_q2_g32_first(K, grid, stream, p);
    return cout ? launch_q2_g32_middle(K, grid, stream, p) : launch_q2_g32_last(K, grid, stream, p);
}

cudaError_t launch_wavefront(bool lane32, const Config &cfg, int grid, cudaStream_t stream, const WfParams &p)
{
    if (cfg.global_profile) return lane32 ? launch_wf_l32_gp(grid, stream, p) : launch_wf_l16_gp(grid, stream, p);
    const size_t smem = (size_t)cfg.passes * kPassBytes;
    if (cfg.passes > 1) return lane32 ? launch_wf_l32_g32_mp(cfg.K, grid, smem, stream, p)
                                      : launch_wf_l16_g32_mp(cfg.K, grid, smem, stream, p);
    if (lane32) return launch_wf_l32_g32(cfg.K, grid, smem, stream, p);
    switch (cfg.G) {
        case 4: return launch_wf_l16_g4(cfg.K, grid, smem, stream, p);
        case 8: return launch_wf_l16_g8(cfg.K, grid, smem, stream, p);
        case 16: return launch_wf_l16_g16(cfg.K, grid, smem, stream, p);
        case 32: return launch_wf_l16_g32(cfg.K, grid, smem, stream, p);
        default: return cudaErrorInvalidValue;
    }
}

void free_db(swg_ctx *c)
{
    c->d_db.release();
    c->d_tile_off.release();
    c->d_tile_cols.release();
    c->d_line_off.release();
    c->d_lines.release();
    c->db_ready = false;
}

// Shared tail of the two database loaders: `lengths`/`offsets` describe the LOCAL sequences (padded to a
// multiple of 16 with zero-length entries), d_res holds their residues back to back.
int finish_load(swg_ctx *ctx, const std::vector<uint16_t> &len, const std::vector<uint64_t> &off, const int8_t *d_res)
{
    const uint32_t ntiles = (uint32_t)(len.size() / kTileSeqs);
    std::vector<uint64_t> tile_off(ntiles + 1, 0), line_off(ntiles + 1, 0);
    ctx->h_tile_cols.assign(ntiles, 0);
    uint32_t maxcols = kChunkCols;
    double sum_cols = 0;
    for (uint32_t t = 0; t < ntiles; ++t) {
        uint32_t mx = 1;
        for (int k = 0; k < kTileSeqs; ++k) mx = std::max<uint32_t>(mx, len[(size_t)t * kTileSeqs + k]);
        const uint32_t cols = (mx + kChunkCols - 1) / kChunkCols * kChunkCols;
        ctx->h_tile_cols[t] = cols;
        tile_off[t + 1] = tile_off[t] + (uint64_t)(cols / kChunkCols) * kTilePairs;
        // pass lines of the query-pair kernel: 16 per tile, each max(cols, kQ2MinSegCols) + kQ2LineSlack entries
        line_off[t + 1] = line_off[t] + (uint64_t)kTileSeqs * (std::max<uint32_t>(cols, kQ2MinSegCols) + kQ2LineSlack);
        maxcols = std::max(maxcols, cols);
        sum_cols += cols;
    }
    ctx->ntiles = ntiles;
    ctx->maxcols = maxcols;
    ctx->avg_cols = ntiles ? sum_cols / ntiles : 8.0;
    ctx->db_units = tile_off[ntiles];
    ctx->line_units = line_off[ntiles];
    ctx->h_cols_prefix.assign(ntiles + 1, 0);
    for (uint32_t t = 0; t < ntiles; ++t) ctx->h_cols_prefix[t + 1] = ctx->h_cols_prefix[t] + ctx->h_tile_cols[t];
    ctx->h_tile_cols_mono = ctx->h_tile_cols;
    for (uint32_t t = 1; t < ntiles; ++t)
        ctx->h_tile_cols_mono[t] = std::max(ctx->h_tile_cols_mono[t], ctx->h_tile_cols_mono[t - 1]);
    ctx->first_long_tile = ntiles;
    if (ctx->long_cols > 0)         // lengths ascend: long tiles are last
        ctx->first_long_tile = (uint32_t)(std::upper_bound(ctx->h_tile_cols_mono.begin(), ctx->h_tile_cols_mono.end(),
                                                           (uint32_t)ctx->long_cols) - ctx->h_tile_cols_mono.begin());

    DeviceBuf d_off, d_len;
    SWG_CUDA(ctx, ctx->d_db.reserve(std::max<uint64_t>(ctx->db_units, 1) * sizeof(uint4)));
    SWG_CUDA(ctx, ctx->d_tile_off.reserve((ntiles + 1) * sizeof(uint64_t)));
    SWG_CUDA(ctx, ctx->d_tile_cols.reserve(std::max<uint32_t>(ntiles, 1) * sizeof(uint32_t)));
    SWG_CUDA(ctx, ctx->d_line_off.reserve((ntiles + 1) * sizeof(uint64_t)));
    SWG_CUDA(ctx, d_off.reserve(off.size() * sizeof(uint64_t)));
    SWG_CUDA(ctx, d_len.reserve(std::max<size_t>(len.size(), 1) * sizeof(uint16_t)));
    cudaError_t e = cudaMemcpyAsync(ctx->d_tile_off.p, tile_off.data(), (ntiles + 1) * sizeof(uint64_t),
                                    cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(ctx->d_line_off.p, line_off.data(), (ntiles + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice,
                            ctx->stream);
    if (e == cudaSuccess && ntiles)
        e = cudaMemcpyAsync(ctx->d_tile_cols.p, ctx->h_tile_cols.data(), ntiles * sizeof(uint32_t), cudaMemcpyHostToDevice,
                            ctx->stream);
    if (e == cudaSuccess)
        e = cudaMemcpyAsync(d_off.p, off.data(), off.size() * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess && !len.empty())
        e = cudaMemcpyAsync(d_len.p, len.data(), len.size() * sizeof(uint16_t), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess)
        e = launch_build_tiles(d_res, d_off.as<uint64_t>(), d_len.as<uint16_t>(), ctx->d_tile_off.as<uint64_t>(), ntiles,
                               ctx->db_units, ctx->d_db.as<uint4>(), ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    d_off.release();
    d_len.release();
    if (e != cudaSuccess) return cuda_fail(ctx, e, "database layout build");
    ctx->stats.db_bytes = ctx->db_units * sizeof(uint4);
    ctx->db_ready = true;
    ctx->run_done = false;
    return SWG_OK;
}

}  // namespace

// =================================================================================================
extern "C" {

int swg_gpu_device_count(int *count)
{
    if (!count) return fail(nullptr, SWG_ERR_ARG, "count is NULL");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { *count = 0; return cuda_fail(nullptr, e, "cudaGetDeviceCount"); }
    *count = n;
    return n > 0 ? SWG_OK : fail(nullptr, SWG_ERR_NO_DEVICE, "no CUDA device visible");
}

int swg_gpu_create(int device, swg_ctx **out)
{
    if (!out) return fail(nullptr, SWG_ERR_ARG, "ctx is NULL");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) return cuda_fail(nullptr, e, "cudaGetDeviceCount");
    if (n <= 0) return fail(nullptr, SWG_ERR_NO_DEVICE, "no CUDA device visible (there is no CPU fallback)");
    if (device < 0 || device >= n) return fail(nullptr, SWG_ERR_ARG, "device %d out of range (0..%d)", device, n - 1);
    swg_ctx *ctx = new (std::nothrow) swg_ctx();
    if (!ctx) return fail(nullptr, SWG_ERR_NOMEM, "context");
    ctx->device = device;
    memset(&ctx->stats, 0, sizeof(ctx->stats));
    e = cudaSetDevice(device);
    cudaDeviceProp prop;
    if (e == cudaSuccess) e = cudaGetDeviceProperties(&prop, device);
    if (e == cudaSuccess && prop.major < 10) {
        delete ctx;
        return fail(nullptr, SWG_ERR_NO_DEVICE, "device %d is sm_%d%d; this library is built for sm_100a only", device,
                    prop.major, prop.minor);
    }
    if (e == cudaSuccess) ctx->sm_count = prop.multiProcessorCount;
    int prio_lo = 0, prio_hi = 0;
    if (e == cudaSuccess) e = cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    if (e == cudaSuccess) e = cudaStreamCreateWithPriority(&ctx->stream, cudaStreamNonBlocking, prio_hi);
    if (e == cudaSuccess) e = cudaStreamCreateWithPriority(&ctx->side_stream, cudaStreamNonBlocking, prio_lo);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_upload, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_bufs_free, cudaEventDisableTiming);
    for (int k = 0; k < 2 && e == cudaSuccess; ++k) {
        e = cudaEventCreateWithFlags(&ctx->slots[k].ev_upload, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->slots[k].ev_bufs_free, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreate(&ctx->slots[k].ev_begin);
        if (e == cudaSuccess) e = cudaEventCreate(&ctx->slots[k].ev_done);
    }
    if (e == cudaSuccess) e = cudaEventCreate(&ctx->ev_begin);
    if (e == cudaSuccess) e = cudaEventCreate(&ctx->ev_end);
    if (e != cudaSuccess) {
        const int st = cuda_fail(nullptr, e, "context creation");
        delete ctx;
        return st;
    }
    *out = ctx;
    return SWG_OK;
}

void swg_gpu_destroy(swg_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->copy_stream) cudaStreamSynchronize(ctx->copy_stream);
    if (ctx->side_stream) cudaStreamSynchronize(ctx->side_stream);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    free_db(ctx);
    for (int k = 0; k < 2; ++k) {
        swg_ctx::QuerySlot &sl = ctx->slots[k];
        sl.d_queries.release();
        sl.d_submat.release();
        sl.d_top_out.release();
        if (sl.h_stage) cudaFreeHost(sl.h_stage);
        if (sl.h_keys) cudaFreeHost(sl.h_keys);
        for (cudaEvent_t ev : {sl.ev_upload, sl.ev_bufs_free, sl.ev_begin, sl.ev_done})
            if (ev) cudaEventDestroy(ev);
    }
    if (ctx->ev_upload) cudaEventDestroy(ctx->ev_upload);
    if (ctx->ev_bufs_free) cudaEventDestroy(ctx->ev_bufs_free);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    ctx->d_queries.release();
    ctx->d_submat.release();
    if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
    ctx->d_scores.release();
    ctx->d_profile.release();
    ctx->d_profile32.release();
    ctx->d_boundary.release();
    ctx->d_counters.release();
    ctx->d_resc_list.release();
    ctx->d_topk_scratch.release();
    ctx->d_top_out.release();
    ctx->d_profile_q2.release();
    ctx->d_resc_list2.release();
    ctx->d_q2_counters.release();
    ctx->d_profile_xw.release();
    ctx->d_xw_counters.release();
    ctx->d_xw_list.release();
    for (int k = 0; k < 2; ++k) {
        ctx->d_vt[k].release();
        if (ctx->h_vt[k]) cudaFreeHost(ctx->h_vt[k]);
        if (ctx->ev_vt[k]) cudaEventDestroy(ctx->ev_vt[k]);
    }
    ctx->d_q_off.release();
    ctx->d_align_lines.release();
    ctx->d_coords.release();
    for (cudaEvent_t ev : ctx->q_events) cudaEventDestroy(ev);
    for (cudaEvent_t ev : ctx->item_events) cudaEventDestroy(ev);
    for (cudaEvent_t ev : ctx->chunk_events) cudaEventDestroy(ev);
    for (cudaEvent_t ev : ctx->trace_pool) cudaEventDestroy(ev);
    if (ctx->ev_begin) cudaEventDestroy(ctx->ev_begin);
    if (ctx->ev_end) cudaEventDestroy(ctx->ev_end);
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
    if (ctx->side_stream) cudaStreamDestroy(ctx->side_stream);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

const char *swg_gpu_last_error(const swg_ctx *ctx) { return ctx ? ctx->err : g_last_error; }

int swg_gpu_set_option(swg_ctx *ctx, const char *name, long value)
{
    if (!ctx || !name) return fail(ctx, SWG_ERR_ARG, "NULL argument");
    if (!strcmp(name, "long_threshold")) {
        if (value != 0 && value < 8) return fail(ctx, SWG_ERR_ARG, "long_threshold must be 0 (automatic) or >= 8");
        ctx->long_cols = value;
        if (ctx->db_ready)
            ctx->first_long_tile = value == 0 ? ctx->ntiles
                                              : (uint32_t)(std::upper_bound(ctx->h_tile_cols_mono.begin(), ctx->h_tile_cols_mono.end(),
                                                                            (uint32_t)value) - ctx->h_tile_cols_mono.begin());
    } else if (!strcmp(name, "long_kernel")) {
        if (value != 0 && value != 1) return fail(ctx, SWG_ERR_ARG, "long_kernel must be 0 (32-thread shape) or 1 (cross-warp wavefront)");
        ctx->long_kernel = value;
    } else if (!strcmp(name, "xw_warps")) {
        if (value != 0 && value != 1 && value != 2 && value != 4 && value != 8 && value != 16)
            return fail(ctx, SWG_ERR_ARG, "xw_warps must be 0, 1, 2, 4, 8 or 16");
        ctx->xw_warps = value;
    } else if (!strcmp(name, "trace")) {
        ctx->trace = value;
    } else if (!strcmp(name, "chunk_columns")) {
        if (value < 0) return fail(ctx, SWG_ERR_ARG, "chunk_columns must be 0 (planner), 1 (off) or a column count");
        ctx->chunk_columns = value;
    } else if (!strcmp(name, "xw_rows")) {
        if (value < 0 || value > kMaxRowsPerThread) return fail(ctx, SWG_ERR_ARG, "xw_rows must be 0..32");
        ctx->xw_rows = value;
    } else if (!strcmp(name, "force_group")) {
        if (value != 0 && value != 4 && value != 8 && value != 16 && value != 32)
            return fail(ctx, SWG_ERR_ARG, "force_group must be 0, 4, 8, 16 or 32");
        ctx->force_group = value;
    } else if (!strcmp(name, "force_rows")) {
        if (value < 0 || value > kMaxRowsPerThread) return fail(ctx, SWG_ERR_ARG, "force_rows must be 0..32");
        ctx->force_rows = value;
    } else if (!strcmp(name, "query_pairing")) {
        if (value < 0 || value > 2) return fail(ctx, SWG_ERR_ARG, "query_pairing must be 0 (off), 1 (planner) or 2 (always)");
        ctx->query_pairing = value;
    } else if (!strcmp(name, "q2_group")) {
        if (value != 0 && value != 8 && value != 16 && value != 32) return fail(ctx, SWG_ERR_ARG, "q2_group must be 0, 8, 16 or 32");
        ctx->q2_group = value;
    } else if (!strcmp(name, "q2_rows")) {
        if (value < 0 || value > kMaxRowsPerThread || value % 2 || (value > 0 && value < 8))
            return fail(ctx, SWG_ERR_ARG, "q2_rows must be 0 or an even number in 8..32");
        ctx->q2_rows = value;
    } else if (!strcmp(name, "pass_lines")) {
        if (value != 0 && value != 1) return fail(ctx, SWG_ERR_ARG, "pass_lines must be 0 or 1");
        ctx->pass_lines = value;
    } else if (!strcmp(name, "score_budget_mb")) {
        if (value < 1) return fail(ctx, SWG_ERR_ARG, "score_budget_mb must be >= 1");
        ctx->score_budget = value << 20;
    } else if (!strcmp(name, "verbose")) {
        ctx->verbose = value;
    } else if (!strcmp(name, "grid_blocks")) {
        if (value < 0) return fail(ctx, SWG_ERR_ARG, "grid_blocks must be >= 0");
        ctx->grid_blocks = value;
    } else if (!strcmp(name, "block_threads")) {
        if (value != kBlockThreads) return fail(ctx, SWG_ERR_ARG, "block_threads is fixed at %d in this build", kBlockThreads);
    } else {
        return fail(ctx, SWG_ERR_ARG, "unknown option '%s'", name);
    }
    return SWG_OK;
}

// ---- database ----------------------------------------------------------------------------------
// Upload `local_res` bytes described by `len`/`off` (local sequences, padded to whole tiles) and build the layout.
// src_of_tile[lt] = offset of local tile lt's first residue inside `residues` (NULL: the tiles are back to back).
static int upload_and_build(swg_ctx *ctx, const std::vector<uint16_t> &len, const std::vector<uint64_t> &off,
                            const signed char *residues, const uint64_t *src_of_tile)
{
    const uint64_t ltiles = len.size() / kTileSeqs;
    const uint64_t local_res = off[len.size()];
    DeviceBuf d_res;
    SWG_CUDA(ctx, d_res.reserve(std::max<uint64_t>(local_res, 1)));
    cudaError_t e = cudaSuccess;
    if (!src_of_tile) {
        if (local_res) e = cudaMemcpyAsync(d_res.p, residues, local_res, cudaMemcpyHostToDevice, ctx->stream);
    } else {
        // tile ranges gathered through two pinned staging buffers
        const size_t kStage = 32u << 20;
        char *stage[2] = {nullptr, nullptr};
        cudaEvent_t done[2] = {nullptr, nullptr};
        for (int b = 0; b < 2 && e == cudaSuccess; ++b) {
            e = cudaMallocHost((void **)&stage[b], kStage);
            if (e == cudaSuccess) e = cudaEventCreate(&done[b]);
        }
        uint64_t lt = 0, dst = 0;
        int b = 0;
        while (e == cudaSuccess && lt < ltiles) {
            e = cudaEventSynchronize(done[b]);          // a never-recorded event is complete
            size_t fill = 0;
            while (lt < ltiles) {
                const uint64_t bytes = off[(lt + 1) * kTileSeqs] - off[lt * kTileSeqs];
                if (bytes > kStage) { e = cudaErrorInvalidValue; break; }     // 16 x 65535 < 32 MiB: cannot happen
                if (fill + bytes > kStage) break;
                memcpy(stage[b] + fill, residues + src_of_tile[lt], bytes);
                fill += bytes;
                ++lt;
            }
            if (e == cudaSuccess && fill) {
                e = cudaMemcpyAsync(d_res.as<char>() + dst, stage[b], fill, cudaMemcpyHostToDevice, ctx->stream);
                if (e == cudaSuccess) e = cudaEventRecord(done[b], ctx->stream);
                dst += fill;
            }
            b ^= 1;
        }
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        for (int k = 0; k < 2; ++k) {
            if (stage[k]) cudaFreeHost(stage[k]);
            if (done[k]) cudaEventDestroy(done[k]);
        }
    }
    if (e != cudaSuccess) { d_res.release(); return cuda_fail(ctx, e, "database upload"); }
    const int st = finish_load(ctx, len, off, d_res.as<int8_t>());
    d_res.release();
    return st;
}

static int load_db_impl(swg_ctx *ctx, const uint16_t *lengths, const uint64_t *offsets, const signed char *residues,
                        uint64_t n_sequences, uint64_t n_residues, int shard, int num_shards)
{
    if (!ctx) return fail(nullptr, SWG_ERR_ARG, "ctx is NULL");
    if (num_shards < 1 || shard < 0 || shard >= num_shards) return fail(ctx, SWG_ERR_ARG, "shard %d of %d", shard, num_shards);
    if (n_sequences && (!lengths || (!residues && n_residues))) return fail(ctx, SWG_ERR_ARG, "NULL database arrays");
    if (n_sequences >= (1ull << 32) - kTileSeqs) return fail(ctx, SWG_ERR_ARG, "more than 2^32 sequences");
    SWG_CUDA(ctx, cudaSetDevice(ctx->device));
    free_db(ctx);
    ctx->n_total = n_sequences;
    ctx->shard = (uint32_t)shard;
    ctx->num_shards = (uint32_t)num_shards;

    // local tiles: global tile t (16 consecutive sorted sequences) belongs to shard t % num_shards
    const uint64_t gtiles = (n_sequences + kTileSeqs - 1) / kTileSeqs;
    const uint64_t ltiles = gtiles > (uint64_t)shard ? (gtiles - shard + num_shards - 1) / num_shards : 0;
    std::vector<uint16_t> len(ltiles * kTileSeqs, 0);
    std::vector<uint64_t> off(ltiles * kTileSeqs + 1, 0);
    std::vector<uint64_t> tile_src(ltiles + 1, 0);   // where each local tile's residues start in the caller's array
    if (offsets) {
        // the caller holds the prefix sums: only this shard's tiles are visited
        if (offsets[n_sequences] != n_residues)
            return fail(ctx, SWG_ERR_ARG, "offsets end at %llu residues, caller said %llu", (unsigned long long)offsets[n_sequences],
                        (unsigned long long)n_residues);
        for (uint64_t lt = 0; lt < ltiles; ++lt) {
            const uint64_t g0 = (lt * num_shards + shard) * kTileSeqs;
            tile_src[lt] = offsets[g0];
            for (int k = 0; k < kTileSeqs; ++k)
                if (g0 + k < n_sequences) len[lt * kTileSeqs + k] = lengths[g0 + k];
        }
    } else {
        uint64_t pos = 0, lt = 0;
        for (uint64_t t = 0; t < gtiles; ++t) {
            const bool mine = (t % num_shards) == (uint64_t)shard;
            if (mine) tile_src[lt] = pos;
            for (int k = 0; k < kTileSeqs; ++k) {
                const uint64_t s = t * kTileSeqs + k;
                const uint16_t l = s < n_sequences ? lengths[s] : 0;
                if (mine) len[lt * kTileSeqs + k] = l;
                pos += l;
            }
            if (mine) ++lt;
        }
        if (pos != n_residues) return fail(ctx, SWG_ERR_ARG, "lengths sum to %llu residues, caller said %llu",
                                           (unsigned long long)pos, (unsigned long long)n_residues);
    }
    uint64_t local_res = 0, local_seqs = 0;
    for (size_t i = 0; i < len.size(); ++i) {
        off[i] = local_res;
        local_res += len[i];
    }
    off[len.size()] = local_res;
    for (uint64_t lt = 0; lt < ltiles; ++lt) {
        const uint64_t g0 = (lt * num_shards + shard) * kTileSeqs;
        local_seqs += std::min<uint64_t>(kTileSeqs, n_sequences - g0);
    }
    ctx->local_seqs = local_seqs;
    ctx->local_residues = local_res;
    return upload_and_build(ctx, len, off, residues, num_shards == 1 ? nullptr : tile_src.data());
}

int swg_gpu_load_db(swg_ctx *ctx, const uint16_t *lengths, const signed char *residues, uint64_t n_sequences,
                    uint64_t n_residues, int shard, int num_shards)
{
    return load_db_impl(ctx, lengths, nullptr, residues, n_sequences, n_residues, shard, num_shards);
}

int swg_gpu_load_db_offsets(swg_ctx *ctx, const uint16_t *lengths, const uint64_t *offsets, const signed char *residues,
                            uint64_t n_sequences, uint64_t n_residues, int shard, int num_shards)
{
    if (n_sequences && !offsets) return fail(ctx, SWG_ERR_ARG, "offsets is NULL");
    return load_db_impl(ctx, lengths, offsets, residues, n_sequences, n_residues, shard, num_shards);
}

int swg_gpu_load_db_shard(swg_ctx *ctx, const uint16_t *local_lengths, const signed char *local_residues,
                          uint64_t n_local, uint64_t n_local_residues, int shard, int num_shards, uint64_t n_total)
{
    if (!ctx) return fail(nullptr, SWG_ERR_ARG, "ctx is NULL");
    if (num_shards < 1 || shard < 0 || shard >= num_shards) return fail(ctx, SWG_ERR_ARG, "shard %d of %d", shard, num_shards);
    if (n_local && (!local_lengths || (!local_residues && n_local_residues))) return fail(ctx, SWG_ERR_ARG, "NULL database arrays");
    if (n_total >= (1ull << 32) - kTileSeqs) return fail(ctx, SWG_ERR_ARG, "more than 2^32 sequences");
    // the caller's sequences must be exactly the tiles t = shard, shard + num_shards, ... of the whole database
    const uint64_t gtiles = (n_total + kTileSeqs - 1) / kTileSeqs;
    const uint64_t ltiles = gtiles > (uint64_t)shard ? (gtiles - shard + num_shards - 1) / num_shards : 0;
    uint64_t expect = 0;
    for (uint64_t lt = 0; lt < ltiles; ++lt) {
        const uint64_t g0 = (lt * num_shards + shard) * kTileSeqs;
        expect += std::min<uint64_t>(kTileSeqs, n_total - g0);
    }
    if (expect != n_local)
        return fail(ctx, SWG_ERR_ARG, "shard %d of %d of a %llu-sequence database holds %llu sequences, caller passed %llu",
                    shard, num_shards, (unsigned long long)n_total, (unsigned long long)expect, (unsigned long long)n_local);
    SWG_CUDA(ctx, cudaSetDevice(ctx->device));
    free_db(ctx);
    ctx->n_total = n_total;
    ctx->shard = (uint32_t)shard;
    ctx->num_shards = (uint32_t)num_shards;
    std::vector<uint16_t> len(ltiles * kTileSeqs, 0);
    std::vector<uint64_t> off(ltiles * kTileSeqs + 1, 0);
    uint64_t acc = 0;
    for (uint64_t i = 0; i < len.size(); ++i) {
        off[i] = acc;
        if (i < n_local) { len[i] = local_lengths[i]; acc += local_lengths[i]; }    // only the last tile can be partial
    }
    off[len.size()] = acc;
    if (acc != n_local_residues) return fail(ctx, SWG_ERR_ARG, "lengths sum to %llu residues, caller said %llu",
                                             (unsigned long long)acc, (unsigned long long)n_local_residues);
    ctx->local_seqs = n_local;
    ctx->local_residues = acc;
    return upload_and_build(ctx, len, off, local_residues, nullptr);
}

// Undo the reference's lane interleave (sequences.c:704-723): lane k of group g holds sequence g*vector_length + k;
// its residues are vect_db[disp[g] + j*vector_length + k], padded with code 24.
static void deinterleave(const signed char *vect_db, const uint16_t *vect_lengths, const uint64_t *vect_disp, int vector_length,
                         uint64_t n_sequences, std::vector<uint16_t> &lengths, std::vector<signed char> &flat)
{
    lengths.assign(n_sequences, 0);
    std::vector<uint64_t> start(n_sequences + 1, 0);
    for (uint64_t s = 0; s < n_sequences; ++s) {
        const uint64_t g = s / vector_length, k = s % vector_length;
        const signed char *col = vect_db + vect_disp[g] + k;
        uint32_t l = vect_lengths[g];
        while (l > 0 && col[(uint64_t)(l - 1) * vector_length] >= kPadCode) --l;
        lengths[s] = (uint16_t)l;
        start[s + 1] = start[s] + l;
    }
    flat.assign(std::max<uint64_t>(start[n_sequences], 1), 0);
    for (uint64_t s = 0; s < n_sequences; ++s) {
        const uint64_t g = s / vector_length, k = s % vector_length;
        const signed char *col = vect_db + vect_disp[g] + k;
        signed char *dst = flat.data() + start[s];
        for (uint32_t j = 0; j < lengths[s]; ++j) dst[j] = col[(uint64_t)j * vector_length];
    }
    flat.resize(start[n_sequences]);
}

int swg_gpu_load_db_interleaved(swg_ctx *ctx, const signed char *vect_db, const uint16_t *vect_lengths, uint64_t vect_count,
                                const uint64_t *vect_disp, int vector_length, uint64_t n_sequences, int shard, int num_shards)
{
    if (!ctx) return fail(nullptr, SWG_ERR_ARG, "ctx is NULL");
    if (vector_length != 16 && vector_length != 32) return fail(ctx, SWG_ERR_ARG, "vector_length must be 16 or 32");
    if (n_sequences > vect_count * (uint64_t)vector_length) return fail(ctx, SWG_ERR_ARG, "more sequences than lanes");
    if (n_sequences && (!vect_db || !vect_lengths || !vect_disp)) return fail(ctx, SWG_ERR_ARG, "NULL database arrays");
    std::vector<uint16_t> lengths;
    std::vector<signed char> flat;
    deinterleave(vect_db, vect_lengths, vect_disp, vector_length, n_sequences, lengths, flat);
    return swg_gpu_load_db(ctx, lengths.data(), flat.data(), n_sequences, flat.size(), shard, num_shards);
}

uint64_t swg_gpu_db_local_sequences(const swg_ctx *ctx) { return ctx && ctx->db_ready ? ctx->local_seqs : 0; }
uint64_t swg_gpu_db_local_residues(const swg_ctx *ctx) { return ctx && ctx->db_ready ? ctx->local_residues : 0; }

// ---- queries -----------------------------------------------------------------------------------
int swg_gpu_set_queries(swg_ctx *ctx, const signed char *queries, const uint16_t *q_lengths, const uint32_t *q_disp,
                        uint64_t q_count, const signed char *submat, int open_gap, int extend_gap)
{
    if (!ctx) return fail(nullptr, SWG_ERR_ARG, "ctx is NULL");
    if (q_count && (!queries || !q_lengths || !q_disp)) return fail(ctx, SWG_ERR_ARG, "NULL query arrays");
    if (!submat) return fail(ctx, SWG_ERR_ARG, "NULL substitution matrix");
    if (open_gap < 0 || extend_gap < 0 || open_gap + extend_gap > 4096)
        return fail(ctx, SWG_ERR_ARG, "gap penalties %d/%d out of range", open_gap, extend_gap);
    SWG_CUDA(ctx, cudaSetDevice(ctx->device));
    ctx->q_count = q_count;
    ctx->q_len.assign(q_lengths, q_lengths + q_count);
    ctx->q_off.assign(q_count + 1, 0);
    for (uint64_t i = 0; i < q_count; ++i) ctx->q_off[i + 1] = ctx->q_off[i] + q_lengths[i];
    const uint64_t total = ctx->q_off[q_count];
    SWG_CUDA(ctx, ctx->d_queries.reserve(std::max<uint64_t>(total, 1)));
    SWG_CUDA(ctx, ctx->d_submat.reserve(768));
    // queries are packed back to back (the caller's q_disp may leave gaps) in a pinned staging buffer together with
    // the matrix, and go to the device in ONE copy: a batch of thousands of queries pays the copy latency once
    {
        const size_t need = (size_t)total + 768;
        if (need > ctx->h_stage_cap) {
            if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
            ctx->h_stage = nullptr;
            ctx->h_stage_cap = 0;
            SWG_CUDA(ctx, cudaMallocHost((void **)&ctx->h_stage, need + need / 2));
            ctx->h_stage_cap = need + need / 2;
        }
        SWG_CUDA(ctx, cudaEventSynchronize(ctx->ev_upload));     // the previous upload from this staging buffer is complete
        for (uint64_t i = 0; i < q_count; ++i)
            if (q_lengths[i]) memcpy(ctx->h_stage + ctx->q_off[i], queries + q_disp[i], q_lengths[i]);
        memcpy(ctx->h_stage + total, submat, 768);
        SWG_CUDA(ctx, ctx->d_queries.reserve(need));
        // the upload runs on the copy stream, behind the last run that read this buffer set and ahead of the next one
        // (with swg_gpu_submit's two sets it overlaps the kernels of the batch in flight)
        SWG_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_bufs_free, 0));
        SWG_CUDA(ctx, cudaMemcpyAsync(ctx->d_queries.p, ctx->h_stage, need, cudaMemcpyHostToDevice, ctx->copy_stream));
        SWG_CUDA(ctx, cudaMemcpyAsync(ctx->d_submat.p, ctx->d_queries.as<char>() + total, 768, cudaMemcpyDeviceToDevice,
                                      ctx->copy_stream));
        SWG_CUDA(ctx, cudaEventRecord(ctx->ev_upload, ctx->copy_stream));
        SWG_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_upload, 0));
    }
    ctx->open_gap = open_gap;
    ctx->extend_gap = extend_gap;
    ctx->submat_max = 0;
    for (int i = 0; i < 768; ++i) ctx->submat_max = std::max(ctx->submat_max, (int)submat[i]);
    ctx->queries_ready = true;
    ctx->run_done = false;
    ctx->stats.h2d_bytes = total + 768;
    return SWG_OK;
}

int swg_gpu_run(swg_ctx *ctx, uint64_t top, int keep_scores)
{
    if (!ctx) return fail(nullptr, SWG_ERR_ARG, "ctx is NULL");
    if (!ctx->db_ready) return fail(ctx, SWG_ERR_STATE, "no database loaded");
    if (!ctx->queries_ready) return fail(ctx, SWG_ERR_STATE, "no queries set");
    SWG_CUDA(ctx, cudaSetDevice(ctx->device));
    const uint64_t nq = ctx->q_count;
    const uint64_t n_pad = (uint64_t)ctx->ntiles * kTileSeqs;
    ctx->run_top_stride = top;
    if (top > ctx->n_total) top = ctx->n_total;      // like the reference (swimm.c:51); the rest of a row stays 0
    ctx->run_top = top;
    ctx->stats.launches = 0;
    ctx->stats.cells = 0;
    ctx->stats.padded_cells = 0;
    ctx->stats.rescored = 0;
    ctx->stats.pair_launches = 0;
    ctx->stats.stream_bytes = 0;

    const int grid = ctx->grid_blocks > 0 ? (int)std::min<long>(ctx->grid_blocks, ctx->sm_count) : ctx->sm_count;
    const int warps_per_block = kBlockThreads / 32;
    const size_t warps = (size_t)grid * warps_per_block;
    // the pass lines of multi-launch groups take 8 bytes per database column: only when that fits comfortably (the
    // driver is asked for free memory only when a plan needs lines that are not allocated yet, see below)
    bool lines_fit = ctx->pass_lines != 0;
    // ---- score rows: a window of queries ----
    // All nq rows when the caller wants the full score vectors.  Otherwise at most `score_budget` bytes of rows: the
    // batch is searched in chunks of that many queries, each chunk followed by its own top-r selection, so that a
    // batch of a thousand queries on a multi-million-sequence shard does not need nq x n x 4 bytes.
    uint64_t window = nq;
    if (!keep_scores && nq > 2 && n_pad)
        window = std::min<uint64_t>(nq, std::max<uint64_t>(2, (uint64_t)ctx->score_budget / (n_pad * sizeof(int32_t))));
    const uint64_t nchunks = nq ? (nq + window - 1) / window : 1;
    window = nq ? (nq + nchunks - 1) / nchunks : 0;           // chunks of equal size
    ctx->run_kept_scores = nchunks == 1;
    ctx->run_window = window;

    ShardShape shape;
    shape.residues = ctx->local_residues;
    shape.maxcols = ctx->maxcols;
    shape.ntiles = ctx->ntiles;
    shape.lines_fit = lines_fit;
    PlanOptions opts;
    opts.query_pairing = ctx->query_pairing;
    opts.q2_group = ctx->q2_group;
    opts.q2_rows = ctx->q2_rows;
    opts.force_group = ctx->force_group;
    opts.force_rows = ctx->force_rows;
    std::vector<Config> main_cfgs(nq), wide_cfgs(nq);
    std::vector<WorkItem> &items = ctx->items;
    std::vector<size_t> chunk_end;                      // items [chunk_end[c-1], chunk_end[c]) belong to chunk c
    uint32_t max_passes = 1, q2_launches = 0;
    bool q2_lines = false;
    const uint64_t dummy_lines = warps * 4 * (uint64_t)(kQ2MinSegCols + kQ2LineSlack);     // one per thread group (G >= 8)
    for (int attempt = 0; attempt < 2; ++attempt) {
        items.clear();
        chunk_end.clear();
        for (uint64_t c = 0; c < nchunks; ++c) {
            // every chunk is planned on its own (queries are paired inside a chunk); ids become batch-wide afterwards
            const uint64_t q0 = c * window, q1 = std::min(nq, q0 + window);
            const std::vector<uint16_t> sub_len(ctx->q_len.begin() + q0, ctx->q_len.begin() + q1);
            std::vector<Config> sub_main, sub_wide;
            std::vector<WorkItem> sub_items;
            plan_batch(sub_len, shape, opts, sub_main, sub_wide, sub_items);
            for (uint64_t i = 0; i < q1 - q0; ++i) { main_cfgs[q0 + i] = sub_main[i]; wide_cfgs[q0 + i] = sub_wide[i]; }
            for (WorkItem &it : sub_items) {
                it.qa += (uint32_t)q0;
                for (uint32_t &m : it.members) m += (uint32_t)q0;
                for (Q2Launch &L : it.launches)
                    for (int l = 0; l < 2; ++l)
                        if (L.lane[l].q >= 0) L.lane[l].q += (int32_t)q0;
                items.push_back(std::move(it));
            }
            chunk_end.push_back(items.size());
        }
        max_passes = 1;
        q2_launches = 0;
        q2_lines = false;
        for (uint64_t q = 0; q < nq; ++q) max_passes = std::max(max_passes, std::max(main_cfgs[q].passes, wide_cfgs[q].passes));
        for (const WorkItem &it : items)
            if (it.pair) {
                q2_launches += (uint32_t)it.launches.size();
                if (it.launches.size() > 1) q2_lines = true;
            }
        if (!q2_lines) break;
        const size_t line_bytes = (ctx->line_units + dummy_lines + 2) * sizeof(uint2);
        if (ctx->d_lines.cap < line_bytes && attempt == 0) {
            size_t free_b = 0, total_b = 0;
            if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) free_b = 0;
            if ((double)ctx->line_units * sizeof(uint2) >= 0.6 * (double)free_b) {
                shape.lines_fit = false;           // plan again without multi-launch groups
                continue;
            }
        }
        const cudaError_t le = ctx->d_lines.reserve(line_bytes);
        if (le == cudaSuccess) break;
        if (le != cudaErrorMemoryAllocation || attempt == 1) return cuda_fail(ctx, le, "pass lines");
        cudaGetLastError();              // not enough memory for the pass lines after all: plan again without them
        shape.lines_fit = false;
    }
    if (ctx->verbose) fputs(describe_plan(ctx->q_len, main_cfgs, items).c_str(), stderr);

    // ---- long tiles (K3) ----
    // A warp advances its sequences one column per step, so a sequence with more columns than the warp's share of the
    // shard outlasts the launch it is part of.  Per work item: tiles above 0.8 x (residues / warps / sequences a warp
    // holds) are "long" (option long_threshold > 0: a fixed column count instead).  They leave the item's main launches
    // and are searched, query by query, by the long-sequence kernel (wavefront_xw.cuh) on the high-priority stream with
    // a few SMs of their own, while the main launches fill the other SMs from the side stream -- when the time model
    // says that this beats keeping them (the long-sequence kernel has fewer rows per thread: lower throughput).
    std::vector<uint32_t> item_first_long(items.size(), ctx->ntiles);
    std::vector<int> item_long_grid(items.size(), 0);
    std::vector<XwConfig> xw_cfgs(nq);
    std::vector<uint4> vt_all;                          // column-chunk tables, query after query
    std::vector<std::pair<size_t, size_t>> vt_of(nq, std::make_pair((size_t)0, (size_t)0));   // (offset, count) in vt_all
    bool any_xw = false;
    if (ctx->ntiles) {
        const double res = (double)ctx->local_residues, res9 = res * 1e-9;
        const double total_cols = (double)ctx->h_cols_prefix[ctx->ntiles];
        for (size_t ii = 0; ii < items.size(); ++ii) {
            const WorkItem &it = items[ii];
            const std::vector<uint32_t> qs = it.pair ? it.members : std::vector<uint32_t>(1, it.qa);
            // the item's launches: throughput seconds on the whole shard, step seconds, sequences a warp holds at once
            struct LaunchModel { double thr, step; int passes; double seqs_per_warp; };
            std::vector<LaunchModel> lm;
            bool xw_ok = ctx->long_kernel != 0;
            if (it.pair) {
                for (const Q2Launch &L : it.launches) {
                    const double rate = q2_rate(L.G, L.K, it.launches.size() > 1);
                    lm.push_back({2.0 * L.G * L.K / rate * res9, step_seconds_loaded(L.K, rate), 1, 32.0 / L.G});
                }
            } else {
                const Config &c = main_cfgs[it.qa];
                const double rate = shape_rate(c.G, c.K, c.passes);
                lm.push_back({(double)c.passes * c.G * c.K / rate * res9, step_seconds_loaded(c.K, rate), (int)c.passes, 64.0 / c.G});
            }
            for (uint32_t q : qs) xw_ok = xw_ok && ctx->q_len[q] <= (uint32_t)kXwMaxRows;
            double limit = 1e300;
            for (const LaunchModel &l : lm) limit = std::min(limit, 0.8 * res / ((double)warps * l.seqs_per_warp));
            uint32_t fl = ctx->ntiles;
            const bool forced = ctx->long_cols > 0;
            if (forced) fl = ctx->first_long_tile;
            else if ((double)ctx->maxcols > 1.125 * limit)
                fl = (uint32_t)(std::upper_bound(ctx->h_tile_cols_mono.begin(), ctx->h_tile_cols_mono.end(),
                                                 (uint32_t)std::min(limit, 4.0e9)) - ctx->h_tile_cols_mono.begin());
            if (fl >= ctx->ntiles) continue;
            // outliers only: at most the longest 30 % of the shard's columns leave the main launches (a small shard is
            // "long" as a whole by the column limit; its bulk still belongs on the main kernels)
            if (!forced && total_cols - (double)ctx->h_cols_prefix[fl] > 0.3 * total_cols) {
                const uint64_t keep = (uint64_t)(0.7 * total_cols);
                fl = (uint32_t)(std::lower_bound(ctx->h_cols_prefix.begin(), ctx->h_cols_prefix.end(), keep) - ctx->h_cols_prefix.begin());
                fl = std::min(fl, ctx->ntiles - 1);
            }
            const double long_cols = total_cols - (double)ctx->h_cols_prefix[fl];
            if (!xw_ok) {
                // queries the long-sequence kernel does not cover (or long_kernel = 0): a single query falls back to the
                // 32-thread shape of its own kernel, a query group keeps every tile
                if (!it.pair) item_first_long[ii] = fl;
                continue;
            }
            // SMs for the long tiles: in proportion to their share of the columns, at least kMaxLongBlocks
            const double pairs = (double)(ctx->ntiles - fl) * kTilePairs;
            int lg = (int)std::max<double>((double)kMaxLongBlocks, long_cols / total_cols * grid + 0.999);
            if (fl == 0) lg = grid;
            else lg = std::max(1, std::min(lg, grid - 1));
            // per query: the long-sequence kernel, or -- queries of one pass (<= 1024 rows) -- the 32-thread single-pass
            // shape of the sequence-pair kernel when the same time model favours it (short queries: the cross-warp
            // wavefront has too few rows per thread to pay for its hand-off)
            double t_long = 0.0;
            bool all_ok = true;
            for (uint32_t q : qs) {
                const uint32_t m = std::max<uint32_t>(ctx->q_len[q], 1);
                XwConfig xc = choose_xw_config(m, pairs, long_cols * kTilePairs, (double)ctx->maxcols, lg, ctx->xw_warps, ctx->xw_rows);
                if (m <= (uint32_t)kMaxPassRows && !ctx->xw_warps && !ctx->xw_rows) {
                    const int K = (int)((m + 31) / 32);
                    // (column chunks: swg_plan.cu, plan_column_chunks)
                    size_t n_chunks = 0;
                    uint32_t C = 0;
                    {
                        std::vector<ColumnChunk> chunks;
                        C = plan_column_chunks(m, ctx->submat_max, ctx->extend_gap, ctx->chunk_columns, ctx->h_tile_cols.data(), fl,
                                               ctx->ntiles, ctx->maxcols, chunks);
                        if (!chunks.empty()) {
                            const size_t first = vt_all.size();
                            for (const ColumnChunk &c : chunks) vt_all.push_back(make_uint4(c.tile, c.col0, c.cols, 0));
                            n_chunks = chunks.size();
                            vt_of[q] = std::make_pair(first, n_chunks);
                        }
                    }
                    double chunk_cols = 0.0;
                    for (size_t k = 0; k < n_chunks; ++k) chunk_cols += vt_all[vt_of[q].first + k].z;
                    const double work_cols = n_chunks ? chunk_cols : long_cols;
                    const double longest = n_chunks ? (double)C : (double)ctx->maxcols;
                    const double active = std::min(16.0, std::max(1.0, (n_chunks ? (double)n_chunks : (double)(ctx->ntiles - fl)) * kTilePairs / lg));
                    const double step = std::max(step_seconds_alone(K), step_seconds_loaded(K, shape_rate(32, K, 1)) * active / 16.0);
                    const double t_wide = std::max(work_cols * kTilePairs / (lg * 16.0), longest) * step;
                    // (the long-sequence kernel's hand-off makes its steps slower than this model says: it must win clearly)
                    if (!xc.ok() || n_chunks || t_wide < 1.4 * xc.seconds) {
                        xc = XwConfig();
                        xc.K = K;
                        xc.W = 1;
                        xc.groups = 16;
                        xc.wide = true;
                        xc.seconds = t_wide;
                    } else {
                        vt_of[q] = std::make_pair((size_t)0, (size_t)0);
                    }
                }
                xw_cfgs[q] = xc;
                all_ok = all_ok && xc.ok();
                t_long += xc.seconds;
            }
            if (all_ok && !it.pair && xw_cfgs[it.qa].wide) {
                const Config &c = main_cfgs[it.qa];
                if (c.G == 32 && c.passes == 1 && c.K == xw_cfgs[it.qa].K) all_ok = false;       // the main shape IS that shape
            }
            if (ctx->verbose) {
                // what the time model expects (reported only: the split itself follows the column limit)
                const double last_main = fl ? (double)ctx->h_tile_cols_mono[fl - 1] : 0.0;
                const double main_share = (1.0 - long_cols / total_cols) * grid / std::max(1, grid - lg);
                double t_keep = 0.0, t_split = 0.0;
                for (const LaunchModel &l : lm) {
                    t_keep += std::max(l.thr, (double)ctx->maxcols * l.passes * l.step);
                    t_split += std::max(l.thr * main_share, last_main * l.passes * l.step);
                }
                fprintf(stderr, "[swg] item %zu: tiles %u.. long (limit %.0f columns): model keep %.3f ms, split %.3f ms main + %.3f ms long tiles on %d SMs\n",
                        ii, fl, limit, t_keep * 1e3, t_split * 1e3, t_long * 1e3, lg);
            }
            if (!all_ok) {                     // (a forced shape too small for the query, or nothing to gain)
                for (uint32_t q : qs) { xw_cfgs[q] = XwConfig(); vt_of[q] = std::make_pair((size_t)0, (size_t)0); }
                if (!it.pair && forced) item_first_long[ii] = fl;      // forced threshold: the 32-thread shape
                continue;
            }
            item_first_long[ii] = fl;
            item_long_grid[ii] = lg;
            any_xw = true;
            if (ctx->verbose)
                for (uint32_t q : qs)
                    fprintf(stderr, "[swg] query %u (%u rows): tiles %u..%u (%.0f columns x 16) on %s, %d warps x %d rows, %d pairs per CTA, %d CTAs, %zu column chunks\n",
                            q, (unsigned)ctx->q_len[q], fl, ctx->ntiles - 1, long_cols,
                            xw_cfgs[q].wide ? "the 32-thread shape of the sequence-pair kernel" : "the long-sequence kernel", xw_cfgs[q].W,
                            xw_cfgs[q].K, xw_cfgs[q].groups, lg, vt_of[q].second);
        }
    }

    SWG_CUDA(ctx, ctx->d_scores.reserve(std::max<uint64_t>(window * n_pad, 1) * sizeof(int32_t)));
    SWG_CUDA(ctx, ctx->d_profile.reserve((size_t)max_passes * kPassBytes));
    SWG_CUDA(ctx, ctx->d_profile32.reserve((size_t)max_passes * kPassBytes));
    SWG_CUDA(ctx, ctx->d_boundary.reserve(2 * warps * (size_t)(ctx->maxcols + kBoundarySlack) * sizeof(uint2)));
    SWG_CUDA(ctx, ctx->d_counters.reserve(std::max<uint64_t>(nq, 1) * 4 * sizeof(uint32_t)));
    SWG_CUDA(ctx, ctx->d_xw_counters.reserve(std::max<uint64_t>(nq, 1) * 4 * sizeof(uint32_t)));
    SWG_CUDA(ctx, ctx->d_resc_list.reserve(std::max<uint64_t>(n_pad, 1) * sizeof(uint32_t)));
    if (q2_launches) {
        SWG_CUDA(ctx, ctx->d_profile_q2.reserve((size_t)kQ2ProfileBytes));
        SWG_CUDA(ctx, ctx->d_resc_list2.reserve(std::max<uint64_t>(n_pad, 1) * sizeof(uint32_t)));
        SWG_CUDA(ctx, ctx->d_q2_counters.reserve((size_t)q2_launches * sizeof(uint32_t)));
    }
    const int vset = ctx->vt_set;
    if (!vt_all.empty()) {
        ctx->vt_set ^= 1;
        const size_t bytes = vt_all.size() * sizeof(uint4);
        if (!ctx->ev_vt[vset]) SWG_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_vt[vset], cudaEventDisableTiming));
        SWG_CUDA(ctx, cudaEventSynchronize(ctx->ev_vt[vset]));      // the run before last, which read this set, is over
        if (bytes > ctx->h_vt_cap[vset]) {
            if (ctx->h_vt[vset]) cudaFreeHost(ctx->h_vt[vset]);
            ctx->h_vt[vset] = nullptr;
            ctx->h_vt_cap[vset] = 0;
            SWG_CUDA(ctx, cudaMallocHost((void **)&ctx->h_vt[vset], 2 * bytes));
            ctx->h_vt_cap[vset] = 2 * bytes;
        }
        SWG_CUDA(ctx, ctx->d_vt[vset].reserve(bytes));
        memcpy(ctx->h_vt[vset], vt_all.data(), bytes);
        SWG_CUDA(ctx, cudaMemcpyAsync(ctx->d_vt[vset].p, ctx->h_vt[vset], bytes, cudaMemcpyHostToDevice, ctx->stream));
    }
    if (any_xw) {
        SWG_CUDA(ctx, ctx->d_profile_xw.reserve((size_t)16 * kPassBytes));
        SWG_CUDA(ctx, ctx->d_xw_list.reserve(std::max<uint64_t>(n_pad, 1) * sizeof(uint32_t)));
    }
    const TopkPlan tp = topk_plan(n_pad, top, window);
    SWG_CUDA(ctx, ctx->d_topk_scratch.reserve(tp.scratch_keys * sizeof(uint64_t)));
    SWG_CUDA(ctx, ctx->d_top_out.reserve(std::max<uint64_t>(nq * top, 1) * sizeof(uint64_t)));

    // The CUDA runtime loads a kernel's code the first time it is launched (a millisecond or two per instantiation,
    // and a batch may use a dozen shapes).  Every instantiation this run needs and this context has not launched yet
    // is launched once on an empty task list BEFORE the timed region, so that the reported search time is the search.
    {
        WfParams w;
        memset(&w, 0, sizeof(w));
        w.gap_open_extend = ctx->open_gap + ctx->extend_gap;
        w.gap_extend = ctx->extend_gap;
        w.task_counter = ctx->d_counters.as<uint32_t>();        // incremented by the empty launches, zeroed again below
        w.resc_count = ctx->d_counters.as<uint32_t>() + 1;      // stays 0: "nothing to recompute"
        w.resc_count2 = ctx->d_counters.as<uint32_t>() + 1;
        w.passes = 1;
        cudaError_t we = cudaMemsetAsync(ctx->d_counters.p, 0, 4 * sizeof(uint32_t), ctx->stream);
        // which instantiation family the penalties select: defaults as immediates, gap-extend 1 or 2 as an immediate, generic
        const uint32_t fast = (w.gap_open_extend == kFastGapOpenExtend && w.gap_extend == kFastGapExtend) ? 1u
                              : w.gap_extend == 1 ? 2u : w.gap_extend == 2 ? 3u : 0u;
        auto fresh = [&](uint32_t key) {
            key = key * 4 + fast;
            if (std::find(ctx->warmed.begin(), ctx->warmed.end(), key) != ctx->warmed.end()) return false;
            ctx->warmed.push_back(key);
            return true;
        };
        auto warm_seqpair = [&](bool lane32, const Config &c) {
            w.profile = lane32 ? ctx->d_profile32.as<uint8_t>() : ctx->d_profile.as<uint8_t>();
            w.passes = c.passes;
            const uint32_t key = ((((lane32 ? 1u : 0u) * 64 + (uint32_t)c.G) * 64 + (uint32_t)c.K) * 4 + (c.passes > 1 ? 1u : 0u) +
                                  (c.global_profile ? 2u : 0u));
            if (we == cudaSuccess && fresh(key)) we = launch_wavefront(lane32, c, 1, ctx->stream, w);
        };
        auto warm_xw = [&](const XwConfig &xc) {
            if (!xc.ok()) return;
            if (xc.wide) {
                const Config c = {32, xc.K, 1, false};
                warm_seqpair(false, c);
                warm_seqpair(true, c);
                return;
            }
            w.profile = ctx->d_profile_xw.as<uint8_t>();
            w.xw_warps = (uint32_t)xc.W;
            w.xw_groups = (uint32_t)xc.groups;
            if (we == cudaSuccess && fresh((2u << 20) + (uint32_t)xc.K)) we = launch_xw_l16(xc.K, 1, ctx->stream, w);
            if (we == cudaSuccess && fresh((3u << 20) + (uint32_t)xc.K)) we = launch_xw_l32(xc.K, 1, ctx->stream, w);
        };
        for (const WorkItem &it : items) {
            if (!ctx->ntiles) break;
            if (!it.pair) {
                warm_seqpair(false, main_cfgs[it.qa]);
                warm_seqpair(false, wide_cfgs[it.qa]);
                warm_seqpair(true, wide_cfgs[it.qa]);
                warm_xw(xw_cfgs[it.qa]);
                continue;
            }
            for (uint32_t q : it.members) { warm_seqpair(true, wide_cfgs[q]); warm_xw(xw_cfgs[q]); }
            w.profile = ctx->d_profile_q2.as<uint8_t>();
            for (const Q2Launch &L : it.launches) {
                bool cin = false, cout = false;
                for (int l = 0; l < 2; ++l)
                    if (L.lane[l].q >= 0) { cin |= !L.lane[l].first; cout |= !L.lane[l].last; }
                const uint32_t key = (1u << 20) + (((uint32_t)L.G * 64 + (uint32_t)L.K) * 4 + (cin ? 1u : 0u) + (cout ? 2u : 0u));
                if (we == cudaSuccess && fresh(key)) we = launch_q2(L.G, L.K, cin, cout, 1, ctx->stream, w);
            }
        }
        if (we != cudaSuccess) return cuda_fail(ctx, we, "kernel warm-up");
    }

    SWG_CUDA(ctx, cudaEventRecord(ctx->ev_begin, ctx->stream));
    if (ctx->begin_mark) SWG_CUDA(ctx, cudaEventRecord(ctx->begin_mark, ctx->stream));
    ctx->trace_marks.clear();
    auto mark = [&](const char *label, cudaStream_t st) {
        if (!ctx->trace) return;
        if (ctx->trace_marks.size() >= ctx->trace_pool.size()) {
            cudaEvent_t ev;
            if (cudaEventCreate(&ev) != cudaSuccess) return;
            ctx->trace_pool.push_back(ev);
        }
        cudaEvent_t ev = ctx->trace_pool[ctx->trace_marks.size()];
        cudaEventRecord(ev, st);
        ctx->trace_marks.push_back(std::make_pair(std::string(label), ev));
    };
    SWG_CUDA(ctx, cudaMemsetAsync(ctx->d_counters.p, 0, std::max<uint64_t>(nq, 1) * 4 * sizeof(uint32_t), ctx->stream));
    SWG_CUDA(ctx, cudaMemsetAsync(ctx->d_xw_counters.p, 0, std::max<uint64_t>(nq, 1) * 4 * sizeof(uint32_t), ctx->stream));
    if (q2_launches) SWG_CUDA(ctx, cudaMemsetAsync(ctx->d_q2_counters.p, 0, (size_t)q2_launches * sizeof(uint32_t), ctx->stream));

    WfParams p;
    memset(&p, 0, sizeof(p));
    p.db = ctx->d_db.as<uint4>();
    p.tile_off = ctx->d_tile_off.as<uint64_t>();
    p.tile_cols = ctx->d_tile_cols.as<uint32_t>();
    p.ntiles = ctx->ntiles;
    p.n_total = ctx->n_total;
    p.shard = ctx->shard;
    p.num_shards = ctx->num_shards;
    p.boundary = ctx->d_boundary.as<uint2>();
    p.maxcols = ctx->maxcols + kBoundarySlack;      // stride of a warp's scratch line (multiple of 8)
    p.resc_list = ctx->d_resc_list.as<uint32_t>();
    p.gap_open_extend = ctx->open_gap + ctx->extend_gap;
    p.gap_extend = ctx->extend_gap;

    while (ctx->item_events.size() < items.size() + 1) {
        cudaEvent_t ev;
        SWG_CUDA(ctx, cudaEventCreate(&ev));
        ctx->item_events.push_back(ev);
    }
    SWG_CUDA(ctx, cudaEventRecord(ctx->item_events[0], ctx->stream));
    uint64_t padded = 0;
    uint32_t q2_next_counter = 0;
    uint64_t chunk_q0 = 0;                               // first query of the chunk being enqueued
    auto score_row = [&](uint64_t q) { return ctx->d_scores.as<int32_t>() + (q - chunk_q0) * n_pad; };
    while (ctx->chunk_events.size() < 2 * nchunks) {
        cudaEvent_t ev;
        SWG_CUDA(ctx, cudaEventCreate(&ev));
        ctx->chunk_events.push_back(ev);
    }
    ctx->run_chunks = 0;

    // 32-bit recomputation (K2) of the sequences query q left in `list`
    auto recompute32 = [&](uint64_t q, uint32_t *list, bool build, cudaStream_t st) -> cudaError_t {
        const Config wide_cfg = wide_cfgs[q];
        uint32_t *cnt = ctx->d_counters.as<uint32_t>() + q * 4;
        cudaError_t e = cudaSuccess;
        if (build) {
            e = launch_build_profile(ctx->d_queries.as<int8_t>() + ctx->q_off[q], ctx->q_len[q], ctx->d_submat.as<int8_t>(),
                                     wide_cfg.G, wide_cfg.K, wide_cfg.passes, ctx->d_profile32.as<uint8_t>(), st);
            ctx->stats.launches += 1;
        }
        if (e != cudaSuccess) return e;
        WfParams r = p;
        r.scores = score_row(q);
        r.profile = ctx->d_profile32.as<uint8_t>();
        r.passes = wide_cfg.passes;
        r.tile_first = 0;
        r.tile_count = ctx->ntiles;
        r.task_counter = cnt + 2;
        r.resc_count = cnt + 3;
        r.resc_list = list;
        ctx->stats.launches += 1;
        e = launch_wavefront(true, wide_cfg, grid, st, r);
        mark("32-bit recomputation", st);
        return e;
    };

    // Long tiles [fl, ntiles) of query q on the long-sequence kernel (K3): profile of all W passes, the 16-bit launch
    // (scores merged with atomicMax, hence zeroed first), and the 32-bit recomputation of what it listed, same shape.
    // `phase` 0: only what precedes the search launch (profile build, zeroing); 1: only the launches; 2: both.  The
    // caller forks the side stream BETWEEN the two phases: the long tiles' launch and the main launch then become
    // ready at the same moment and the high-priority stream wins the SMs it needs.  (When the main kernel, whose CTAs
    // are persistent, got all SMs first, the long tiles waited for the whole main kernel: a 0.9 ms search took 1.9.)
    int long_ctas = 0;                                  // CTAs of the last long-tile launch
    auto enqueue_xw = [&](uint64_t q, uint32_t fl, int xgrid, cudaStream_t st, int phase) -> cudaError_t {
        const XwConfig xc = xw_cfgs[q];
        uint32_t *xcnt = ctx->d_xw_counters.as<uint32_t>() + q * 4;
        int32_t *sc = score_row(q);
        const bool zero_first = !xc.wide || vt_of[q].second != 0;       // scores merged with atomicMax
        if (phase != 1) {
            cudaError_t e = launch_build_profile(ctx->d_queries.as<int8_t>() + ctx->q_off[q], ctx->q_len[q], ctx->d_submat.as<int8_t>(),
                                                 32, xc.K, xc.wide ? 1u : (uint32_t)xc.W, ctx->d_profile_xw.as<uint8_t>(), st);
            if (e == cudaSuccess && zero_first)
                e = cudaMemsetAsync(sc + (size_t)fl * kTileSeqs, 0, (size_t)(ctx->ntiles - fl) * kTileSeqs * sizeof(int32_t), st);
            if (e != cudaSuccess || phase == 0) return e;
        }
        if (xc.wide) {
            // one pass of 32 threads x K rows per pair (wavefront.cuh), its own profile, list and counters; then the
            // 32-bit recomputation of what it listed, same shape
            const Config c = {32, xc.K, 1, false};
            cudaError_t e = cudaSuccess;
            WfParams x = p;
            x.scores = sc;
            x.profile = ctx->d_profile_xw.as<uint8_t>();
            x.passes = 1;
            x.tile_first = fl;
            x.tile_count = ctx->ntiles - fl;
            x.task_counter = xcnt + 0;
            x.resc_count = xcnt + 1;
            x.resc_list = ctx->d_xw_list.as<uint32_t>();
            uint64_t tasks = (uint64_t)x.tile_count * kTilePairs;
            if (vt_of[q].second) {
                // column chunks instead of whole tiles (scores merged with atomicMax: zeroed in phase 0)
                x.vt = ctx->d_vt[vset].as<uint4>() + vt_of[q].first;
                x.vt_count = (uint32_t)vt_of[q].second;
                tasks = (uint64_t)x.vt_count * kTilePairs;
            }
            xgrid = (int)std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)xgrid, (tasks + 15) / 16));
            long_ctas = xgrid;
            e = launch_wavefront(false, c, xgrid, st, x);
            x.vt = nullptr;
            x.vt_count = 0;
            x.task_counter = xcnt + 2;
            x.tile_first = 0;
            x.tile_count = ctx->ntiles;
            mark("long tiles: 32-thread shape", st);
            if (e == cudaSuccess) e = launch_wavefront(true, c, xgrid, st, x);
            mark("long tiles: 32-bit recomputation", st);
            ctx->stats.launches += 3;
            for (uint32_t t = fl; t < ctx->ntiles; ++t) padded += (uint64_t)32 * xc.K * (ctx->h_tile_cols[t] + 31) * kTileSeqs;
            return e;
        }
        cudaError_t e = cudaSuccess;
        WfParams x = p;
        x.scores = sc;
        x.profile = ctx->d_profile_xw.as<uint8_t>();
        x.passes = (uint32_t)xc.W;
        x.tile_first = fl;
        x.tile_count = ctx->ntiles - fl;
        x.task_counter = xcnt + 0;
        x.resc_count = xcnt + 1;
        x.resc_list = ctx->d_xw_list.as<uint32_t>();
        x.xw_warps = (uint32_t)xc.W;
        x.xw_groups = (uint32_t)xc.groups;
        // no more CTAs than there are pairs to keep their groups busy
        xgrid = (int)std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)xgrid, ((uint64_t)x.tile_count * kTilePairs + xc.groups - 1) / xc.groups));
        long_ctas = xgrid;
        e = launch_xw_l16(xc.K, xgrid, st, x);
        mark("long tiles: long-sequence kernel", st);
        x.task_counter = xcnt + 2;
        if (e == cudaSuccess) e = launch_xw_l32(xc.K, xgrid, st, x);
        mark("long tiles: 32-bit recomputation", st);
        ctx->stats.launches += 3;
        ctx->stats.cells += 0;
        for (uint32_t t = fl; t < ctx->ntiles; ++t)
            padded += (uint64_t)xc.W * 32 * xc.K * (ctx->h_tile_cols[t] + 40) * kTileSeqs;
        return e;
    };
    // CTAs for the long tiles of one query: in proportion to their share of the columns (the long and the main launches
    // run side by side, one CTA per SM), at least kMaxLongBlocks when there is that much work
    auto long_grid_for = [&](uint32_t fl, uint64_t groups_needed, bool main_runs) -> int {
        const double long_cols_sum = (double)(ctx->h_cols_prefix[ctx->ntiles] - ctx->h_cols_prefix[fl]);
        const double share = long_cols_sum / std::max(1.0, (double)ctx->h_cols_prefix[ctx->ntiles]);
        const uint64_t by_share = (uint64_t)(share * grid + 0.999);
        int lg = (int)std::min<uint64_t>(std::max<uint64_t>(groups_needed, 1), std::max<uint64_t>(kMaxLongBlocks, by_share));
        if (lg > grid - 1 && main_runs) lg = std::max(1, grid - 1);
        if (!main_runs) lg = (int)std::min<uint64_t>(std::max<uint64_t>(groups_needed, 1), (uint64_t)grid);
        return lg;
    };

    size_t chunk_of_item = 0;
    for (size_t ii = 0; ii < items.size() && ctx->ntiles; ++ii) {
        while (ii >= chunk_end[chunk_of_item]) ++chunk_of_item;
        chunk_q0 = chunk_of_item * window;
        const bool chunk_ends_here = ii + 1 == chunk_end[chunk_of_item];
        // the top-r selection of a finished chunk (its score rows are reused by the next chunk)
        auto finish_chunk = [&]() -> int {
            const uint64_t q1 = std::min(nq, chunk_q0 + window);
            SWG_CUDA(ctx, cudaEventRecord(ctx->chunk_events[2 * chunk_of_item], ctx->stream));
            if (top) {
                cudaError_t te = launch_topk(tp, ctx->d_scores.as<int32_t>(), q1 - chunk_q0, ctx->n_total, ctx->shard,
                                             ctx->num_shards, ctx->d_topk_scratch.as<uint64_t>(),
                                             ctx->d_top_out.as<uint64_t>() + chunk_q0 * top, ctx->stream, &ctx->stats.launches);
                if (te != cudaSuccess) return cuda_fail(ctx, te, "top-r launch");
            }
            SWG_CUDA(ctx, cudaEventRecord(ctx->chunk_events[2 * chunk_of_item + 1], ctx->stream));
            ctx->run_chunks = chunk_of_item + 1;
            return SWG_OK;
        };
        const WorkItem &it = items[ii];
        if (it.pair) {
            // ---- two queries per register: one launch per pass over the shard's tiles below the long-tile threshold ----
            cudaError_t e = cudaSuccess;
            const uint32_t main_tiles = item_first_long[ii];
            cudaStream_t ks = ctx->stream;               // stream of the pair kernel's launches
            bool forked = false;
            if (main_tiles < ctx->ntiles) {
                e = enqueue_xw(it.members[0], main_tiles, item_long_grid[ii], ctx->stream, 0);
                if (e == cudaSuccess && main_tiles) {
                    e = cudaEventRecord(ctx->ev_fork, ctx->stream);
                    if (e == cudaSuccess) e = cudaStreamWaitEvent(ctx->side_stream, ctx->ev_fork, 0);
                    forked = e == cudaSuccess;
                    if (forked) ks = ctx->side_stream;
                }
                if (e == cudaSuccess) e = enqueue_xw(it.members[0], main_tiles, item_long_grid[ii], ctx->stream, 1);
                for (size_t k = 1; k < it.members.size() && e == cudaSuccess; ++k)
                    e = enqueue_xw(it.members[k], main_tiles, item_long_grid[ii], ctx->stream, 2);
                if (e != cudaSuccess) return cuda_fail(ctx, e, "long-sequence kernel launch");
            }
            auto lane_q = [&](const LaneSlice &sl) { return sl.q >= 0 ? ctx->d_queries.as<int8_t>() + ctx->q_off[sl.q] : nullptr; };
            auto lane_m = [&](const LaneSlice &sl) { return sl.q >= 0 ? (uint32_t)ctx->q_len[sl.q] : 0u; };
            WfParams pq = p;
            pq.resc_list = ctx->d_resc_list.as<uint32_t>();
            pq.resc_list2 = ctx->d_resc_list2.as<uint32_t>();
            pq.boundary = ctx->d_lines.as<uint2>();
            pq.line_off = ctx->d_line_off.as<uint64_t>();
            pq.line_dummy = ctx->line_units;
            pq.tile_first = 0;
            pq.tile_count = main_tiles;
            pq.passes = 1;
            for (size_t li = 0; li < it.launches.size() && e == cudaSuccess && main_tiles; ++li) {
                const Q2Launch &L = it.launches[li];
                bool cin = false, cout = false;
                uint32_t cin_mask = 0;
                pq.lane_flags = 0;
                for (int l = 0; l < 2; ++l) {
                    const LaneSlice &sl = L.lane[l];
                    int32_t *sc = nullptr;
                    uint32_t *rc = nullptr;
                    if (sl.q >= 0) {
                        sc = score_row((uint64_t)sl.q);
                        rc = ctx->d_counters.as<uint32_t>() + (uint64_t)sl.q * 4 + 3;
                        pq.lane_flags |= (kLaneActive | (sl.first ? kLaneFirst : 0u)) << (2 * l);
                        if (!sl.first) { cin = true; cin_mask |= 0xffffu << (16 * l); }
                        if (!sl.last) cout = true;
                    }
                    if (l == 0) { pq.scores = sc; pq.resc_count = rc; }
                    else { pq.scores2 = sc; pq.resc_count2 = rc; }
                }
                if (cin && cin_mask != 0xffffffffu) {
                    // one lane continues, the other starts a query (or idles): its half of the lines must read as zero
                    e = launch_clear_lane(ctx->d_lines.as<uint2>(), ctx->line_units, cin_mask, ks);
                    ctx->stats.launches += 1;
                    ctx->stats.stream_bytes += 2 * ctx->line_units * sizeof(uint2);
                    if (e != cudaSuccess) break;
                }
                // the launch's profile (one slot: the stream orders the build behind the previous launch, 100 KB, ~5 us)
                e = launch_build_profile_q2(lane_q(L.lane[0]), lane_m(L.lane[0]), lane_q(L.lane[1]), lane_m(L.lane[1]),
                                            ctx->d_submat.as<int8_t>(), L.G, L.K, L.lane[0].row0, L.lane[1].row0,
                                            ctx->d_profile_q2.as<uint8_t>(), ks);
                ctx->stats.launches += 1;
                if (e != cudaSuccess) break;
                pq.profile = ctx->d_profile_q2.as<uint8_t>();
                pq.task_counter = ctx->d_q2_counters.as<uint32_t>() + q2_next_counter++;
                mark("query-pair profile", ks);
                e = launch_q2(L.G, L.K, cin, cout, grid, ks, pq);
                mark("query-pair kernel", ks);
                ctx->stats.launches += 1;
                ctx->stats.pair_launches += 1;
                ctx->stats.stream_bytes += ctx->db_units * sizeof(uint4) + ((int)cin + (int)cout) * ctx->line_units * sizeof(uint2);
                padded += 2ull * L.G * L.K * (uint64_t)((ctx->avg_cols + L.G - 1) * main_tiles) * kTileSeqs;
                // a query that ends here: its overflowed sequences are recomputed before the lane's list is reused
                for (int l = 0; l < 2 && e == cudaSuccess; ++l)
                    if (L.lane[l].q >= 0 && L.lane[l].last)
                        e = recompute32((uint64_t)L.lane[l].q, l == 0 ? ctx->d_resc_list.as<uint32_t>()
                                                                       : ctx->d_resc_list2.as<uint32_t>(), true, ks);
            }
            if (e == cudaSuccess && forked) {
                e = cudaEventRecord(ctx->ev_join, ctx->side_stream);
                if (e == cudaSuccess) e = cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0);
            }
            if (e != cudaSuccess) return cuda_fail(ctx, e, "query-pair kernel launch");
            SWG_CUDA(ctx, cudaEventRecord(ctx->item_events[ii + 1], ctx->stream));
            ctx->stats.cells += it.member_rows * ctx->local_residues;
            if (chunk_ends_here) { const int fs = finish_chunk(); if (fs != SWG_OK) return fs; }
            continue;
        }
        const uint64_t q = it.qa;
        const uint32_t m = ctx->q_len[q];
        const Config main_cfg = main_cfgs[q];
        const Config wide_cfg = wide_cfgs[q];
        uint32_t *cnt = ctx->d_counters.as<uint32_t>() + q * 4;
        const int8_t *d_q = ctx->d_queries.as<int8_t>() + ctx->q_off[q];
        p.scores = score_row(q);
        p.resc_count = cnt + 3;

        cudaError_t e = launch_build_profile(d_q, m, ctx->d_submat.as<int8_t>(), wide_cfg.G, wide_cfg.K, wide_cfg.passes,
                                             ctx->d_profile32.as<uint8_t>(), ctx->stream);
        ctx->stats.launches += 1;
        const bool same = main_cfg.G == wide_cfg.G && main_cfg.K == wide_cfg.K && main_cfg.passes == wide_cfg.passes &&
                          main_cfg.global_profile == wide_cfg.global_profile;
        if (e == cudaSuccess && !same) {
            e = launch_build_profile(d_q, m, ctx->d_submat.as<int8_t>(), main_cfg.G, main_cfg.K, main_cfg.passes,
                                     ctx->d_profile.as<uint8_t>(), ctx->stream);
            ctx->stats.launches += 1;
        }
        // Long tiles start first, on a few SMs of their own (high-priority stream), while the main kernel (launched on
        // the low-priority side stream right behind) fills the others.  They run the long-sequence kernel; queries it
        // does not cover (more than 8192 rows) and option long_kernel = 0 fall back to the 32-thread shape of this kernel.
        uint32_t first_long = item_first_long[ii];
        const bool use_xw = first_long < ctx->ntiles && xw_cfgs[q].ok();
        if (!use_xw && same) first_long = ctx->ntiles;          // the main shape IS the 32-thread shape
        uint32_t main_tiles = ctx->ntiles;
        bool forked = false;
        if (e == cudaSuccess && first_long < ctx->ntiles) {
            main_tiles = first_long;
            const uint32_t long_tiles = ctx->ntiles - first_long;
            if (use_xw && e == cudaSuccess) e = enqueue_xw(q, first_long, item_long_grid[ii], ctx->stream, 0);
            if (main_tiles && e == cudaSuccess) {
                e = cudaEventRecord(ctx->ev_fork, ctx->stream);
                if (e == cudaSuccess) e = cudaStreamWaitEvent(ctx->side_stream, ctx->ev_fork, 0);
                forked = e == cudaSuccess;
            }
            if (use_xw) {
                if (e == cudaSuccess) e = enqueue_xw(q, first_long, item_long_grid[ii], ctx->stream, 1);
            } else {
                const uint64_t long_warps = (uint64_t)long_tiles * kTilePairs;          // one pair per warp at G = 32
                const int long_grid = long_grid_for(first_long, (long_warps + warps_per_block - 1) / warps_per_block, main_tiles > 0);
                p.profile = ctx->d_profile32.as<uint8_t>();
                p.passes = wide_cfg.passes;
                p.tile_first = first_long;
                p.tile_count = long_tiles;
                p.task_counter = cnt + 0;
                p.boundary = ctx->d_boundary.as<uint2>() + warps * (size_t)p.maxcols;      // its own scratch lines
                if (e == cudaSuccess) e = launch_wavefront(false, wide_cfg, long_grid, ctx->stream, p);
                p.boundary = ctx->d_boundary.as<uint2>();
                ctx->stats.launches += 1;
                for (uint32_t t = p.tile_first; t < ctx->ntiles; ++t)
                    padded += (uint64_t)wide_cfg.passes * wide_cfg.G * wide_cfg.K * (ctx->h_tile_cols[t] + wide_cfg.G - 1) * kTileSeqs;
            }
        }
        if (e == cudaSuccess && main_tiles) {
            p.profile = same ? ctx->d_profile32.as<uint8_t>() : ctx->d_profile.as<uint8_t>();
            p.passes = main_cfg.passes;
            p.tile_first = 0;
            p.tile_count = main_tiles;
            p.task_counter = cnt + 1;
            mark("before the sequence-pair kernel", forked ? ctx->side_stream : ctx->stream);
            // With long tiles running beside it the main kernel leaves them their SMs (its CTAs are persistent: had it
            // taken every SM first, the long tiles would wait for all of it), and a second wave of the same kernel --
            // same task counter -- follows the long tiles on their stream and takes over their SMs when they are done.
            const int wave2 = (forked && use_xw && long_ctas > 0 && long_ctas < grid) ? long_ctas : 0;
            e = launch_wavefront(false, main_cfg, grid - wave2, forked ? ctx->side_stream : ctx->stream, p);
            mark("sequence-pair kernel", forked ? ctx->side_stream : ctx->stream);
            if (e == cudaSuccess && forked) e = cudaEventRecord(ctx->ev_join, ctx->side_stream);
            ctx->stats.launches += 1;
            if (e == cudaSuccess && wave2) {
                WfParams p2 = p;
                p2.boundary = p.boundary + (size_t)(grid - wave2) * warps_per_block * (size_t)p.maxcols;     // its own scratch lines
                e = launch_wavefront(false, main_cfg, wave2, ctx->stream, p2);
                mark("sequence-pair kernel, second wave", ctx->stream);
                ctx->stats.launches += 1;
            }
        }
        if (e == cudaSuccess && forked) e = cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0);
        mark("joined", ctx->stream);
        if (e == cudaSuccess) e = recompute32(q, ctx->d_resc_list.as<uint32_t>(), false, ctx->stream);
        if (e != cudaSuccess) return cuda_fail(ctx, e, "search kernel launch");
        SWG_CUDA(ctx, cudaEventRecord(ctx->item_events[ii + 1], ctx->stream));
        ctx->stats.cells += (uint64_t)m * ctx->local_residues;
        ctx->stats.stream_bytes += ctx->db_units * sizeof(uint4);
        padded += (uint64_t)main_cfg.passes * main_cfg.G * main_cfg.K *
                  (uint64_t)((ctx->avg_cols + main_cfg.G - 1) * main_tiles) * kTileSeqs;
        if (chunk_ends_here) { const int fs = finish_chunk(); if (fs != SWG_OK) return fs; }
    }
    ctx->stats.padded_cells = padded;
    if (!ctx->ntiles && nq * top) SWG_CUDA(ctx, cudaMemsetAsync(ctx->d_top_out.p, 0, nq * top * sizeof(uint64_t), ctx->stream));
    SWG_CUDA(ctx, cudaEventRecord(ctx->ev_end, ctx->stream));
    SWG_CUDA(ctx, cudaEventRecord(ctx->ev_bufs_free, ctx->stream));
    if (!vt_all.empty()) SWG_CUDA(ctx, cudaEventRecord(ctx->ev_vt[vset], ctx->stream));
    ctx->run_done = true;
    return SWG_OK;
}

int swg_gpu_sync(swg_ctx *ctx)
{
    if (!ctx) return fail(nullptr, SWG_ERR_ARG, "ctx is NULL");
    SWG_CUDA(ctx, cudaSetDevice(ctx->device));
    SWG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (ctx->run_done) {
        float ms_all = 0.f, ms_topr = 0.f;
        SWG_CUDA(ctx, cudaEventElapsedTime(&ms_all, ctx->ev_begin, ctx->ev_end));
        for (uint64_t c = 0; c < ctx->run_chunks; ++c) {
            float ms = 0.f;
            SWG_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->chunk_events[2 * c], ctx->chunk_events[2 * c + 1]));
            ms_topr += ms;
        }
        if (ctx->trace)
            for (const auto &mk : ctx->trace_marks) {
                float ms = 0.f;
                if (cudaEventElapsedTime(&ms, ctx->ev_begin, mk.second) == cudaSuccess)
                    fprintf(stderr, "[swg trace] %9.3f ms  %s\n", ms, mk.first.c_str());
            }
        ctx->stats.device_seconds = ms_all * 1e-3;
        ctx->stats.search_seconds = (ms_all - ms_topr) * 1e-3;
        ctx->stats.topr_seconds = ms_topr * 1e-3;
        ctx->q_seconds.assign(ctx->q_count, 0.0);
        for (size_t ii = 0; ii < ctx->items.size() && ctx->ntiles; ++ii) {
            float ms = 0.f;
            SWG_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->item_events[ii], ctx->item_events[ii + 1]));
            const WorkItem &it = ctx->items[ii];
            if (!it.pair) ctx->q_seconds[it.qa] = ms * 1e-3;
            else      // a group's time is split in proportion to the cells of its queries
                for (uint32_t q : it.members)
                    ctx->q_seconds[q] = ms * 1e-3 * (double)ctx->q_len[q] / (double)std::max<uint64_t>(it.member_rows, 1);
        }
    }
    return SWG_OK;
}

int swg_gpu_fetch(swg_ctx *ctx, int32_t *scores, uint64_t *top_keys)
{
    if (!ctx) return fail(nullptr, SWG_ERR_ARG, "ctx is NULL");
    if (!ctx->run_done) return fail(ctx, SWG_ERR_STATE, "fetch without a run");
    if (scores && !ctx->run_kept_scores)
        return fail(ctx, SWG_ERR_STATE, "the run did not keep every query's score row (swg_gpu_run with keep_scores = 0)");
    SWG_CUDA(ctx, cudaSetDevice(ctx->device));
    const uint64_t nq = ctx->q_count, n_pad = (uint64_t)ctx->ntiles * kTileSeqs;
    uint64_t d2h = 0;
    if (top_keys && nq * ctx->run_top_stride) {
        // rows of the caller's array are `top` keys apart; the device holds min(top, n) keys per query
        if (ctx->run_top != ctx->run_top_stride) memset(top_keys, 0, nq * ctx->run_top_stride * sizeof(uint64_t));
        if (ctx->run_top)
            SWG_CUDA(ctx, cudaMemcpy2DAsync(top_keys, ctx->run_top_stride * sizeof(uint64_t), ctx->d_top_out.p,
                                            ctx->run_top * sizeof(uint64_t), ctx->run_top * sizeof(uint64_t), nq,
                                            cudaMemcpyDeviceToHost, ctx->stream));
        d2h += nq * ctx->run_top * sizeof(uint64_t);
    }
    std::vector<int32_t> local;
    if (scores && nq * n_pad) {
        // local (tile-sharded) order -> positions in the whole length-sorted database.  One shard: a query's local
        // row IS the global row (2-D copy, the pad lanes of the last tile are dropped); several shards: local tile lt is
        // global tile lt * num_shards + shard, so the rows are scattered tile by tile (64 bytes) from a staging copy.
        if (ctx->num_shards == 1) {
            SWG_CUDA(ctx, cudaMemcpy2DAsync(scores, ctx->n_total * sizeof(int32_t), ctx->d_scores.p, n_pad * sizeof(int32_t),
                                            ctx->n_total * sizeof(int32_t), nq, cudaMemcpyDeviceToHost, ctx->stream));
        } else {
            local.resize(nq * n_pad);
            SWG_CUDA(ctx, cudaMemcpyAsync(local.data(), ctx->d_scores.p, nq * n_pad * sizeof(int32_t), cudaMemcpyDeviceToHost,
                                          ctx->stream));
        }
        d2h += nq * ctx->local_seqs * sizeof(int32_t);
    }
    uint32_t resc = 0;
    std::vector<uint32_t> cnt(std::max<uint64_t>(nq, 1) * 4, 0), xcnt(std::max<uint64_t>(nq, 1) * 4, 0);
    if (nq) {
        SWG_CUDA(ctx, cudaMemcpyAsync(cnt.data(), ctx->d_counters.p, nq * 4 * sizeof(uint32_t), cudaMemcpyDeviceToHost,
                                      ctx->stream));
        SWG_CUDA(ctx, cudaMemcpyAsync(xcnt.data(), ctx->d_xw_counters.p, nq * 4 * sizeof(uint32_t), cudaMemcpyDeviceToHost,
                                      ctx->stream));
    }
    const int st = swg_gpu_sync(ctx);
    if (st != SWG_OK) return st;
    for (uint64_t q = 0; q < nq; ++q) resc += cnt[q * 4 + 3] + xcnt[q * 4 + 1];
    ctx->stats.rescored = resc;
    ctx->stats.d2h_bytes = d2h;
    if (!local.empty()) {
        const uint64_t ltiles = ctx->ntiles;
        for (uint64_t q = 0; q < nq; ++q) {
            int32_t *row = scores + q * ctx->n_total;
            const int32_t *src = local.data() + q * n_pad;
            for (uint64_t lt = 0; lt < ltiles; ++lt) {
                const uint64_t g0 = (lt * ctx->num_shards + ctx->shard) * kTileSeqs;
                if (g0 >= ctx->n_total) break;
                memcpy(row + g0, src + lt * kTileSeqs, std::min<uint64_t>(kTileSeqs, ctx->n_total - g0) * sizeof(int32_t));
            }
        }
    }
    return SWG_OK;
}

int swg_gpu_search(swg_ctx *ctx, const signed char *queries, const uint16_t *q_lengths, const uint32_t *q_disp,
                   uint64_t q_count, const signed char *submat, int open_gap, int extend_gap, uint64_t top, int32_t *scores,
                   uint64_t *top_keys, double *work_seconds)
{
    int st = swg_gpu_set_queries(ctx, queries, q_lengths, q_disp, q_count, submat, open_gap, extend_gap);
    if (st == SWG_OK) st = swg_gpu_run(ctx, top_keys ? top : 0, scores != nullptr);
    if (st == SWG_OK) st = swg_gpu_fetch(ctx, scores, top_keys);
    if (st == SWG_OK && work_seconds) *work_seconds = ctx->stats.search_seconds;
    return st;
}

// ---- opt-in: where the alignments behind the hits start and end --------------------------------------------
int swg_gpu_align_ends(swg_ctx *ctx, int32_t *coords)
{
    if (!ctx || !coords) return fail(ctx, SWG_ERR_ARG, "NULL argument");
    if (!ctx->run_done || !ctx->queries_ready) return fail(ctx, SWG_ERR_STATE, "swg_gpu_align_ends needs a completed swg_gpu_run of the current queries");
    SWG_CUDA(ctx, cudaSetDevice(ctx->device));
    const uint64_t nq = ctx->q_count, top = ctx->run_top, stride = ctx->run_top_stride;
    for (uint64_t i = 0; i < nq * stride * 4; ++i) coords[i] = -1;
    if (!nq || !top || !ctx->ntiles) return SWG_OK;
    const uint64_t line_stride = ctx->maxcols;
    const uint64_t line_bytes = nq * top * line_stride * sizeof(int2);
    if (line_bytes > (8ull << 30)) return fail(ctx, SWG_ERR_ARG, "%llu hits x %llu columns: too many for the coordinate pass",
                                                (unsigned long long)(nq * top), (unsigned long long)line_stride);
    SWG_CUDA(ctx, ctx->d_q_off.reserve((nq + 1) * sizeof(uint32_t)));
    SWG_CUDA(ctx, ctx->d_align_lines.reserve(line_bytes));
    SWG_CUDA(ctx, ctx->d_coords.reserve(nq * top * 4 * sizeof(int32_t)));
    SWG_CUDA(ctx, cudaMemcpyAsync(ctx->d_q_off.p, ctx->q_off.data(), (nq + 1) * sizeof(uint32_t), cudaMemcpyHostToDevice, ctx->stream));
    cudaError_t e = launch_align_ends(ctx->d_top_out.as<uint64_t>(), nq, top, ctx->d_queries.as<int8_t>(), ctx->d_q_off.as<uint32_t>(),
                                      ctx->d_submat.as<int8_t>(), ctx->open_gap + ctx->extend_gap, ctx->extend_gap,
                                      ctx->d_db.as<uint4>(), ctx->d_tile_off.as<uint64_t>(), ctx->d_tile_cols.as<uint32_t>(),
                                      ctx->num_shards, ctx->d_align_lines.as<int2>(), (uint32_t)line_stride,
                                      ctx->d_coords.as<int32_t>(), ctx->stream);
    if (e != cudaSuccess) return cuda_fail(ctx, e, "coordinate pass launch");
    SWG_CUDA(ctx, cudaMemcpy2DAsync(coords, stride * 4 * sizeof(int32_t), ctx->d_coords.p, top * 4 * sizeof(int32_t),
                                    top * 4 * sizeof(int32_t), nq, cudaMemcpyDeviceToHost, ctx->stream));
    SWG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return SWG_OK;
}

// ---- streaming: two batches in flight ----------------------------------------------------------------
static void swap_slot(swg_ctx *ctx, swg_ctx::QuerySlot &sl)
{
    std::swap(ctx->d_queries, sl.d_queries);
    std::swap(ctx->d_submat, sl.d_submat);
    std::swap(ctx->d_top_out, sl.d_top_out);
    std::swap(ctx->h_stage, sl.h_stage);
    std::swap(ctx->h_stage_cap, sl.h_stage_cap);
    std::swap(ctx->ev_upload, sl.ev_upload);
    std::swap(ctx->ev_bufs_free, sl.ev_bufs_free);
}

int swg_gpu_submit(swg_ctx *ctx, const signed char *queries, const uint16_t *q_lengths, const uint32_t *q_disp,
                   uint64_t q_count, const signed char *submat, int open_gap, int extend_gap, uint64_t top, int *ticket)
{
    if (!ctx || !ticket) return fail(ctx, SWG_ERR_ARG, "NULL argument");
    swg_ctx::QuerySlot &sl = ctx->slots[ctx->next_ticket & 1];
    if (sl.busy) return fail(ctx, SWG_ERR_STATE, "two batches are in flight: poll ticket %d first", sl.ticket);
    SWG_CUDA(ctx, cudaSetDevice(ctx->device));
    swap_slot(ctx, sl);                  // the batch uses the slot's buffer set
    cudaError_t e = cudaSuccess;
    int st = swg_gpu_set_queries(ctx, queries, q_lengths, q_disp, q_count, submat, open_gap, extend_gap);
    // the batch's clock starts where a run's does: behind the one-off loading of kernel code a first call may need
    ctx->begin_mark = sl.ev_begin;
    if (st == SWG_OK) st = swg_gpu_run(ctx, top, 0);
    ctx->begin_mark = nullptr;
    const uint64_t n_keys = q_count * ctx->run_top;
    if (st == SWG_OK && e == cudaSuccess && n_keys > sl.h_keys_cap) {
        if (sl.h_keys) cudaFreeHost(sl.h_keys);
        sl.h_keys = nullptr;
        sl.h_keys_cap = 0;
        e = cudaMallocHost((void **)&sl.h_keys, (n_keys + n_keys / 2) * sizeof(uint64_t));
        if (e == cudaSuccess) sl.h_keys_cap = n_keys + n_keys / 2;
    }
    if (st == SWG_OK && e == cudaSuccess && n_keys)
        e = cudaMemcpyAsync(sl.h_keys, ctx->d_top_out.p, n_keys * sizeof(uint64_t), cudaMemcpyDeviceToHost, ctx->stream);
    if (st == SWG_OK && e == cudaSuccess) e = cudaEventRecord(sl.ev_done, ctx->stream);
    if (st == SWG_OK && e == cudaSuccess) e = cudaEventRecord(ctx->ev_bufs_free, ctx->stream);     // behind the download
    swap_slot(ctx, sl);
    ctx->run_done = false;               // the legacy fetch must not read another batch's buffers
    ctx->queries_ready = false;
    if (st != SWG_OK) return st;
    if (e != cudaSuccess) return cuda_fail(ctx, e, "swg_gpu_submit");
    sl.busy = true;
    sl.ticket = ctx->next_ticket++;
    sl.q_count = q_count;
    sl.top = ctx->run_top;
    sl.top_stride = ctx->run_top_stride;
    ctx->stats.h2d_bytes = ctx->q_off[q_count] + 768;
    ctx->stats.d2h_bytes = n_keys * sizeof(uint64_t);
    *ticket = sl.ticket;
    return SWG_OK;
}

int swg_gpu_poll(swg_ctx *ctx, int ticket, int wait, uint64_t *top_keys, double *device_seconds, int *done)
{
    if (!ctx || !done) return fail(ctx, SWG_ERR_ARG, "NULL argument");
    *done = 0;
    swg_ctx::QuerySlot &sl = ctx->slots[ticket & 1];
    if (ticket < 0 || !sl.busy || sl.ticket != ticket) return fail(ctx, SWG_ERR_STATE, "ticket %d is not in flight", ticket);
    SWG_CUDA(ctx, cudaSetDevice(ctx->device));
    if (wait) SWG_CUDA(ctx, cudaEventSynchronize(sl.ev_done));
    else {
        const cudaError_t q = cudaEventQuery(sl.ev_done);
        if (q == cudaErrorNotReady) return SWG_OK;
        if (q != cudaSuccess) return cuda_fail(ctx, q, "cudaEventQuery");
    }
    if (top_keys && sl.top_stride) {
        // rows of the caller's array are top_stride keys apart; the batch produced min(top, n) keys per query
        if (sl.top != sl.top_stride) memset(top_keys, 0, sl.q_count * sl.top_stride * sizeof(uint64_t));
        for (uint64_t q = 0; q < sl.q_count && sl.top; ++q)
            memcpy(top_keys + q * sl.top_stride, sl.h_keys + q * sl.top, sl.top * sizeof(uint64_t));
    }
    if (device_seconds) {
        float ms = 0.f;
        SWG_CUDA(ctx, cudaEventElapsedTime(&ms, sl.ev_begin, sl.ev_done));
        *device_seconds = ms * 1e-3;
    }
    sl.busy = false;
    *done = 1;
    return SWG_OK;
}

int swg_gpu_get_stats(swg_ctx *ctx, swg_stats *out)
{
    if (!ctx || !out) return fail(ctx, SWG_ERR_ARG, "NULL argument");
    *out = ctx->stats;
    return SWG_OK;
}

int swg_gpu_get_query_seconds(swg_ctx *ctx, double *seconds, uint64_t max_queries)
{
    if (!ctx || !seconds) return fail(ctx, SWG_ERR_ARG, "NULL argument");
    for (uint64_t q = 0; q < max_queries; ++q) seconds[q] = q < ctx->q_seconds.size() ? ctx->q_seconds[q] : 0.0;
    return SWG_OK;
}

int swg_gpu_get_query_kernels(swg_ctx *ctx, int32_t *kind, uint64_t max_queries)
{
    if (!ctx || !kind) return fail(ctx, SWG_ERR_ARG, "NULL argument");
    for (uint64_t q = 0; q < max_queries; ++q) kind[q] = 0;
    for (const WorkItem &it : ctx->items)
        if (it.pair)
            for (uint32_t q : it.members)
                if (q < max_queries) kind[q] = 1;
    return SWG_OK;
}

int swg_plan_describe(const uint16_t *q_lengths, uint64_t q_count, uint64_t n_sequences, uint64_t n_residues,
                      uint32_t longest_sequence, int query_pairing, char *text, uint64_t capacity)
{
    if ((q_count && !q_lengths) || !text || capacity == 0) return fail(nullptr, SWG_ERR_ARG, "NULL argument");
    if (query_pairing < 0 || query_pairing > 2) return fail(nullptr, SWG_ERR_ARG, "query_pairing must be 0, 1 or 2");
    ShardShape shape;
    shape.residues = n_residues;
    shape.maxcols = std::max<uint32_t>(kChunkCols, (longest_sequence + kChunkCols - 1) / kChunkCols * kChunkCols);
    shape.ntiles = (uint32_t)((n_sequences + kTileSeqs - 1) / kTileSeqs);
    shape.lines_fit = true;
    PlanOptions opts;
    opts.query_pairing = query_pairing;
    std::vector<uint16_t> q_len(q_lengths, q_lengths + q_count);
    std::vector<Config> main_cfgs, wide_cfgs;
    std::vector<WorkItem> items;
    plan_batch(q_len, shape, opts, main_cfgs, wide_cfgs, items);
    const std::string out = describe_plan(q_len, main_cfgs, items);
    const size_t n = std::min<size_t>(out.size(), (size_t)capacity - 1);
    memcpy(text, out.data(), n);
    text[n] = 0;
    return SWG_OK;
}

int swg_plan_column_chunks(uint32_t m, int smax, int extend_gap, long option, const uint32_t *tile_cols, uint32_t n_tiles,
                           uint32_t *chunks /* [capacity][3] = tile, first column, columns */, uint64_t capacity, uint64_t *n_chunks,
                           uint64_t *span_bound)
{
    if (!n_chunks || (n_tiles && !tile_cols)) return fail(nullptr, SWG_ERR_ARG, "NULL argument");
    uint32_t maxcols = 8;
    for (uint32_t t = 0; t < n_tiles; ++t) maxcols = std::max(maxcols, tile_cols[t]);
    std::vector<ColumnChunk> out;
    plan_column_chunks(m, smax, extend_gap, option, tile_cols, 0, n_tiles, maxcols, out);
    *n_chunks = out.size();
    if (span_bound) *span_bound = alignment_span_bound(m, smax, extend_gap);
    for (uint64_t i = 0; chunks && i < out.size() && i < capacity; ++i) {
        chunks[3 * i] = out[i].tile;
        chunks[3 * i + 1] = out[i].col0;
        chunks[3 * i + 2] = out[i].cols;
    }
    return SWG_OK;
}

int swg_gpu_debug_read(swg_ctx *ctx, const char *name, void *out, uint64_t max_bytes, uint64_t *bytes)
{
    if (!ctx || !name || !bytes) return fail(ctx, SWG_ERR_ARG, "NULL argument");
    const DeviceBuf *b = nullptr;
    uint64_t n = 0;
    if (!strcmp(name, "db")) { b = &ctx->d_db; n = ctx->db_units * sizeof(uint4); }
    else if (!strcmp(name, "tile_off")) { b = &ctx->d_tile_off; n = ((uint64_t)ctx->ntiles + 1) * sizeof(uint64_t); }
    else if (!strcmp(name, "tile_cols")) { b = &ctx->d_tile_cols; n = (uint64_t)ctx->ntiles * sizeof(uint32_t); }
    else if (!strcmp(name, "profile")) { b = &ctx->d_profile; n = b->cap; }
    else if (!strcmp(name, "profile32")) { b = &ctx->d_profile32; n = b->cap; }
    else if (!strcmp(name, "scores")) { b = &ctx->d_scores; n = ctx->run_window * ctx->ntiles * kTileSeqs * sizeof(int32_t); }
    else if (!strcmp(name, "counters")) { b = &ctx->d_counters; n = ctx->q_count * 4 * sizeof(uint32_t); }
    else return fail(ctx, SWG_ERR_ARG, "unknown buffer '%s'", name);
    *bytes = n;
    if (!out || !n) return SWG_OK;
    if (n > max_bytes) n = max_bytes;
    SWG_CUDA(ctx, cudaSetDevice(ctx->device));
    SWG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    SWG_CUDA(ctx, cudaMemcpy(out, b->p, n, cudaMemcpyDeviceToHost));
    return SWG_OK;
}

int swg_gpu_pipebench(swg_ctx *ctx, int max_probes, double *ginstr_per_s, double *sm_mhz, const char **names, int *n_probes,
                      int *sm_count)
{
    if (!ctx || !ginstr_per_s || !sm_mhz || !n_probes) return fail(ctx, SWG_ERR_ARG, "NULL argument");
    const int n = pipebench_probe_count();
    if (max_probes < n) return fail(ctx, SWG_ERR_ARG, "need room for %d probes", n);
    SWG_CUDA(ctx, cudaSetDevice(ctx->device));
    SWG_CUDA(ctx, run_pipebench_all(ginstr_per_s, sm_mhz, sm_count, ctx->stream));
    if (names)
        for (int p = 0; p < n; ++p) names[p] = pipebench_probe_name(p);
    *n_probes = n;
    return SWG_OK;
}

// ---- the reference's kernel signature (CPUsearch.h:37-39) ------------------------------------------
int swimm_gpu_search_avx2_compat(char *query_sequences, unsigned short int *query_sequences_lengths,
                                 unsigned long int query_sequences_count, unsigned int *query_disp, char *vect_sequences_db,
                                 unsigned short int *vect_sequences_db_lengths, unsigned short int *vect_sequences_db_blocks,
                                 unsigned long int vect_sequences_db_count, unsigned long int *vect_sequences_db_disp,
                                 char *submat, int open_gap, int extend_gap, int n_threads, int cpu_block_size, int *scores,
                                 double *workTime)
{
    (void)vect_sequences_db_blocks;
    (void)cpu_block_size;
    const int vl = 32;
    if (!scores) return fail(nullptr, SWG_ERR_ARG, "scores is NULL");
    // n_threads is read as the number of GPUs (0 or less: all visible ones)
    int visible = 0;
    cudaError_t ce = cudaGetDeviceCount(&visible);
    if (ce != cudaSuccess) return cuda_fail(nullptr, ce, "cudaGetDeviceCount");
    if (visible <= 0) return fail(nullptr, SWG_ERR_NO_DEVICE, "no CUDA device visible (there is no CPU fallback)");
    const int ngpu = n_threads <= 0 ? visible : std::min(n_threads, visible);
    // every lane is searched; lanes that are pure padding have length 0 and score 0, as in the reference
    const uint64_t lanes = (uint64_t)vect_sequences_db_count * vl;
    std::vector<uint64_t> disp(vect_sequences_db_count + 1);
    for (uint64_t g = 0; g <= vect_sequences_db_count; ++g) disp[g] = vect_sequences_db_disp[g];
    std::vector<uint16_t> lengths;
    std::vector<signed char> flat;
    deinterleave((const signed char *)vect_sequences_db, vect_sequences_db_lengths, disp.data(), vl, lanes, lengths, flat);
    // one context per GPU, each with the tiles t % ngpu == g; all of them run at the same time, every fetch fills the
    // score entries of its own shard
    std::vector<swg_ctx *> ctxs((size_t)ngpu, nullptr);
    int st = SWG_OK;
    char err[512] = "";
    for (int g = 0; g < ngpu && st == SWG_OK; ++g) {
        st = swg_gpu_create(g, &ctxs[g]);
        if (st == SWG_OK) st = swg_gpu_load_db(ctxs[g], lengths.data(), flat.data(), lanes, flat.size(), g, ngpu);
        if (st == SWG_OK)
            st = swg_gpu_set_queries(ctxs[g], (const signed char *)query_sequences, query_sequences_lengths, query_disp,
                                     query_sequences_count, (const signed char *)submat, open_gap, extend_gap);
        if (st == SWG_OK) st = swg_gpu_run(ctxs[g], 0, 1);
        if (st != SWG_OK) snprintf(err, sizeof(err), "%s", ctxs[g] ? ctxs[g]->err : g_last_error);
    }
    double work = 0.0;
    for (int g = 0; g < ngpu && st == SWG_OK; ++g) {
        st = swg_gpu_fetch(ctxs[g], scores, nullptr);
        if (st != SWG_OK) snprintf(err, sizeof(err), "%s", ctxs[g]->err);
        work = std::max(work, ctxs[g]->stats.search_seconds);
    }
    for (int g = 0; g < ngpu; ++g) swg_gpu_destroy(ctxs[g]);
    if (st != SWG_OK) snprintf(g_last_error, sizeof(g_last_error), "%s", err);
    else if (workTime) *workTime = work;
    return st;
}

}  // extern "C"
