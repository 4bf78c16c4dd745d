// swg_plan.h -- the host-side planner of libswimm_cuda.so: how a batch of queries is mapped onto the two search
// kernels.  Pure host code (no CUDA call): a query length, the shape of the resident shard and the measured rate
// tables (swg_rates.inc) go in, kernel shapes and a launch schedule come out.  swg_gpu_run() executes the plan;
// swg_plan_describe() (include/swimm_gpu.h) prints it, with or without a GPU.
#pragma once

#include <stdint.h>

#include <string>
#include <vector>

namespace swg {

constexpr int kMaxSmemPasses = 7;            // sequence-pair kernel: 7 x 32 KB profile slices per CTA
constexpr double kSmHz = 1.9e9;              // SM clock the time estimates assume

struct Config {          // how one query is mapped onto thread groups
    int G, K;
    uint32_t passes;
    bool global_profile;
};

struct XwConfig {        // how one query is mapped onto the long-sequence kernel (wavefront_xw.cuh)
    int K = 0;           // rows per thread
    int W = 0;           // warps (= concurrent passes of 32*K rows) per sequence pair; 0: the query does not fit (> 8192 rows)
    int groups = 0;      // sequence pairs a CTA works on at the same time (<= 16 / W): fewer groups = shorter steps
    double seconds = 0;  // estimated run time on `ctas` SMs
    bool wide = false;   // instead: the 32-thread single-pass shape of the sequence-pair kernel (W = 1, queries <= 1024 rows)
    bool ok() const { return W > 0; }
};

struct PairConfig {      // how a query pair is mapped onto the query-pair kernel (wavefront_q2.cuh)
    int G;
    std::vector<int> K;  // rows per thread of every pass (one launch per pass; a single-pass pair has one entry)
    double cost;         // estimated run time per database residue (units shared with the single-query planner)
    uint32_t passes() const { return (uint32_t)K.size(); }
    uint32_t rows() const { uint32_t r = 0; for (int k : K) r += (uint32_t)(G * k); return r; }
};

struct LaneSlice {       // what one 16-bit lane does in one launch of the query-pair kernel
    int32_t q;           // query index, -1: the lane is idle
    uint32_t row0;       // first row of the query this launch computes
    bool first, last;    // the query's first / last launch
};

struct Q2Launch {
    int G, K;
    LaneSlice lane[2];
};

struct WorkItem {        // one entry of a batch's schedule
    bool pair;           // false: query qa on the sequence-pair kernel; true: a group of queries on the query-pair kernel
    uint32_t qa;
    std::vector<uint32_t> members;      // the group's queries
    std::vector<Q2Launch> launches;     // the group's launches, in stream order
    uint64_t member_rows = 0;
};

struct ShardShape {      // what the planner needs to know about the resident database shard
    uint64_t residues = 0;
    uint32_t maxcols = 8;     // columns of the longest tile
    uint32_t ntiles = 0;
    bool lines_fit = true;    // the pass lines of multi-launch groups (8 B per database column) fit in device memory
};

struct PlanOptions {     // swg_gpu_set_option values the planner honours
    long query_pairing = 1;   // 0: never pair queries, 1: where the planner expects a gain, 2: always
    long q2_group = 0, q2_rows = 0;          // forced shape of the query-pair kernel
    long force_group = 0, force_rows = 0;    // forced shape of the sequence-pair kernel
};

// measured GCUPS of a kernel shape on a large database (swg_rates.inc)
double shape_rate(int G, int K, uint32_t passes);
double q2_rate(int G, int K, bool multi);

// sequence-pair kernel: shape for one query of m rows on this shard; the 32-thread shape of long tiles / 32-bit redo
Config choose_config(uint32_t m, double residues, double maxcols, long force_group, long force_rows);
Config wide_config(uint32_t m);

// Time model of a sequence's serial chain.  A warp advances its sequences one column per STEP; on a busy SM (16 warps)
// a step of K rows takes 64 * K cells / (the kernel's measured rate per warp), on a nearly idle SM it is bounded by the
// dependent instruction chain instead (about 26 cycles per row + 120 per column).
double step_seconds_loaded(int K, double rate_gcups);
double step_seconds_alone(int K);
double xw_rate(int K);
// long-sequence kernel: W warps x 32 threads x K rows >= m with the profile of all W passes in shared memory; the shape
// and the number of concurrent pairs per CTA that minimise the estimated time of `pairs` sequence pairs with
// `pair_columns` columns in total, the longest `maxcols` columns, on `ctas` SMs
XwConfig choose_xw_config(uint32_t m, double pairs, double pair_columns, double maxcols, int ctas, long force_warps = 0,
                          long force_rows = 0);

// Column chunks of long tiles for a query of one pass (m <= 1024 rows).  An alignment with a positive score has fewer
// than m * Smax / ge database-only columns (each costs at least ge, the aligned pairs earn at most m * Smax), so it spans
// at most B = m + m * Smax / ge + 1 columns.  Chunks of C columns whose starts are (C - overlap) apart with overlap >= B
// therefore contain every such alignment entirely in at least one chunk, and a sequence's score is the best score of
// its chunks -- exactly; the chunks are independent tasks: no serial chain is left.
struct ColumnChunk { uint32_t tile, col0, cols; };
uint64_t alignment_span_bound(uint32_t m, int smax, int ge);       // B; 0 when there is no bound (ge < 1 or smax < 1)
// tiles [first_tile, ntiles) with tile_cols[] columns each (multiples of 8) -> chunks, longest first.  option: 0 = chunk
// length max(2B, 2048), 1 = off, > 1 = that many columns (at least B + 8).  Returns C, or 0 (and no chunks) when
// chunking is off, unbounded, or would not at least halve the longest tile (2 C > maxcols).
uint32_t plan_column_chunks(uint32_t m, int smax, int ge, long option, const uint32_t *tile_cols, uint32_t first_tile,
                            uint32_t ntiles, uint32_t maxcols, std::vector<ColumnChunk> &out);

// query-pair kernel: launch heights covering m rows; the two lanes as streams of queries
PairConfig choose_pair_config(uint32_t m, long force_group, long force_rows);
double plan_stream(const std::vector<uint32_t> lanes[2], const std::vector<uint16_t> &q_len, long force_rows,
                   std::vector<Q2Launch> &out);

// The whole batch: per-query shapes for the sequence-pair kernel (main_cfgs, wide_cfgs) and the schedule (items).
void plan_batch(const std::vector<uint16_t> &q_len, const ShardShape &shard, const PlanOptions &opt,
                std::vector<Config> &main_cfgs, std::vector<Config> &wide_cfgs, std::vector<WorkItem> &items);
std::string describe_plan(const std::vector<uint16_t> &q_len, const std::vector<Config> &main_cfgs,
                          const std::vector<WorkItem> &items);

}  // namespace swg
