// swg_plan.cu -- the host-side planner (swg_plan.h): kernel shapes and the launch schedule of a batch, from the
// measured rate tables.  No CUDA call in this file.
#include "swg_plan.h"

#include <algorithm>
#include <cstdio>

#include "swg_common.cuh"

namespace swg {

// ---- mapping a query onto thread groups ---------------------------------------------------------
// The group computes passes * G * K rows of which m are useful; how fast a shape runs (per-column overhead,
// pipeline skew, registers) was measured once per shape on a B200 (tools/calibrate.py -> swg_rates.inc).
// The planner minimises rows / rate.
#include "swg_rates.inc"

double shape_rate(int G, int K, uint32_t passes)
{
    if (passes > 1) return kRateMulti_G32[K];
    switch (G) {
        case 4: return kRateSingle_G4[K];
        case 8: return kRateSingle_G8[K];
        case 16: return kRateSingle_G16[K];
        default: return kRateSingle_G32[K];
    }
}

Config wide_config(uint32_t m);

// `residues` and `maxcols` describe the shard: on a small one the longest sequence's serial chain (one column per
// step of its thread group, about 24 cycles per row) can outlast the throughput time, and a shape with more threads
// per sequence (fewer rows per thread, a shorter step) wins although its saturated rate is lower.
Config choose_config(uint32_t m, double residues, double maxcols, long force_group, long force_rows)
{
    Config best = {32, 32, 1, false};
    if (m == 0) m = 1;
    double best_cost = 1e300;
    const Config wide = wide_config(m);
    const double wide_chain = maxcols * wide.passes * (24.0 * wide.K + 100.0);
    for (int G = 4; G <= 32; G *= 2) {
        if (force_group && G != force_group) continue;
        for (int K = 1; K <= kMaxRowsPerThread; ++K) {
            if (force_rows && K != force_rows) continue;
            const uint32_t rows = (uint32_t)(G * K);
            const uint32_t passes = (m + rows - 1) / rows;
            if (passes > 1 && G != 32) continue;     // the pass boundary line is per warp: one pair per warp
            if (passes > (uint32_t)kMaxSmemPasses && !(force_group || force_rows)) continue;
            const double thr = (double)passes * rows / shape_rate(G, K, passes) * residues * 1e-9;
            // (the longest tiles can be handed to the 32-thread shape, which caps the chain at that shape's)
            const double chain = std::min(maxcols * passes * (24.0 * K + 100.0), wide_chain) / kSmHz;
            const double c = std::max(thr, chain) + 1e-3 * thr;       // ties: the higher throughput
            if (c < best_cost) { best_cost = c; best = {G, K, passes, passes > (uint32_t)kMaxSmemPasses}; }
        }
    }
    if (best_cost == 1e300) {       // nothing fits in shared memory (or forced values cannot hold the query)
        const uint32_t passes = (m + 1023) / 1024;
        best = {32, 32, passes, passes > (uint32_t)kMaxSmemPasses};
    }
    if (best.global_profile) { best.G = 32; best.K = 32; best.passes = (m + 1023) / 1024; }
    return best;
}

// Shape of the query-pair kernel for a pair whose longer query has m rows.  K is even, 8..32.  One pass when
// G*K >= m for some G in {8, 16, 32}; otherwise several passes of 32-thread groups, each pass with its own K: the
// cheapest multiset of pass heights that covers m rows (unbounded knapsack over the measured rates), so that the
// rows a pair computes exceed its length by less than 64.
double q2_rate(int G, int K, bool multi)
{
    const float *tab = multi ? kRateQ2Multi_G32 : G == 8 ? kRateQ2_G8 : G == 16 ? kRateQ2_G16 : kRateQ2_G32;
    return tab[K / 2];
}

PairConfig choose_pair_config(uint32_t m, long force_group, long force_rows)
{
    if (m == 0) m = 1;
    PairConfig best;
    best.G = 32;
    best.cost = 1e300;
    for (int G = 8; G <= 32; G *= 2) {
        if (force_group && G != force_group) continue;
        for (int K = 8; K <= kMaxRowsPerThread; K += 2) {
            if (force_rows && K != force_rows) continue;
            if ((uint32_t)(G * K) < m) continue;
            // the rate tables count the cells of BOTH queries: cost per database residue = rows / (rate / 2)
            const double c = 2.0 * G * K / q2_rate(G, K, false);
            if (c < best.cost) { best.G = G; best.K.assign(1, K); best.cost = c; }
        }
    }
    if (best.cost < 1e300) return best;
    // several passes, 32 threads per sequence: min-cost cover of ceil(m / 64) units with passes of K/2 units each
    const uint32_t units = (m + 63) / 64;
    std::vector<double> cost(units + 1, 1e300);
    std::vector<int> pick(units + 1, 0);
    cost[0] = 0.0;
    for (uint32_t u = 1; u <= units; ++u)
        for (int K = 8; K <= kMaxRowsPerThread; K += 2) {
            if (force_rows && K != force_rows) continue;
            const uint32_t prev = u > (uint32_t)(K / 2) ? u - K / 2 : 0;
            const double c = cost[prev] + 2.0 * 32 * K / q2_rate(32, K, true);
            if (c < cost[u]) { cost[u] = c; pick[u] = K; }
        }
    best.G = 32;
    best.K.clear();
    for (uint32_t u = units; u > 0;) {
        const int K = pick[u];
        best.K.push_back(K);
        u = u > (uint32_t)(K / 2) ? u - K / 2 : 0;
    }
    std::sort(best.K.begin(), best.K.end(), [](int a, int b) { return a > b; });     // tallest passes first
    best.cost = cost[units];
    return best;
}

// The two lanes as independent streams of queries.  `lanes[l]` lists lane l's queries in order; every launch computes
// 32*K rows of both lanes' current queries.  Launch boundaries are put where a query ends (the other lane simply
// continues), so the rows a lane computes exceed its queries' lengths by < 64 per query plus the idle tail of the
// shorter lane.  Returns the launches and their estimated cost (same units as PairConfig::cost).
double plan_stream(const std::vector<uint32_t> lanes[2], const std::vector<uint16_t> &q_len, long force_rows,
                   std::vector<Q2Launch> &out)
{
    size_t pos[2] = {0, 0};
    uint32_t done[2] = {0, 0};
    double cost = 0.0;
    for (;;) {
        int32_t q[2];
        uint32_t rem[2];
        for (int l = 0; l < 2; ++l) {
            q[l] = pos[l] < lanes[l].size() ? (int32_t)lanes[l][pos[l]] : -1;
            rem[l] = q[l] >= 0 ? std::max<uint32_t>(q_len[q[l]], 1u) - done[l] : 0xffffffffu;
        }
        if (q[0] < 0 && q[1] < 0) break;
        uint32_t seg = std::min(rem[0], rem[1]);
        const uint32_t far = std::max(rem[0], rem[1]);
        if (far != 0xffffffffu && far - seg <= 192) seg = far;       // ends this close finish in the same launches
        // the lane whose query ends first has nothing after it: its tail is idle whatever happens, so no launch
        // boundary (and no short, slow launch) is spent on that end
        if (far != 0xffffffffu && seg != far) {
            const int ls = rem[0] <= rem[1] ? 0 : 1;
            if (pos[ls] + 1 >= lanes[ls].size()) seg = far;
        }
        const PairConfig pc = choose_pair_config(seg, 32, force_rows);
        bool fin[2] = {false, false};
        for (int K : pc.K) {
            Q2Launch L;
            L.G = pc.G;
            L.K = K;
            const uint32_t rows = (uint32_t)(pc.G * K);
            for (int l = 0; l < 2; ++l) {
                if (q[l] >= 0 && !fin[l]) {
                    const uint32_t len = std::max<uint32_t>(q_len[q[l]], 1u);
                    L.lane[l] = {q[l], done[l], done[l] == 0, done[l] + rows >= len};
                    done[l] += rows;
                    fin[l] = L.lane[l].last;
                } else {
                    L.lane[l] = {-1, 0, false, false};
                }
            }
            out.push_back(L);
            cost += 2.0 * rows / q2_rate(pc.G, K, true);
        }
        for (int l = 0; l < 2; ++l)
            if (fin[l]) { ++pos[l]; done[l] = 0; }
    }
    return cost;
}

// the 32-thread shape used for long tiles and for the 32-bit re-computation
Config wide_config(uint32_t m)
{
    if (m == 0) m = 1;
    Config c;
    c.G = 32;
    if (m <= 1024) { c.K = (int)((m + 31) / 32); c.passes = 1; }
    else {
        // fewest passes, then the fewest rows per thread that still cover the query
        c.passes = (m + 1023) / 1024;
        c.K = (int)((m + 32 * c.passes - 1) / (32 * c.passes));
    }
    c.global_profile = c.passes > (uint32_t)kMaxSmemPasses;
    if (c.global_profile) c.K = 32;
    return c;
}

constexpr double kGpuWarps = 148.0 * 16.0;      // resident warps the rate tables were measured with

double step_seconds_loaded(int K, double rate_gcups) { return 64.0 * K * kGpuWarps / (rate_gcups * 1e9); }
double step_seconds_alone(int K) { return (26.0 * K + 120.0) / kSmHz; }
double xw_rate(int K) { return kRateXw[K]; }

XwConfig choose_xw_config(uint32_t m, double pairs, double pair_columns, double maxcols, int ctas, long force_warps,
                          long force_rows)
{
    XwConfig best;
    if (m == 0) m = 1;
    if (ctas < 1) ctas = 1;
    double best_cost = 1e300;
    for (int W = 1; W <= 16; W *= 2)
        for (int K = 1; K <= kMaxRowsPerThread; ++K) {
            if ((force_warps && W != force_warps) || (force_rows && K != force_rows)) continue;
            if ((uint32_t)(W * 32 * K) < m) continue;
            if (W * ((K + 15) / 16) > 16) continue;                 // W passes of 12.5 KB (K <= 16) or 25 KB in shared memory
            for (int groups = 16 / W; groups >= 1; groups /= 2) {
                // a step is issue-bound when the SM's 16 warp slots are busy and latency-bound when few warps run
                const double active = (double)groups * W;
                const double step = std::max(step_seconds_alone(K), step_seconds_loaded(K, xw_rate(K)) * active / 16.0);
                const double slots = (double)ctas * groups;         // pairs in flight
                const double per_slot = std::max(pair_columns / slots, std::min(pairs, 1.0) * (maxcols + 40.0 * W));
                const double c = per_slot * step * (1.0 + 1e-3 * K);          // ties: fewer rows per thread
                if (c < best_cost) { best_cost = c; best.K = K; best.W = W; best.groups = groups; best.seconds = per_slot * step; }
            }
        }
    return best;
}

uint64_t alignment_span_bound(uint32_t m, int smax, int ge)
{
    if (ge < 1 || smax < 1) return 0;
    if (m == 0) m = 1;
    return (uint64_t)m + (uint64_t)m * (uint64_t)smax / (uint64_t)ge + 1;
}

uint32_t plan_column_chunks(uint32_t m, int smax, int ge, long option, const uint32_t *tile_cols, uint32_t first_tile,
                            uint32_t ntiles, uint32_t maxcols, std::vector<ColumnChunk> &out)
{
    out.clear();
    const uint64_t B = alignment_span_bound(m, smax, ge);
    if (B == 0 || option == 1) return 0;
    const uint64_t want = option > 1 ? (uint64_t)option : std::max<uint64_t>(2 * B, 2048);
    const uint32_t C = (uint32_t)std::min<uint64_t>((std::max(want, B + 8) + 7) / 8 * 8, 1u << 20);
    const uint32_t stride = (uint32_t)((C - B) / 8 * 8);            // chunk starts stay multiples of 8 (the layout's column chunks)
    if (stride < 8 || (uint64_t)C * 2 > maxcols) return 0;
    for (uint32_t t = ntiles; t-- > first_tile;) {
        const uint32_t cols = tile_cols[t];
        if (cols <= C) { out.push_back({t, 0, cols}); continue; }
        for (uint32_t c0 = 0;; c0 += stride) {
            const uint32_t len = std::min(C, cols - c0);
            out.push_back({t, c0, len});
            if (c0 + len >= cols) break;
        }
    }
    std::stable_sort(out.begin(), out.end(), [](const ColumnChunk &a, const ColumnChunk &b) { return a.cols > b.cols; });
    return C;
}

// ---- the batch ------------------------------------------------------------------------------------
void plan_batch(const std::vector<uint16_t> &q_len, const ShardShape &shard, const PlanOptions &opt,
                std::vector<Config> &main_cfgs, std::vector<Config> &wide_cfgs, std::vector<WorkItem> &items)
{
    const uint64_t nq = q_len.size();
    main_cfgs.assign(nq, Config());
    wide_cfgs.assign(nq, Config());
    for (uint64_t q = 0; q < nq; ++q) {
        main_cfgs[q] = choose_config(q_len[q], (double)shard.residues, (double)shard.maxcols, opt.force_group, opt.force_rows);
        wide_cfgs[q] = wide_config(q_len[q]);
    }

    // ---- schedule ----
    // Batches of at least two queries use the query-pair kernel where the planner expects it to beat the
    // sequence-pair kernel: (1) the queries longer than one pass (1024 rows) are dealt to the two 16-bit lanes
    // (longest first, to the lane with fewer rows) and run as two streams, pass by pass; (2) the shorter ones are
    // paired with their neighbour in length for one single-pass launch.  Everything else runs one query at a time.
    items.clear();
    {
        std::vector<uint32_t> order(nq);
        for (uint64_t q = 0; q < nq; ++q) order[q] = (uint32_t)q;
        std::stable_sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) { return q_len[x] > q_len[y]; });
        // Estimated seconds of a launch: its throughput time, or the serial chain of the longest sequence when that
        // is longer (a sequence advances one column per step of its thread group: about 24 cycles per row on a busy
        // SM, measured on the long-sequence workload).  The same model for both kernels, so that they compare fairly.
        const double res9 = (double)shard.residues * 1e-9;
        auto chain_seconds = [&](int K, uint32_t passes) {
            return (double)shard.maxcols * passes * (24.0 * K + 100.0) / kSmHz;
        };
        auto single_cost = [&](uint32_t q) {
            const Config &c = main_cfgs[q];
            return std::max((double)c.passes * c.G * c.K / shape_rate(c.G, c.K, c.passes) * res9, chain_seconds(c.K, c.passes));
        };
        auto q2_cost = [&](const std::vector<Q2Launch> &ls) {
            double t = 0.0;
            for (const Q2Launch &L : ls)
                t += std::max(2.0 * L.G * L.K / q2_rate(L.G, L.K, ls.size() > 1) * res9, chain_seconds(L.K, 1));
            return t;
        };
        // One candidate schedule: queries longer than `stream_above` rows go to the two streams, the others are
        // paired with their neighbour in length.  Returns the estimated seconds of the whole batch.
        auto build = [&](uint32_t stream_above, std::vector<WorkItem> &out) -> double {
            std::vector<char> taken(nq, 0);
            double total = 0.0;
            auto add_group = [&](const std::vector<uint32_t> &members, std::vector<Q2Launch> &launches, double cost) {
                WorkItem it;
                it.pair = true;
                it.qa = members[0];
                it.members = members;
                it.launches.swap(launches);
                for (uint32_t q : members) { taken[q] = 1; it.member_rows += q_len[q]; }
                out.push_back(std::move(it));
                total += cost;
            };
            // (1) two streams; the longest queries may be left to the sequence-pair kernel when that balances the lanes
            std::vector<uint32_t> longq;
            for (uint32_t q : order)
                if (q_len[q] > stream_above) longq.push_back(q);
            if (longq.size() >= 2 && shard.lines_fit) {
                double best_cost = 1e300;
                size_t best_skip = 0;
                std::vector<Q2Launch> best_launches;
                const size_t max_skip = opt.query_pairing == 2 ? 0 : std::min<size_t>(2, longq.size() - 2);
                double skipped_cost = 0.0;
                for (size_t skip = 0; skip <= max_skip; ++skip) {
                    std::vector<uint32_t> lanes[2];
                    uint64_t rows[2] = {0, 0};
                    for (size_t i = skip; i < longq.size(); ++i) {
                        const int l = rows[1] < rows[0] ? 1 : 0;
                        lanes[l].push_back(longq[i]);
                        rows[l] += q_len[longq[i]];
                    }
                    std::vector<Q2Launch> ls;
                    plan_stream(lanes, q_len, opt.q2_rows, ls);
                    const double c = q2_cost(ls);
                    if (c + skipped_cost < best_cost) { best_cost = c + skipped_cost; best_skip = skip; best_launches.swap(ls); }
                    skipped_cost += single_cost(longq[skip]);
                }
                double all_single = 0.0, skipped = 0.0;
                for (uint32_t q : longq) all_single += single_cost(q);
                for (size_t i = 0; i < best_skip; ++i) skipped += single_cost(longq[i]);
                if (opt.query_pairing == 2 || best_cost < all_single) {
                    std::vector<uint32_t> members(longq.begin() + best_skip, longq.end());
                    add_group(members, best_launches, best_cost - skipped);
                }
            }
            // (2) single-pass pairs of neighbours among the rest
            std::vector<uint32_t> rest;
            for (uint32_t q : order)
                if (!taken[q] && q_len[q] <= std::min<uint32_t>(stream_above, kMaxPassRows)) rest.push_back(q);
            for (size_t i = 0; i + 1 < rest.size(); i += 2) {
                const uint32_t qb = rest[i], qa = rest[i + 1];          // qb is the longer one
                const PairConfig pc = choose_pair_config(q_len[qb], opt.q2_group, opt.q2_rows);
                std::vector<Q2Launch> ls;
                if (pc.passes() == 1) {
                    Q2Launch L;
                    L.G = pc.G;
                    L.K = pc.K[0];
                    L.lane[0] = {(int32_t)qa, 0, true, true};
                    L.lane[1] = {(int32_t)qb, 0, true, true};
                    ls.push_back(L);
                } else if (shard.lines_fit) {                                   // forced rows too few for one pass
                    std::vector<uint32_t> lanes[2] = {{qa}, {qb}};
                    plan_stream(lanes, q_len, opt.q2_rows, ls);
                } else {
                    continue;
                }
                const double cost = q2_cost(ls);
                if (opt.query_pairing == 2 || cost < single_cost(qa) + single_cost(qb))
                    add_group({qa, qb}, ls, cost);
            }
            for (uint64_t q = 0; q < nq; ++q)
                if (!taken[q]) {
                    WorkItem it;
                    it.pair = false;
                    it.qa = (uint32_t)q;
                    out.push_back(std::move(it));
                    total += single_cost((uint32_t)q);
                }
            return total;
        };
        if (opt.query_pairing && nq >= 2 && shard.ntiles) {
            // where the streams end and the single-pass pairs begin is a planner choice too
            double best = 1e300;
            for (uint32_t above : {(uint32_t)kMaxPassRows, 832u, 640u, 448u, 256u, 0u}) {
                std::vector<WorkItem> cand;
                const double c = build(above, cand);
                if (c < best) { best = c; items.swap(cand); }
                if (opt.q2_rows || opt.q2_group) break;          // forced shapes (tests): the first candidate
            }
        } else {
            for (uint64_t q = 0; q < nq; ++q) {
                WorkItem it;
                it.pair = false;
                it.qa = (uint32_t)q;
                items.push_back(std::move(it));
            }
        }
    }

}

std::string describe_plan(const std::vector<uint16_t> &q_len, const std::vector<Config> &main_cfgs,
                          const std::vector<WorkItem> &items)
{
    std::string out;
    char line[512];
    for (const WorkItem &it : items) {
        if (!it.pair) {
            const Config &c = main_cfgs[it.qa];
            snprintf(line, sizeof(line), "[swg] query %u (%u rows): sequence-pair kernel G=%d K=%d passes=%u\n", it.qa,
                     (unsigned)q_len[it.qa], c.G, c.K, c.passes);
            out += line;
            continue;
        }
        snprintf(line, sizeof(line), "[swg] query-pair kernel, %zu queries (%llu rows), %zu launches:\n", it.members.size(),
                 (unsigned long long)it.member_rows, it.launches.size());
        out += line;
        for (const Q2Launch &L : it.launches) {
            snprintf(line, sizeof(line),
                     "[swg]   G=%d K=%d | lane 0: q %d (%u rows) from row %u%s%s | lane 1: q %d (%u rows) from row %u%s%s\n", L.G,
                     L.K, L.lane[0].q, L.lane[0].q >= 0 ? (unsigned)q_len[L.lane[0].q] : 0u, L.lane[0].row0,
                     L.lane[0].first ? " first" : "", L.lane[0].last ? " last" : "", L.lane[1].q,
                     L.lane[1].q >= 0 ? (unsigned)q_len[L.lane[1].q] : 0u, L.lane[1].row0, L.lane[1].first ? " first" : "",
                     L.lane[1].last ? " last" : "");
            out += line;
        }
    }
    return out;
}

}  // namespace swg
