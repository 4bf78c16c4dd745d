// wavefront_q2.cuh -- the query-pair variant of the search kernel (K1q): TWO QUERIES against one database
// sequence in the two 16-bit halves of every register (the 16-bit search kernel of wavefront.cuh packs two
// database sequences against one query).  Same recurrence and same systolic pipeline; what changes:
//
//   * both halves of a register see the SAME database letter, so the packed substitution score of a row is one
//     32-bit profile entry {S[qB[i]][letter], S[qA[i]][letter]} read straight from shared memory -- the PRMT that
//     packs the scores of two database letters disappears.  Per cell pair: 3.5 ALU-pipe instructions
//     (VIMNMX3.RELU, 2 x VIADDMNMX, VIMNMX3 / 2) + 2 VIADD.16x2 on the FMA-heavy pipe, against 4.5 + 2
//     (pipebench: mix_q2_3p5alu_2viadd_immediate vs mix_v2_immediate_penalties);
//   * the profile is 4 bytes per (row, letter): 25 letters x 4 KB = 100 KB for 1024 rows, so a CTA keeps the rows of
//     ONE launch in shared memory and longer queries are searched PASS BY PASS, one launch per pass over the whole
//     shard.  The last row (H, F) of a pass is parked in a per-sequence line in HBM (8 bytes per database column,
//     written once and read once by the next launch: ~70 GB/s on the Swiss-Prot-sized benchmark, 1 % of the HBM
//     roofline);
//   * the two lanes are INDEPENDENT STREAMS of queries (swg_api.cu, plan_stream): a launch computes G*K rows of each
//     lane's current query, launch boundaries fall where a query ends, and a lane may start its next query while the
//     other continues one.  Per lane and launch: whether the rows continue from the line (the host clears the half of
//     a lane that starts fresh), whether the score is merged with the stored one, whether the lane is idle;
//   * a group of G threads owns ONE database sequence (the half of a tile pair is picked by the PRMT selector
//     that builds the column word).  Groups of 8 and 16 threads serve single-launch pairs of short queries.
//
// Exactness is unchanged: wrapping 16-bit lanes, a lane (= query) whose running best reaches kOverflow16 in any
// launch is listed once and recomputed by the 32-bit kernel of wavefront.cuh when the query ends.
//
// Used by swg_gpu_run() for batches of at least two queries; a single query runs the sequence-pair kernel.  Role
// in the reference: the same cpu_search_avx2_sp task loop (CPUsearch.c:482-548), which also walks Qcount x groups
// tasks.
#pragma once

#include <type_traits>

#include "swg_common.cuh"

namespace swg {

// G threads x K rows (K even); CIN: a lane's rows continue from the line of the previous launch; COUT: the last
// row is parked for the next launch; GOE/GE > 0: gap penalties as immediates.
//
// A STEP IS TWO COLUMNS.  Thread t works two columns behind thread t-1, receives the (H, F) of the row above for both
// columns in one round of shuffles, and walks the rows of the two columns interleaved -- row x of the first, then row
// x-1 of the second.  The two walks are independent dependency chains (the second needs only E and the diagonal term of
// the row the first has just finished), so the fixed latency of max3 -> add -> addmax of one chain is covered by the
// other inside the same warp (ncu, one chain per step: 32 % of a warp's cycles in stall_wait, ALU pipe 84.5 % busy).
// Thread 0's inputs are selects (one SEL each: measured 0.9 % faster than an IMAD pair that keeps them off the ALU
// pipe -- an IMAD costs the ALU pipe about half a slot, pipebench pair_viaddmnmx_imad).
template <int G, int K, bool CIN, bool COUT, int GOE, int GE>
__global__ void __launch_bounds__(kBlockThreads, 1) wavefront_q2_kernel(const WfParams p)
{
    typedef uint32_t reg;
    typedef Lane16 L;
    static_assert(G == 8 || G == 16 || G == 32, "group size");
    static_assert(K >= 2 && K <= kMaxRowsPerThread && K % 2 == 0, "rows per thread");
    constexpr int GPW = 32 / G;                       // groups per warp
    constexpr int TPT = kTileSeqs / GPW;              // warp tasks per tile: every group takes one sequence
    constexpr int KCH = (K + 3) / 4;                  // 4-row (16-byte) profile chunks per thread
    constexpr int NC = kTripCols;
    constexpr int SPT = NC / 2;                       // steps per trip
    constexpr int SKEW = 2 * G;                       // columns between the words entering thread 0 and the rows leaving thread G-1
    constexpr uint32_t FI = SKEW / NC;                // head trips: thread G-1 is still on the previous segment
    constexpr uint32_t kMinTrips = FI + 4;
    static_assert(NC == 4, "two steps of two columns per trip");
    static_assert(kMinTrips * NC <= kQ2MinSegCols, "line stride too short for padded segments");

    extern __shared__ __align__(16) uint8_t prof_smem[];
    {
        const uint4 *src = reinterpret_cast<const uint4 *>(p.profile);
        uint4 *dst = reinterpret_cast<uint4 *>(prof_smem);
        // only the part of a letter row that this shape reads
        constexpr int row16 = G * KCH;
        for (int i = threadIdx.x; i < kLetters * row16; i += blockDim.x) {
            const int letter = i / row16, o = i % row16;
            dst[letter * (kQ2LetterStride / 16) + o] = src[letter * (kQ2LetterStride / 16) + o];
        }
    }
    __syncthreads();

    const int lane = threadIdx.x & 31;
    const int t = lane % G;
    const int g = lane / G;
    const uint32_t warp_global = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    // GE > 0: the gap-extend penalty is an immediate (it is the third operand of the two VIADDMNMX per cell: with it in
    // a register they read three registers and take two register-bank cycles); GOE > 0: go + ge as an immediate too
    // (second operand of a VIADD.16x2: no difference in rate, kept for SWIMM's defaults)
    const reg nge = GE > 0 ? L::splat(-GE) : L::splat(-p.gap_extend);
    const reg ngoe = GOE > 0 ? L::splat(-GOE) : L::splat(-p.gap_open_extend);
    const uint32_t pad_pk = 0x60000060u;                         // two pad letters (see the step's word below)
    const uint32_t ntasks = p.tile_count * TPT;
    // a line nobody reads: where the flush segments (and the kernel's first head steps) park their rows
    uint2 *const dummy_line = p.boundary + p.line_dummy + (size_t)(warp_global * GPW + g) * (kQ2MinSegCols + kQ2LineSlack);

    auto fetch_task = [&]() -> uint32_t {
        uint32_t task = 0;
        if (lane == 0) task = atomicAdd(p.task_counter, 1u);
        return __shfl_sync(0xffffffffu, task, 0);
    };

    reg DS[K], E[K];
#pragma unroll
    for (int x = 0; x < K; ++x) { DS[x] = L::splat(0); E[x] = L::splat(0); }
    reg best = L::splat(0), bsave = L::splat(0);
    reg out_ha = L::splat(0), out_fa = L::splat(0), out_hb = L::splat(0), out_fb = L::splat(0);
    uint32_t pk = pad_pk;
    const uint32_t slice_off = (uint32_t)t * 16;
    const bool is_first = t == 0;                                // the thread that feeds the pipeline

    // One step: process the two columns (a, b) whose letters arrived in the previous step with the (H, F) of the row
    // above handed down now, and form DS for b and for the a of the next step from the letters that arrive now (see
    // wavefront.cuh).  The step's word: byte 0 = letter of b, byte 3 = letter of the next a, both as code * 4.
    // restart: the next a is the first column of a new sequence (head steps only: thread t meets it in step t).
    auto step = [&](uint32_t in_pk, uint2 in_a, uint2 in_b, uint2 *store_to, bool restart) {
        uint32_t pkn = __shfl_up_sync(0xffffffffu, pk, 1, G);
        reg r_ha = __shfl_up_sync(0xffffffffu, out_ha, 1, G);
        reg r_fa = __shfl_up_sync(0xffffffffu, out_fa, 1, G);
        reg r_hb = __shfl_up_sync(0xffffffffu, out_hb, 1, G);
        reg r_fb = __shfl_up_sync(0xffffffffu, out_fb, 1, G);
        if (is_first) {
            pkn = in_pk;
            r_ha = CIN ? in_a.x : 0u; r_fa = CIN ? in_a.y : 0u;
            r_hb = CIN ? in_b.x : 0u; r_fb = CIN ? in_b.y : 0u;
        }
        pk = pkn;

        // packed scores of the following columns: letter offset code * 4096.  pkn * 1024 drops byte 3 from the 32-bit
        // product; pkn >> 14 = byte 3 * 1024.  The 16-byte loads are consumed chunk by chunk inside the row loop (short
        // live ranges); the rare segment restart below reads them again.
        const uint4 *qa = reinterpret_cast<const uint4 *>(prof_smem + (pkn * 1024u + slice_off));
        const uint4 *qb = reinterpret_cast<const uint4 *>(prof_smem + ((pkn >> 14) + slice_off));
        reg hpa = r_ha, fa = r_fa, dspa = L::splat(0);
        reg hpb = r_hb, fb = r_fb, dspb = L::splat(0);
        uint4 va = make_uint4(0u, 0u, 0u, 0u), vb = va;
#pragma unroll
        for (int x = 0; x <= K; ++x) {
            if (x < K) {                                       // row x of column a
                if (x % 4 == 0) va = qa[(x / 4) * G];
                const uint32_t wx = x % 4 == 0 ? va.x : x % 4 == 1 ? va.y : x % 4 == 2 ? va.z : va.w;
                const reg ds = DS[x];
                const reg h = L::max3_relu(ds, E[x], fa);      // max(ds, E(i,j), F(i,j), 0)
                const reg open = L::add(h, ngoe);              // H(i,j) - (go+ge)
                E[x] = L::addmax(E[x], nge, open);             // E(i,j+1)
                fa = L::addmax(fa, nge, open);                 // F(i+1,j)
                if (x & 1) best = L::max3(best, dspa, ds);
                dspa = ds;
                DS[x] = L::add(hpa, wx);                       // H(i-1,j) + S(i,j+1)
                hpa = h;
            }
            if (x >= 1) {                                      // row x-1 of column b
                const int y = x - 1;
                if (y % 4 == 0) vb = qb[(y / 4) * G];
                const uint32_t wy = y % 4 == 0 ? vb.x : y % 4 == 1 ? vb.y : y % 4 == 2 ? vb.z : vb.w;
                const reg ds = DS[y];
                const reg h = L::max3_relu(ds, E[y], fb);
                const reg open = L::add(h, ngoe);
                E[y] = L::addmax(E[y], nge, open);
                fb = L::addmax(fb, nge, open);
                if (y & 1) best = L::max3(best, dspb, ds);
                dspb = ds;
                DS[y] = L::add(hpb, wy);
                hpb = h;
            }
        }
        out_ha = hpa; out_fa = fa;
        out_hb = hpb; out_fb = fb;
        if (COUT && t == G - 1) {
            __stcs(store_to, make_uint2(out_ha, out_fa));
            __stcs(store_to + 1, make_uint2(out_hb, out_fb));
        }
        if (restart) {                                 // the next column starts a new sequence
#pragma unroll
            for (int i = 0; i < KCH; ++i) {
                const uint4 v = qb[i * G];
                DS[4 * i] = v.x; DS[4 * i + 1] = v.y;
                if (4 * i + 2 < K) { DS[4 * i + 2] = v.z; DS[4 * i + 3] = v.w; }
            }
#pragma unroll
            for (int x = 0; x < K; ++x) E[x] = L::splat(0);
            bsave = best;
            best = L::splat(0);
        }
    };
    const std::true_type head_steps;
    const std::false_type steady_steps;

    // Store the two scores of a sequence whose last column has left the pipeline.
    auto finalize = [&](uint32_t lseq) {
        reg b = bsave;
#pragma unroll
        for (int o = G >> 1; o > 0; o >>= 1) b = L::max2(b, __shfl_xor_sync(0xffffffffu, b, o, G));
        if (t == 0 && global_seq_index(p, lseq) < p.n_total) {
            const int s_lo = (int)(short)(b & 0xffffu), s_hi = (int)(short)(b >> 16);
            const uint32_t fl = p.lane_flags;
            if (fl & kLaneActive) {
                const int o = (fl & kLaneFirst) ? 0 : p.scores[lseq];
                const int n = max(s_lo, o);
                p.scores[lseq] = n;
                if (n >= kOverflow16 && o < kOverflow16) p.resc_list[atomicAdd(p.resc_count, 1u)] = lseq;
            }
            if ((fl >> 2) & kLaneActive) {
                const int o = ((fl >> 2) & kLaneFirst) ? 0 : p.scores2[lseq];
                const int n = max(s_hi, o);
                p.scores2[lseq] = n;
                if (n >= kOverflow16 && o < kOverflow16) p.resc_list2[atomicAdd(p.resc_count2, 1u)] = lseq;
            }
        }
    };

    // (a lane that starts a new query while the other one continues finds zeros in its half of the lines: the host
    // clears that half between the two launches, clear_lane_kernel)
    bool pending = false;
    uint32_t pend_lseq = 0;
    uint32_t next_task = fetch_task();
    uint2 hf_in0 = make_uint2(0u, 0u), hf_in1 = make_uint2(0u, 0u);
    uint32_t wlast = kPadWord;                           // the last two columns of the previous trip (thread 0)
    // thread G-1 runs SKEW columns behind the words: during the head steps of a segment it is still parking columns of
    // the PREVIOUS segment.  st_cur[c] / st_prev[c] = where the rows leaving in the step that takes in column c go.
    uint2 *st_cur = dummy_line + kQ2MinSegCols - SKEW;     // as if a flush segment had just ended

    for (;;) {
        const uint32_t task = next_task;
        const bool have = task < ntasks;
        if (!have && !pending) break;
        // (the next task is drawn 16 trips before this one ends, not at its start: see wavefront.cuh)

        uint32_t half = 0, lseq = 0, ncols = 0;
        const uint2 *words = nullptr;
        uint2 *line = dummy_line;
        if (have) {
            const uint32_t tile = p.tile_first + p.tile_count - 1 - task / TPT;        // longest tiles first
            const uint32_t seq = (task % TPT) * GPW + g;
            lseq = tile * kTileSeqs + seq;
            half = seq & 1u;
            ncols = p.tile_cols[tile];
            words = reinterpret_cast<const uint2 *>(p.db + p.tile_off[tile] + (seq >> 1));
            if (CIN || COUT) {
                const uint32_t stride = (ncols < (uint32_t)kQ2MinSegCols ? (uint32_t)kQ2MinSegCols : ncols) + kQ2LineSlack;
                line = p.boundary + p.line_off[tile] + (size_t)seq * stride;
            }
        }
        const uint32_t data_trips = ncols / NC;
        const uint32_t trips = data_trips < kMinTrips ? kMinTrips : data_trips;
        const uint32_t seg_cols = trips * NC;
        // the step's word from two consecutive 4-byte words of the sequence pair (2 columns x 2 sequences each): byte 0 =
        // this sequence's residue in the second column of the first word, byte 3 = in the first column of the second
        // word, bytes 1 and 2 = 0 (sign replication of a byte below 128)
        const uint32_t sel = 0x0882u | half | ((4u + half) << 12);
        uint2 *const st_prev = st_cur;                            // st_cur of the previous segment + its column count
        st_cur = line - SKEW;

        uint2 w = make_uint2(kPadWord, kPadWord);
        if (t == 0 && data_trips > 0) w = words[0];
        uint2 ring[NC];                                           // ring[j]: (H, F) above row 0 for column c0 + j - 2
#pragma unroll
        for (int j = 0; j < NC; ++j) ring[j] = make_uint2(0u, 0u);
        if (CIN) {
            ring[0] = hf_in0;
            ring[1] = hf_in1;
            if (t == 0 && have) {
                ring[2] = __ldcs(line);
                ring[3] = __ldcs(line + 1);
            }
        }
        auto trip_body = [&](auto head_tag, uint32_t trip) {
            constexpr bool HEAD = decltype(head_tag)::value;
            // the words of the next trip, as two 4-byte loads: an 8-byte load wants a register pair, w.y moves on to wlast,
            // and the copy out of the pair that ptxas then places right behind the load waits for the whole latency
            uint2 nw = make_uint2(kPadWord, kPadWord);
            const uint32_t nt = trip + 1;
            if (t == 0 && nt < data_trips) {
                const uint32_t *src = reinterpret_cast<const uint32_t *>(words + ((nt >> 1) * (kTilePairs * 2) + (nt & 1u)));
                nw.x = __ldg(src);
                nw.y = __ldg(src + 1);
            }
            const uint32_t c0 = trip * NC;
            uint2 *const st = (HEAD ? st_prev : st_cur) + c0;
#pragma unroll
            for (int j = 0; j < SPT; ++j) {
                const uint32_t pkn = j == 0 ? prmt(wlast, w.x, sel) : prmt(w.x, w.y, sel);
                const uint2 hfa = CIN ? ring[2 * j] : make_uint2(0u, 0u);
                const uint2 hfb = CIN ? ring[2 * j + 1] : make_uint2(0u, 0u);
                if (CIN && t == 0 && have) {
                    ring[2 * j] = __ldcs(line + c0 + 2 * j + 2);
                    ring[2 * j + 1] = __ldcs(line + c0 + 2 * j + 3);
                }
                step(pkn, hfa, hfb, st + 2 * j, HEAD && trip * SPT + j == (uint32_t)t);
            }
            wlast = w.y;
            w = nw;
        };
#pragma unroll 1
        for (uint32_t trip = 0; trip < FI; ++trip) trip_body(head_steps, trip);
        if (pending) { finalize(pend_lseq); pending = false; }
        const uint32_t fetch_trip = have ? (trips > FI + 16 ? trips - 16 : FI) : 0xffffffffu;
#pragma unroll 1
        for (uint32_t trip = FI; trip < trips; ++trip) {
            if (trip == fetch_trip) next_task = fetch_task();
            trip_body(steady_steps, trip);
        }
        if (CIN) { hf_in0 = ring[0]; hf_in1 = ring[1]; }
        // the next segment reads the carried word with ITS selector (the other sequence of the pair, or another pair):
        // leave this sequence's last residue in every byte
        wlast = prmt(wlast, wlast, 0x2222u | (half * 0x1111u));
        st_cur += seg_cols;                                       // the next segment's head steps continue this line
        if (have) { pending = true; pend_lseq = lseq; }
    }
}

// host-side launchers (wavefront_q2_inst_*.cu); K even, 8..32.  Groups of 8 and 16 threads are instantiated for
// single-pass pairs only (a pair that needs several passes always runs 32-thread groups: fewer, longer passes).
cudaError_t launch_q2_g8(int K, int grid, cudaStream_t stream, const WfParams &p);
cudaError_t launch_q2_g16(int K, int grid, cudaStream_t stream, const WfParams &p);
cudaError_t launch_q2_g32(int K, int grid, cudaStream_t stream, const WfParams &p);
cudaError_t launch_q2_g32_first(int K, int grid, cudaStream_t stream, const WfParams &p);    // pass 0 of several
cudaError_t launch_q2_g32_middle(int K, int grid, cudaStream_t stream, const WfParams &p);
cudaError_t launch_q2_g32_last(int K, int grid, cudaStream_t stream, const WfParams &p);

template <int G, int K, bool CIN, bool COUT, int GOE = 0, int GE = 0>
cudaError_t launch_q2_one(int grid, cudaStream_t stream, const WfParams &p)
{
    cudaError_t e = ensure_dynamic_smem(reinterpret_cast<const void *>(&wavefront_q2_kernel<G, K, CIN, COUT, GOE, GE>), kQ2ProfileBytes);
    if (e != cudaSuccess) return e;
    wavefront_q2_kernel<G, K, CIN, COUT, GOE, GE><<<grid, kBlockThreads, kQ2ProfileBytes, stream>>>(p);
    return cudaGetLastError();
}

template <int G, bool CIN, bool COUT>
cudaError_t launch_q2_family(int K, int grid, cudaStream_t stream, const WfParams &p)
{
    constexpr int FO = kFastGapOpenExtend, FE = kFastGapExtend;
    if (p.gap_open_extend == FO && p.gap_extend == FE) {
        switch (K) {
#define SWG_CASE(k) case k: return launch_q2_one<G, k, CIN, COUT, FO, FE>(grid, stream, p);
            SWG_CASE(8) SWG_CASE(10) SWG_CASE(12) SWG_CASE(14) SWG_CASE(16) SWG_CASE(18) SWG_CASE(20)
            SWG_CASE(22) SWG_CASE(24) SWG_CASE(26) SWG_CASE(28) SWG_CASE(30) SWG_CASE(32)
#undef SWG_CASE
            default: return cudaErrorInvalidValue;
        }
    }
    // gap-extend penalties 1 and 2 (with any gap-open penalty) cover BLAST's, SSEARCH's and SWIMM's usual settings
    if (p.gap_extend == 1) {
        switch (K) {
#define SWG_CASE(k) case k: return launch_q2_one<G, k, CIN, COUT, 0, 1>(grid, stream, p);
            SWG_CASE(8) SWG_CASE(10) SWG_CASE(12) SWG_CASE(14) SWG_CASE(16) SWG_CASE(18) SWG_CASE(20)
            SWG_CASE(22) SWG_CASE(24) SWG_CASE(26) SWG_CASE(28) SWG_CASE(30) SWG_CASE(32)
#undef SWG_CASE
            default: return cudaErrorInvalidValue;
        }
    }
    if (p.gap_extend == 2) {
        switch (K) {
#define SWG_CASE(k) case k: return launch_q2_one<G, k, CIN, COUT, 0, 2>(grid, stream, p);
            SWG_CASE(8) SWG_CASE(10) SWG_CASE(12) SWG_CASE(14) SWG_CASE(16) SWG_CASE(18) SWG_CASE(20)
            SWG_CASE(22) SWG_CASE(24) SWG_CASE(26) SWG_CASE(28) SWG_CASE(30) SWG_CASE(32)
#undef SWG_CASE
            default: return cudaErrorInvalidValue;
        }
    }
    switch (K) {
#define SWG_CASE(k) case k: return launch_q2_one<G, k, CIN, COUT>(grid, stream, p);
        SWG_CASE(8) SWG_CASE(10) SWG_CASE(12) SWG_CASE(14) SWG_CASE(16) SWG_CASE(18) SWG_CASE(20)
        SWG_CASE(22) SWG_CASE(24) SWG_CASE(26) SWG_CASE(28) SWG_CASE(30) SWG_CASE(32)
#undef SWG_CASE
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace swg
