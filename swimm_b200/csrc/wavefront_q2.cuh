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
    constexpr uint32_t FI = G / NC;
    constexpr uint32_t kMinTrips = FI + 4;
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
    const uint32_t pad_pk = 0x00006000u;
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
    reg out_h = L::splat(0), out_f = L::splat(0);
    uint32_t pk = pad_pk;
    const uint32_t slice_off = (uint32_t)t * 16;
    uint32_t first_mask = t == 0 ? 0xffffffffu : 0u;            // all ones in the thread that feeds the pipeline
    asm volatile("" : "+r"(first_mask));                         // opaque: keeps the blends below from turning back into selects

    // One step: process the column whose word arrived in the previous step with the (H, F) of the row above handed
    // down now, and form DS for the column whose word arrives now (see wavefront.cuh).  HEAD steps (the first G of
    // a segment) look at the segment mark; steady-state steps do not.
    auto column = [&](auto head_tag, uint32_t in_pkn, uint2 in_hf, uint2 *store_to) {
        constexpr bool HEAD = decltype(head_tag)::value;
        uint32_t pkn = __shfl_up_sync(0xffffffffu, pk, 1, G);
        reg r_h = __shfl_up_sync(0xffffffffu, out_h, 1, G);
        reg r_f = __shfl_up_sync(0xffffffffu, out_f, 1, G);
        // thread 0 takes the entering column word and the (H, F) above row 0 instead: a mask blend (one LOP3 each,
        // no predicate to keep alive across the row loop)
        pkn ^= (pkn ^ in_pkn) & first_mask;
        r_h = CIN ? (r_h ^ ((r_h ^ in_hf.x) & first_mask)) : (r_h & ~first_mask);
        r_f = CIN ? (r_f ^ ((r_f ^ in_hf.y) & first_mask)) : (r_f & ~first_mask);
        pk = pkn;

        // packed scores of the NEXT column: letter offset code*4096 (the word carries code*4 in byte lane 1).  The
        // 16-byte loads are consumed chunk by chunk inside the row loop (short live ranges: K <= 32 rows fit in 128
        // registers); the rare segment restart below reads them again instead of keeping them alive.
        const uint4 *q = reinterpret_cast<const uint4 *>(prof_smem + (((pkn & 0xff00u) << 2) | slice_off));
        reg hp = r_h, f = r_f, dsprev = L::splat(0);
#pragma unroll
        for (int i = 0; i < KCH; ++i) {
            const uint4 v = q[i * G];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int x = 4 * i + k;
                if (x >= K) continue;                          // K = 4n + 2: the last chunk is half used
                const uint32_t wx = k == 0 ? v.x : k == 1 ? v.y : k == 2 ? v.z : v.w;
                const reg ds = DS[x];
                const reg h = L::max3_relu(ds, E[x], f);       // max(ds, E(i,j), F(i,j), 0)
                const reg open = L::add(h, ngoe);              // H(i,j) - (go+ge)
                E[x] = L::addmax(E[x], nge, open);             // E(i,j+1)
                f = L::addmax(f, nge, open);                   // F(i+1,j)
                if (x & 1) best = L::max3(best, dsprev, ds);
                dsprev = ds;
                DS[x] = L::add(hp, wx);                        // H(i-1,j) + S(i,j+1)
                hp = h;
            }
        }
        out_h = hp;
        out_f = f;
        if (COUT && t == G - 1) __stcs(store_to, make_uint2(out_h, out_f));
        if (HEAD && (pkn & kQ2MarkSegment)) {          // the next column starts a new sequence
#pragma unroll
            for (int i = 0; i < KCH; ++i) {
                const uint4 v = q[i * G];
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
    uint2 hf_in = make_uint2(0u, 0u);
    // thread G-1 runs G columns behind the words: during the head steps of a segment it is still parking columns of
    // the PREVIOUS segment.  st_cur[s] / st_prev[s] = where the row processed in step s of the current segment goes.
    uint2 *st_cur = dummy_line + kQ2MinSegCols - G;     // as if a flush segment had just ended

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
        const uint32_t selA = 0x7604u | (half << 4);              // residue byte of this sequence -> byte lane 1
        const uint32_t selB = 0x7604u | ((2u + half) << 4);
        uint2 *const st_prev = st_cur;                            // st_cur of the previous segment + its column count
        st_cur = line - G;

        uint2 w = make_uint2(kPadWord, kPadWord);
        if (t == 0 && data_trips > 0) w = words[0];
        uint2 ring[NC];
#pragma unroll
        for (int j = 0; j < NC; ++j) ring[j] = make_uint2(0u, 0u);
        if (CIN) {
            ring[0] = hf_in;
            if (t == 0 && have) {
#pragma unroll
                for (int j = 1; j < NC; ++j) ring[j] = __ldcs(line + j - 1);
            }
        }
        auto trip_body = [&](auto head_tag, uint32_t trip) {
            constexpr bool HEAD = decltype(head_tag)::value;
            uint2 nw = make_uint2(kPadWord, kPadWord);
            const uint32_t nt = trip + 1;
            if (t == 0 && nt < data_trips) nw = words[(nt >> 1) * (kTilePairs * 2) + (nt & 1u)];
            const uint32_t c0 = trip * NC;
            uint2 *const st = (HEAD ? st_prev : st_cur) + c0;
#pragma unroll
            for (int j = 0; j < NC; ++j) {
                const uint32_t word = (j < 2) ? w.x : w.y;
                const bool first = HEAD && j == 0 && trip == 0;
                const uint32_t pkn = prmt(word, first ? kQ2MarkSegment : 0u, (j & 1) ? selB : selA);
                const uint2 hf = CIN ? ring[j] : make_uint2(0u, 0u);
                if (CIN && t == 0 && have) ring[j] = __ldcs(line + c0 + j + (NC - 1));
                column(head_tag, pkn, hf, st + j);
            }
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
        if (CIN) hf_in = ring[0];
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
