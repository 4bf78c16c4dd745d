// wavefront.cuh -- the Smith-Waterman search kernel of the B200 build (replaces the reference's
// cpu_search_avx2_sp inner loops, CPUsearch.c:553-956).
//
// Parallelisation (inter-task across groups, anti-diagonal wavefront inside a group):
//   * a GROUP of G threads (G = 4, 8, 16 or 32 lanes of one warp) aligns the query against one PAIR of
//     database sequences; the two sequences live in the two 16-bit halves of every register (Lane16) and
//     are advanced by single packed instructions (VIMNMX3.S16x2.RELU, VIADDMNMX.S16x2, VIADD.16x2);
//   * thread t of the group owns query rows [t*K, (t+1)*K) of the current pass; H and E of those rows stay in
//     registers; database columns stream through the group as a systolic pipeline: at step s thread t works
//     on column s - t and hands (H, F) of its last row plus the column's profile offsets to thread t + 1
//     with warp shuffles;
//   * the pipeline never drains between sequences: when thread 0 has fed the last column of one segment
//     (a pass of a task) it feeds the first column of the next one; a "new segment" mark travels down the
//     pipeline with the column and each thread resets its rows when the mark reaches it.  A task's score is
//     reduced and stored a few loop trips later, when the mark of the following task has passed every thread;
//   * queries longer than G*K rows take several passes; the last row of a pass is parked in a per-warp
//     global scratch line (L2 resident) and re-enters at thread 0 in the next pass.
//
// Per cell pair the recurrence is 4.5 ALU-pipe instructions (PRMT score pack, VIMNMX3.RELU for H, two
// VIADDMNMX for E and F, half a VIMNMX3 for the running best) plus two VIADD.16x2, which issue on the
// FMA-heavy pipe and overlap the ALU pipe completely (pipebench: pair_viaddmnmx_viadd16x2 = 126/clk/SM).
//
// Exactness: lanes use wrapping 16-bit adds.  A lane whose running best reaches kOverflow16 is reported in
// resc_list and recomputed by the Lane32 instantiation, so every stored score is the exact int32 value, like
// the reference's 8 -> 16 -> 32 bit escalation (CPUsearch.c:678-956).
#pragma once

#include <type_traits>

#include "swg_common.cuh"

namespace swg {

// A column travels down the pipeline as one 32-bit word:
//   bits 15:8  = 4 * residue code of sequence A  (so that word & 0xff00 is the profile offset code * 1024)
//   bits 31:24 = 4 * residue code of sequence B
//   bits 23:16 = pass of the segment the column belongs to (selects the 32 KB profile slice)
//   bits  2:0  = marks
constexpr uint32_t kMarkSegment = 1u;   // first column of a segment: the rows restart from H = E = 0
constexpr uint32_t kMarkTask = 2u;      // ... and it is the first pass of a new task: the running best is parked
constexpr uint32_t kMarkCarry = 4u;     // the segment is not the task's last pass: the last row goes to the scratch line

// GOE/GE > 0: the gap penalties are compile-time constants and become immediate operands (fewer register-file
// reads per cell: pipebench mix_v2_immediate_penalties vs mix_v2); 0: taken from the parameter block.  Instantiated:
// (12, 2) = SWIMM's defaults, (0, 1) and (0, 2) = gap-extend 1 or 2 with any gap-open, (0, 0) = generic.
template <class L, int G, int K, bool MP, bool GP, int GOE, int GE>
__global__ void __launch_bounds__(kBlockThreads, 1) wavefront_kernel(const WfParams p)
{
    typedef typename L::reg reg;
    static_assert(G == 4 || G == 8 || G == 16 || G == 32, "group size");
    static_assert(K >= 1 && K <= kMaxRowsPerThread, "rows per thread");
    static_assert(L::kSeqs == 2 || G == 32, "the 32-bit kernel runs one sequence per warp");
    static_assert(!MP || G == 32, "several passes need one pair per warp (the scratch line is per warp)");
    static_assert(!GP || MP, "the global-profile variant is a multi-pass kernel");
    constexpr int GPW = 32 / G;                       // groups per warp
    constexpr int TPT = (L::kSeqs == 2) ? (kTilePairs / GPW) : 1;   // warp tasks per tile
    constexpr int KCH = (K + 15) / 16;                // 16-row profile chunks per thread
    constexpr int NC = kTripCols;                     // columns per loop trip
    constexpr uint32_t FI = G / NC;                   // trip of the NEXT task at which a finished task is stored
    constexpr uint32_t kMinTrips = FI + 4;            // every segment is at least this long (padding columns)
    static_assert(kMinTrips * NC <= kBoundarySlack, "scratch line too short for padded segments");

    // GP ("global profile"): queries with more passes than fit in shared memory read the profile through L1
    extern __shared__ __align__(16) uint8_t prof_smem[];
    const uint8_t *prof = GP ? p.profile : prof_smem;
    if (L::kSeqs == 1 && *p.resc_count == 0) return;      // nothing left the 16-bit range
    if (!GP) {
        const uint4 *src = reinterpret_cast<const uint4 *>(p.profile);
        uint4 *dst = reinterpret_cast<uint4 *>(prof_smem);
        const int n16 = (int)(p.passes * (kPassBytes / 16));
        for (int i = threadIdx.x; i < n16; i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();

    const int lane = threadIdx.x & 31;
    const int t = lane % G;
    const int g = lane / G;
    const uint32_t warp_global = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    uint2 *bnd = p.boundary + (size_t)warp_global * p.maxcols;
    // GE > 0: the gap-extend penalty as an immediate -- it is the third operand of the two VIADDMNMX per cell, which
    // with it in a register read three registers (two register-bank cycles); GOE > 0: go + ge as an immediate too
    const reg nge = GE > 0 ? L::splat(-GE) : L::splat(-p.gap_extend);
    const reg ngoe = GOE > 0 ? L::splat(-GOE) : L::splat(-p.gap_open_extend);
    const uint32_t npass = MP ? p.passes : 1u;
    const uint32_t pad_pk = (L::kSeqs == 2) ? 0x60006000u : 0x00006000u;

    uint32_t ntasks;
    if (L::kSeqs == 2) ntasks = (p.vt ? p.vt_count : p.tile_count) * TPT;
    else ntasks = *p.resc_count;

    auto fetch_task = [&]() -> uint32_t {
        uint32_t task = 0;
        if (lane == 0) task = atomicAdd(p.task_counter, 1u);
        return __shfl_sync(0xffffffffu, task, 0);
    };

    // ---- per-thread pipeline state ----
    // DS[x] = H(i-1, j) + S(i, j+1): the diagonal term of row x for the column this thread processes NEXT.
    // It is formed one column ahead (the profile words loaded in a step are those of the next column), which
    // makes every row slot a plain read-then-overwrite and keeps the hot loop free of register moves.
    reg DS[K], E[K];
#pragma unroll
    for (int x = 0; x < K; ++x) { DS[x] = L::splat(0); E[x] = L::splat(0); }
    reg best = L::splat(0), bsave = L::splat(0);
    reg out_h = L::splat(0), out_f = L::splat(0);
    uint32_t pk = pad_pk;             // the column word received last (= of the column processed in the next step)

    // One step of this thread: process the column whose word arrived in the previous step, using (H, F) of the
    // row above handed down now, and prepare DS for the column whose word arrives now.
    //
    // Two versions.  HEAD (the first G steps of a segment): threads may still be in the previous segment, so the
    // marks of every column word are looked at -- restart of the rows, profile slice of the word's pass, whether
    // the last row is carried.  Steady state (all later steps): every thread is inside the current segment, nothing
    // of that is looked at, the profile slice and the carry decision are per-segment values.
    uint32_t slice_off = (uint32_t)t * 16;     // byte offset of this thread's rows in the current pass's profile slice
    bool store_on = false;                     // steady state: this thread parks its last row (thread G-1, carried pass)
    auto column = [&](auto head_tag, uint32_t in_pkn, uint2 in_hf, uint32_t store_col) {
        constexpr bool HEAD = decltype(head_tag)::value;
        uint32_t pkn = __shfl_up_sync(0xffffffffu, pk, 1, G);
        reg r_h = (reg)__shfl_up_sync(0xffffffffu, out_h, 1, G);
        reg r_f = (reg)__shfl_up_sync(0xffffffffu, out_f, 1, G);
        if (t == 0) {
            pkn = in_pkn;
            r_h = MP ? (reg)in_hf.x : L::splat(0);
            r_f = MP ? (reg)in_hf.y : L::splat(0);
        }
        const uint32_t pk_cur = pk;
        pk = pkn;

        // profile words of the NEXT column
        uint32_t w1[KCH * 4], w2[KCH * 4];
        {
            const uint32_t hi = pkn >> 16;
            if (MP && HEAD) slice_off = (uint32_t)t * 16 + (hi & 0xffu) * kPassBytes;
            // pass (bits 15..), letter (bits 10..14) and thread (bits 4..9) offsets do not overlap: one logic op each
            const uint32_t oa = (pkn & 0xff00u) | slice_off;
            const uint32_t ob = (hi & 0xff00u) | slice_off;
            const uint4 *q1 = reinterpret_cast<const uint4 *>(prof + oa);
#pragma unroll
            for (int i = 0; i < KCH; ++i) {
                const uint4 v = q1[i * G];
                w1[4 * i] = v.x; w1[4 * i + 1] = v.y; w1[4 * i + 2] = v.z; w1[4 * i + 3] = v.w;
            }
            if (L::kSeqs == 2) {
                const uint4 *q2 = reinterpret_cast<const uint4 *>(prof + ob);
#pragma unroll
                for (int i = 0; i < KCH; ++i) {
                    const uint4 v = q2[i * G];
                    w2[4 * i] = v.x; w2[4 * i + 1] = v.y; w2[4 * i + 2] = v.z; w2[4 * i + 3] = v.w;
                }
            } else {
#pragma unroll
                for (int i = 0; i < KCH * 4; ++i) w2[i] = 0u;
            }
        }
        auto score_of = [&](int x) -> reg {
            switch (x & 3) {
                case 0: return L::template score<0>(w1[x >> 2], w2[x >> 2]);
                case 1: return L::template score<1>(w1[x >> 2], w2[x >> 2]);
                case 2: return L::template score<2>(w1[x >> 2], w2[x >> 2]);
                default: return L::template score<3>(w1[x >> 2], w2[x >> 2]);
            }
        };

        // One cell per row: 4.5 ALU-pipe instructions (PRMT, VIMNMX3.RELU, 2 x VIADDMNMX, VIMNMX3 / 2) and two
        // VIADD.16x2 on the FMA-heavy pipe.  H = max(ds, E, F, 0) is ONE instruction because ds is formed by a
        // separate add; the running best is taken over ds instead of H -- the same maximum, since E and F only
        // ever hold earlier H values minus gap penalties.
        reg hp = r_h, f = r_f, dsprev = L::splat(0);
#pragma unroll
        for (int x = 0; x < K; ++x) {
            const reg ds = DS[x];
            const reg h = L::max3_relu(ds, E[x], f);       // max(ds, E(i,j), F(i,j), 0)
            const reg open = L::add(h, ngoe);              // H(i,j) - (go+ge)
            E[x] = L::addmax(E[x], nge, open);             // E(i,j+1)
            f = L::addmax(f, nge, open);                   // F(i+1,j)
            if (x & 1) best = L::max3(best, dsprev, ds);
            else if (x == K - 1) best = L::max2(best, ds);
            dsprev = ds;
            DS[x] = L::add(hp, score_of(x));               // H(i-1,j) + S(i,j+1)
            hp = h;
        }
        out_h = hp;
        out_f = f;
        if (MP && (HEAD ? (t == G - 1 && (pk_cur & kMarkCarry)) : store_on))
            bnd[store_col] = make_uint2((uint32_t)out_h, (uint32_t)out_f);
        if (HEAD && (pkn & kMarkSegment)) {          // rare and divergent: the next column starts a segment
#pragma unroll
            for (int x = 0; x < K; ++x) { DS[x] = score_of(x); E[x] = L::splat(0); }
            if (pkn & kMarkTask) { bsave = best; best = L::splat(0); }
        }
    };
    const std::true_type head_steps;
    const std::false_type steady_steps;

    // Store the scores of a task whose last column has left the pipeline (all threads hold its best in bsave).
    auto finalize = [&](uint32_t lseq) {
        reg b = bsave;
#pragma unroll
        for (int o = G >> 1; o > 0; o >>= 1) b = L::max2(b, (reg)__shfl_xor_sync(0xffffffffu, b, o, G));
        if (t == 0) {
            if (L::kSeqs == 2) {
                const uint32_t bb = (uint32_t)b;
                const int s_lo = (int)(short)(bb & 0xffffu), s_hi = (int)(short)(bb >> 16);
                if (p.vt) {
                    // a column chunk of a long sequence: the sequence's score is the best of its chunks; the chunk
                    // that first lifts it to kOverflow16 lists it
                    if (global_seq_index(p, lseq) < p.n_total) {
                        const int old = atomicMax(p.scores + lseq, s_lo);
                        if (s_lo >= kOverflow16 && old < kOverflow16) p.resc_list[atomicAdd(p.resc_count, 1u)] = lseq;
                    }
                    if (global_seq_index(p, lseq + 1) < p.n_total) {
                        const int old = atomicMax(p.scores + lseq + 1, s_hi);
                        if (s_hi >= kOverflow16 && old < kOverflow16) p.resc_list[atomicAdd(p.resc_count, 1u)] = lseq + 1;
                    }
                    return;
                }
                if (global_seq_index(p, lseq) < p.n_total) {
                    p.scores[lseq] = s_lo;
                    if (s_lo >= kOverflow16) p.resc_list[atomicAdd(p.resc_count, 1u)] = lseq;
                }
                if (global_seq_index(p, lseq + 1) < p.n_total) {
                    p.scores[lseq + 1] = s_hi;
                    if (s_hi >= kOverflow16) p.resc_list[atomicAdd(p.resc_count, 1u)] = lseq + 1;
                }
            } else {
                p.scores[lseq] = (int32_t)b;
            }
        }
    };

    bool pending = false;             // warp-uniform: a finished task waits for its store
    uint32_t pend_lseq = 0;
    uint32_t next_task = fetch_task();
    uint2 hf_in = make_uint2(0u, 0u);   // thread 0: (H, F) entering the column it processes in the next step

    for (;;) {
        const uint32_t task = next_task;
        const bool have = task < ntasks;
        if (!have && !pending) break;
        // The next task is drawn 16 trips (64 columns: far more than the atomic's latency) before this one ends, during
        // its last pass -- not at its start: a warp that reserves its next task while it begins a long one commits to
        // two long tasks in a row, which breaks the longest-first balance when there are only a few tasks per warp.

        // ---- decode the task (a flush segment of padding columns when there is none left) ----
        uint32_t half = 0, lseq = 0, ncols = 0;
        const uint2 *words = nullptr;
        if (have) {
            uint32_t tile, pair, col0 = 0;
            if (L::kSeqs == 2) {
                pair = (task % TPT) * GPW + g;
                if (p.vt) {
                    const uint4 v = p.vt[task / TPT];                        // a column chunk of a long tile
                    tile = v.x;
                    col0 = v.y;
                    ncols = v.z;
                } else {
                    tile = p.tile_first + p.tile_count - 1 - task / TPT;    // longest tiles first
                    ncols = p.tile_cols[tile];
                }
                lseq = tile * kTileSeqs + 2 * pair;
            } else {
                lseq = p.resc_list[task];
                tile = lseq / kTileSeqs;
                pair = (lseq % kTileSeqs) >> 1;
                half = lseq & 1u;
                ncols = p.tile_cols[tile];
            }
            words = reinterpret_cast<const uint2 *>(p.db + p.tile_off[tile] + (size_t)(col0 / kChunkCols) * kTilePairs + pair);
        }
        const uint32_t data_trips = ncols / NC;
        const uint32_t trips = data_trips < kMinTrips ? kMinTrips : data_trips;
        const uint32_t seg_cols = trips * NC;
        // PRMT selectors that build a column word: residue bytes (code*4) of the pair go to byte lanes 1 and 3, byte
        // lanes 0 and 2 (marks, pass) are taken from the second operand -- tagging costs no instruction
        const uint32_t selA = (L::kSeqs == 2) ? 0x1604u : (0x7604u | (half << 4));
        const uint32_t selB = (L::kSeqs == 2) ? 0x3624u : (0x7604u | ((2u + half) << 4));
        const uint32_t seg_passes = have ? npass : 1u;

        for (uint32_t pass = 0; pass < seg_passes; ++pass) {
            const uint32_t first_marks = kMarkSegment | (pass == 0 ? kMarkTask : 0u);
            const uint32_t tag = MP ? ((pass << 16) | ((pass + 1 < seg_passes) ? kMarkCarry : 0u)) : 0u;
            const bool carry_in = MP && pass > 0;

            // thread 0 feeds the pipeline: four columns of the pair per trip (two 32-bit words), fetched one trip
            // ahead, and -- after the first pass -- the (H, F) row parked by the previous pass, four columns ahead
            // (a ring of four register pairs: an L2 round trip is longer than one column of a short-row shape)
            uint2 w = make_uint2(kPadWord, kPadWord);
            if (t == 0 && data_trips > 0) w = words[0];
            uint2 ring[NC];                   // ring[j] enters the column processed in step j of the current trip
#pragma unroll
            for (int j = 0; j < NC; ++j) ring[j] = make_uint2(0u, 0u);
            if (MP) {
                ring[0] = hf_in;              // the last column of the previous segment is processed in step 0
                if (carry_in) {
#pragma unroll
                    for (int j = 1; j < NC; ++j) ring[j] = __ldcg(bnd + j - 1);
                }
            }
            // One trip = four steps.  The words of columns 4*trip .. 4*trip+3 enter; the cells processed are those of
            // columns 4*trip-1 .. 4*trip+2.  Thread G-1 runs G-1 columns behind thread 0: during the head trips its
            // scratch-line column may still lie in the previous pass (same task, same length).
            auto trip_body = [&](auto head_tag, uint32_t trip) {
                uint2 nw = make_uint2(kPadWord, kPadWord);
                const uint32_t nt = trip + 1;
                if (t == 0 && nt < data_trips) nw = words[(nt >> 1) * (kTilePairs * 2) + (nt & 1u)];
                const uint32_t c0 = trip * NC;
                uint32_t sc = c0 + seg_cols - G;
                if (sc >= seg_cols) sc -= seg_cols;         // a multiple of 4: the four columns of a trip never wrap
#pragma unroll
                for (int j = 0; j < NC; ++j) {
                    const uint32_t word = (j < 2) ? w.x : w.y;
                    const bool first = decltype(head_tag)::value && j == 0 && trip == 0;
                    const uint32_t pkn = prmt(word, first ? (tag | first_marks) : tag, (j & 1) ? selB : selA);
                    const uint2 hf = ring[j];         // (H, F) entering column c0 + j - 1, processed in this step
                    // the column processed in step j of the next trip (the line has slack past the segment's end:
                    // entries read there are never used).  In pass 0 nothing is carried in: the slot is cleared during
                    // the head trips (slot 0 held the previous segment's last column, used once in step 0) and stays zero
                    if (MP && pass > 0) ring[j] = __ldcg(bnd + c0 + j + (NC - 1));
                    else if (MP && decltype(head_tag)::value) ring[j] = make_uint2(0u, 0u);
                    column(head_tag, pkn, hf, sc + j);
                }
                w = nw;
                if (MP) __syncwarp();             // orders thread G-1's scratch-line stores before thread 0's later loads
            };
#pragma unroll 1
            for (uint32_t trip = 0; trip < FI; ++trip) trip_body(head_steps, trip);
            if (pending && pass == 0) { finalize(pend_lseq); pending = false; }
            store_on = MP && t == G - 1 && (pass + 1 < seg_passes);
            const uint32_t fetch_trip = (have && pass + 1 == seg_passes) ? (trips > FI + 16 ? trips - 16 : FI) : 0xffffffffu;
#pragma unroll 1
            for (uint32_t trip = FI; trip < trips; ++trip) {
                if (trip == fetch_trip) next_task = fetch_task();
                trip_body(steady_steps, trip);
            }
            if (MP) hf_in = ring[0];          // enters the segment's last column, processed in the next segment's step 0
        }
        if (have) { pending = true; pend_lseq = lseq; }
    }
}

// host-side launchers; one translation unit per (lane policy, group size) family (wavefront_inst_*.cu)
cudaError_t launch_wf_l16_g4(int K, int grid, size_t smem, cudaStream_t stream, const WfParams &p);
cudaError_t launch_wf_l16_g8(int K, int grid, size_t smem, cudaStream_t stream, const WfParams &p);
cudaError_t launch_wf_l16_g16(int K, int grid, size_t smem, cudaStream_t stream, const WfParams &p);
cudaError_t launch_wf_l16_g32(int K, int grid, size_t smem, cudaStream_t stream, const WfParams &p);
cudaError_t launch_wf_l32_g32(int K, int grid, size_t smem, cudaStream_t stream, const WfParams &p);
// multi-pass variants (G = 32): the last row of a pass is carried to the next one through the scratch line
cudaError_t launch_wf_l16_g32_mp(int K, int grid, size_t smem, cudaStream_t stream, const WfParams &p);
cudaError_t launch_wf_l32_g32_mp(int K, int grid, size_t smem, cudaStream_t stream, const WfParams &p);
// the two "profile in global memory" variants (G = 32, K = 32) for queries of more than kMaxSmemPasses passes
cudaError_t launch_wf_l16_gp(int grid, cudaStream_t stream, const WfParams &p);
cudaError_t launch_wf_l32_gp(int grid, cudaStream_t stream, const WfParams &p);

template <class L, int G, int K, bool MP, bool GP, int GOE = 0, int GE = 0>
cudaError_t launch_one(int grid, size_t smem, cudaStream_t stream, const WfParams &p)
{
    cudaError_t e = ensure_dynamic_smem(reinterpret_cast<const void *>(&wavefront_kernel<L, G, K, MP, GP, GOE, GE>), smem);
    if (e != cudaSuccess) return e;
    wavefront_kernel<L, G, K, MP, GP, GOE, GE><<<grid, kBlockThreads, smem, stream>>>(p);
    return cudaGetLastError();
}

}  // namespace swg
