// wavefront_xw.cuh -- K3, the long-sequence kernel: the anti-diagonal wavefront of ONE database sequence (pair) runs
// ACROSS THE WARPS of a CTA.
//
// In the search kernels of wavefront.cuh / wavefront_q2.cuh a sequence is owned by one thread group: its columns are
// a serial chain of (columns x passes) steps, and on a database with a few very long sequences (the format allows
// 65535 residues) that chain outlasts the whole rest of the search.  The reference has no counterpart: it only blocks
// the columns of a group of 32 sequences and carries the last column / the E row between blocks
// (CPUsearch.c:562-579, 657-662).
//
// Here the query's rows are tiled into W passes of 32*K rows and the W passes of a sequence run CONCURRENTLY, one per
// warp: warp w of a group owns rows [w*32*K, (w+1)*32*K); inside the warp the columns stream through the 32 threads
// as in wavefront.cuh (systolic pipeline, H diagonal term and E in registers, DS one column ahead).  The last row
// (H, F) of warp w's pass leaves its thread 31 one column at a time into a 128-entry ring in shared memory and
// enters warp w+1's thread 0 forty steps later: the tiles (pass, 4-column trip) of one sequence form a wavefront over
// the warps, 40 columns apart.  A sequence of n columns therefore takes n + 40*W steps instead of n * W, and a group
// is busy on a sequence for 1/W of the time -- the database's longest sequence stops bounding the launch.
//
// Hand-off protocol (all inside one CTA): every warp counts its steps; after each trip (4 columns) thread 31, which
// wrote the trip's last-row entries, publishes the warp's step count in shared memory with a release store.  A warp
// loads entry p of its input ring only after it has seen (acquire) a count > p + 32 from the warp above; a warp
// overwrites ring entry p + R only after the warp below has published a count that shows entry p consumed.  The task
// list is shared the same way: the group's first warp draws (pair) tasks from the global counter and posts them in a
// small ring for the other warps, so that all W warps walk the same tasks in the same order -- longest tiles first.
//
// The query's byte profile of all W passes is resident in shared memory (W * 25 letters * 512 or 1024 bytes), laid
// out as in wavefront.cuh; queries of up to 8192 rows fit (16 warps x 32 threads x 16 rows, or 8 x 32 x 32).
//
// The Lane32 instantiation is the exact 32-bit recomputation (K2) for the same shapes: a flagged long sequence no
// longer costs columns x passes steps of a single warp.
//
// Scores: every warp holds the best of its own rows; the W partial results of a sequence are merged with atomicMax
// (the host zeroes the 16-bit launch's score range first); the warp whose contribution first lifts a score to
// kOverflow16 lists the sequence for the 32-bit recomputation.
#pragma once

#include <type_traits>

#include "swg_common.cuh"
#include "wavefront.cuh"

namespace swg {

constexpr int kXwTaskRing = 32;            // tasks the first warp of a group may be ahead of the last one
constexpr uint32_t kXwLag = 40;            // steps a warp stays behind the warp above it (32 threads + 2 trips)
constexpr uint32_t kXwRing = 128;          // entries of a warp's last-row ring (shared memory, 8 bytes each)
constexpr uint32_t kXwRingGuard = 8;       // margin of the overwrite test (an entry is dead 34 steps before the test needs it)
constexpr uint32_t kXwDone = 0xffffffffu;  // published step count of a warp that has left the kernel
constexpr int kXwMaxRows = 8192;           // longest query the shapes cover

__device__ __forceinline__ uint32_t ld_acquire_shared(const uint32_t *ptr)
{
    uint32_t v;
    asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(v) : "r"((uint32_t)__cvta_generic_to_shared(ptr)) : "memory");
    return v;
}

__device__ __forceinline__ void st_release_shared(uint32_t *ptr, uint32_t v)
{
    asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(ptr)), "r"(v) : "memory");
}

// spin (whole warp, same address) until the published count reaches `need`; kXwDone satisfies every need
__device__ __forceinline__ uint32_t wait_at_least(const uint32_t *flag, uint32_t need)
{
    uint32_t v = ld_acquire_shared(flag);
    while (v < need) {
        __nanosleep(64);
        v = ld_acquire_shared(flag);
    }
    return v;
}

// W = p.xw_warps warps per sequence (pair), 16 / W groups per CTA; K rows per thread; GOE/GE as in wavefront.cuh.
template <class L, int K, int GOE, int GE>
__global__ void __launch_bounds__(kBlockThreads, 1) wavefront_xw_kernel(const WfParams p)
{
    typedef typename L::reg reg;
    static_assert(K >= 1 && K <= kMaxRowsPerThread, "rows per thread");
    static_assert(kBlockThreads % 32 == 0, "whole warps");      // (the planner assumes 16 warps per CTA: experimental CTA sizes do not run this kernel)
    constexpr int G = 32;
    constexpr int TPT = kTilePairs;                   // Lane16: one pair per group and task
    constexpr int KCH = (K + 15) / 16;                // 16-row profile chunks per thread
    constexpr int NC = kTripCols;
    constexpr uint32_t FI = G / NC;                   // head trips of a segment
    constexpr uint32_t kMinTrips = FI + 4;
    constexpr uint32_t LS = KCH * 512;                // bytes of a letter row in shared memory (32 threads x 16 rows per chunk)
    constexpr uint32_t SLICE = kLetters * LS;         // one pass

    extern __shared__ __align__(16) uint8_t prof_smem[];
    __shared__ uint32_t s_prog[16];                   // steps completed by each warp
    __shared__ uint32_t s_task[16][kXwTaskRing];      // tasks posted by each group's first warp
    __shared__ uint32_t s_task_pub[16], s_task_taken[16];

    if (L::kSeqs == 1 && *p.resc_count == 0) return;      // nothing left the 16-bit range
    const uint32_t W = p.xw_warps;
    {
        // compact copy of the [pass][letter][1024] profile: the first LS bytes of every letter row
        const uint4 *src = reinterpret_cast<const uint4 *>(p.profile);
        uint4 *dst = reinterpret_cast<uint4 *>(prof_smem);
        constexpr uint32_t row16 = LS / 16;
        const uint32_t n16 = W * kLetters * row16;
        for (uint32_t i = threadIdx.x; i < n16; i += blockDim.x) {
            const uint32_t o = i % row16, letter = (i / row16) % kLetters, pass = i / (row16 * kLetters);
            dst[i] = src[pass * (kPassBytes / 16) + letter * (kLetterStride / 16) + o];
        }
        if (threadIdx.x < 16) { s_prog[threadIdx.x] = 0; s_task_pub[threadIdx.x] = 0; s_task_taken[threadIdx.x] = 0; }
    }
    __syncthreads();

    const int lane = threadIdx.x & 31;
    const int t = lane;
    const uint32_t warp = threadIdx.x >> 5;
    const uint32_t w = warp % W;                      // this warp's pass
    const uint32_t grp = warp / W;
    if (grp >= p.xw_groups) return;                   // fewer pairs per CTA = fewer warps per scheduler = shorter steps
    const bool has_in = w > 0, has_out = w + 1 < W;
    // the last-row rings live behind the profile: no long-latency load is ever outstanding when a step count is
    // published (the release store's fence would wait for it)
    constexpr uint32_t R = kXwRing, rmask = R - 1;
    uint2 *const ring_out = reinterpret_cast<uint2 *>(prof_smem + W * SLICE) + warp * R;
    const uint2 *const ring_in = ring_out - R;        // the ring of the warp above (has_in)
    const uint32_t *const flag_prev = &s_prog[warp - (has_in ? 1 : 0)];
    const uint32_t *const flag_next = &s_prog[warp + (has_out ? 1 : 0)];
    const reg nge = GE > 0 ? L::splat(-GE) : L::splat(-p.gap_extend);              // (immediates: see wavefront.cuh)
    const reg ngoe = GOE > 0 ? L::splat(-GOE) : L::splat(-p.gap_open_extend);
    const uint32_t pad_pk = (L::kSeqs == 2) ? 0x60006000u : 0x00006000u;
    const uint8_t *const prof = prof_smem + w * SLICE + (uint32_t)t * 16;

    uint32_t ntasks;
    if (L::kSeqs == 2) ntasks = p.tile_count * TPT;
    else ntasks = *p.resc_count;

    // ---- the group's task list ----
    uint32_t tn = 0;                                  // tasks this warp has taken
    auto get_task = [&]() -> uint32_t {
        uint32_t task = 0;
        if (w == 0) {
            if (lane == 0) task = atomicAdd(p.task_counter, 1u);
            task = __shfl_sync(0xffffffffu, task, 0);
            if (W > 1) {
                if (tn >= (uint32_t)kXwTaskRing) wait_at_least(&s_task_taken[grp], tn - kXwTaskRing + 1);
                if (lane == 0) {
                    s_task[grp][tn % kXwTaskRing] = task;
                    st_release_shared(&s_task_pub[grp], tn + 1);
                }
            }
        } else {
            wait_at_least(&s_task_pub[grp], tn + 1);
            task = reinterpret_cast<volatile uint32_t *>(s_task[grp])[tn % kXwTaskRing];
            if (!has_out) {
                __syncwarp();
                if (lane == 0) st_release_shared(&s_task_taken[grp], tn + 1);
            }
        }
        ++tn;
        return task;
    };

    // ---- per-thread pipeline state (see wavefront.cuh) ----
    reg DS[K], E[K];
#pragma unroll
    for (int x = 0; x < K; ++x) { DS[x] = L::splat(0); E[x] = L::splat(0); }
    reg best = L::splat(0), bsave = L::splat(0);
    reg out_h = L::splat(0), out_f = L::splat(0);
    uint32_t pk = pad_pk;
    const bool store_on = has_out && t == G - 1;

    auto column = [&](auto head_tag, uint32_t in_pkn, uint2 in_hf, uint32_t store_at) {
        constexpr bool HEAD = decltype(head_tag)::value;
        uint32_t pkn = __shfl_up_sync(0xffffffffu, pk, 1, G);
        reg r_h = (reg)__shfl_up_sync(0xffffffffu, out_h, 1, G);
        reg r_f = (reg)__shfl_up_sync(0xffffffffu, out_f, 1, G);
        if (t == 0) {
            pkn = in_pkn;
            r_h = (reg)in_hf.x;
            r_f = (reg)in_hf.y;
        }
        pk = pkn;

        // profile words of the NEXT column
        uint32_t w1[KCH * 4], w2[KCH * 4];
        {
            const uint32_t hi = pkn >> 16;
            const uint32_t oa = KCH == 2 ? (pkn & 0xff00u) : ((pkn & 0xff00u) >> 1);
            const uint32_t ob = KCH == 2 ? (hi & 0xff00u) : ((hi & 0xff00u) >> 1);
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
        if (store_on) ring_out[store_at] = make_uint2((uint32_t)out_h, (uint32_t)out_f);
        if (HEAD && (pkn & kMarkSegment)) {          // the next column starts a new task
#pragma unroll
            for (int x = 0; x < K; ++x) { DS[x] = score_of(x); E[x] = L::splat(0); }
            bsave = best;
            best = L::splat(0);
        }
    };
    const std::true_type head_steps;
    const std::false_type steady_steps;

    // Merge this warp's part of a finished task's scores (all threads hold it in bsave).
    auto emit = [&](uint32_t ls, int s) {
        if (global_seq_index(p, ls) >= p.n_total) return;
        const int old = atomicMax(p.scores + ls, s);
        if (L::kSeqs == 2 && s >= kOverflow16 && old < kOverflow16) p.resc_list[atomicAdd(p.resc_count, 1u)] = ls;
    };
    auto finalize = [&](uint32_t lseq) {
        reg b = bsave;
#pragma unroll
        for (int o = G >> 1; o > 0; o >>= 1) b = L::max2(b, (reg)__shfl_xor_sync(0xffffffffu, b, o, G));
        if (t == 0) {
            if (L::kSeqs == 2) {
                const uint32_t bb = (uint32_t)b;
                emit(lseq, (int)(short)(bb & 0xffffu));
                emit(lseq + 1, (int)(short)(bb >> 16));
            } else {
                emit(lseq, (int)b);
            }
        }
    };

    // ---- the stream of columns ----
    // `base` = steps this warp has completed before the current segment.  In step S thread 0 processes stream column
    // S - 1 (its word arrived in step S - 1), thread 31 processes and parks column S - 32, and the (H, F) entering
    // column S + 3 is loaded into a queue of four register pairs.
    uint32_t base = 0, prod_seen = 0, cons_seen = 0;
    uint2 ring[NC];
#pragma unroll
    for (int j = 0; j < NC; ++j) ring[j] = make_uint2(0u, 0u);
    if (has_in) {
        prod_seen = wait_at_least(flag_prev, kXwLag - NC);
#pragma unroll
        for (int j = 1; j < NC; ++j) ring[j] = ring_in[j - 1];
    }
    bool pending = false;
    uint32_t pend_lseq = 0;
    uint32_t next_task = get_task();

    for (;;) {
        const uint32_t task = next_task;
        const bool have = task < ntasks;
        if (!have && !pending) break;
        // (the next task is drawn 16 trips before this one ends, not at its start: see wavefront.cuh)

        uint32_t half = 0, lseq = 0, ncols = 0;
        const uint2 *words = nullptr;
        if (have) {
            uint32_t tile, pair;
            if (L::kSeqs == 2) {
                tile = p.tile_first + p.tile_count - 1 - task / TPT;        // longest tiles first
                pair = task % TPT;
                lseq = tile * kTileSeqs + 2 * pair;
            } else {
                lseq = p.resc_list[task];
                tile = lseq / kTileSeqs;
                pair = (lseq % kTileSeqs) >> 1;
                half = lseq & 1u;
            }
            ncols = p.tile_cols[tile];
            words = reinterpret_cast<const uint2 *>(p.db + p.tile_off[tile] + pair);
        }
        const uint32_t data_trips = ncols / NC;
        const uint32_t trips = data_trips < kMinTrips ? kMinTrips : data_trips;
        const uint32_t selA = (L::kSeqs == 2) ? 0x1604u : (0x7604u | (half << 4));
        const uint32_t selB = (L::kSeqs == 2) ? 0x3624u : (0x7604u | ((2u + half) << 4));

        uint2 wd = make_uint2(kPadWord, kPadWord);
        if (t == 0 && data_trips > 0) wd = words[0];
        auto trip_body = [&](auto head_tag, uint32_t trip) {
            const uint32_t S0 = base + trip * NC;
            if (has_in && prod_seen < S0 + kXwLag) prod_seen = wait_at_least(flag_prev, S0 + kXwLag);
            if (has_out && S0 > R - kXwRingGuard && cons_seen < S0 - (R - kXwRingGuard))
                cons_seen = wait_at_least(flag_next, S0 - (R - kXwRingGuard));
            uint2 nw = make_uint2(kPadWord, kPadWord);
            const uint32_t nt = trip + 1;
            if (t == 0 && nt < data_trips) nw = words[(nt >> 1) * (kTilePairs * 2) + (nt & 1u)];
#pragma unroll
            for (int j = 0; j < NC; ++j) {
                const uint32_t word = (j < 2) ? wd.x : wd.y;
                const bool first = decltype(head_tag)::value && j == 0 && trip == 0;
                const uint32_t pkn = prmt(word, first ? kMarkSegment : 0u, (j & 1) ? selB : selA);
                const uint2 hf = ring[j];
                if (has_in) ring[j] = ring_in[(S0 + j + NC - 1) & rmask];
                column(head_tag, pkn, hf, (S0 + j - G) & rmask);
            }
            wd = nw;
            // the trip's last-row entries are written: publish the step count (thread 31 wrote them)
            if (lane == G - 1) st_release_shared(&s_prog[warp], S0 + NC);
        };
#pragma unroll 1
        for (uint32_t trip = 0; trip < FI; ++trip) trip_body(head_steps, trip);
        if (pending) { finalize(pend_lseq); pending = false; }
        const uint32_t fetch_trip = have ? (trips > FI + 16 ? trips - 16 : FI) : 0xffffffffu;
#pragma unroll 1
        for (uint32_t trip = FI; trip < trips; ++trip) {
            if (trip == fetch_trip) next_task = get_task();
            trip_body(steady_steps, trip);
        }
        base += trips * NC;
        if (have) { pending = true; pend_lseq = lseq; }
    }
    if (lane == G - 1) st_release_shared(&s_prog[warp], kXwDone);
}

// host-side launchers (wavefront_xw_inst_*.cu): K = 1..32, W = p.xw_warps in {1, 2, 4, 8, 16}
cudaError_t launch_xw_l16(int K, int grid, cudaStream_t stream, const WfParams &p);
cudaError_t launch_xw_l32(int K, int grid, cudaStream_t stream, const WfParams &p);

// dynamic shared memory of a shape: the compact profile of W passes + one last-row ring per warp
inline size_t xw_smem_bytes(int K, int W)
{
    return (size_t)W * kLetters * (size_t)((K + 15) / 16) * 512 + 16 * (size_t)kXwRing * sizeof(uint2);
}

template <class L, int K, int GOE = 0, int GE = 0>
cudaError_t launch_xw_one(int grid, cudaStream_t stream, const WfParams &p)
{
    const size_t smem = xw_smem_bytes(K, (int)p.xw_warps);
    cudaError_t e = ensure_dynamic_smem(reinterpret_cast<const void *>(&wavefront_xw_kernel<L, K, GOE, GE>), smem);
    if (e != cudaSuccess) return e;
    wavefront_xw_kernel<L, K, GOE, GE><<<grid, kBlockThreads, smem, stream>>>(p);
    return cudaGetLastError();
}

}  // namespace swg
