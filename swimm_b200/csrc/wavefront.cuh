// wavefront.cuh -- the Smith-Waterman search kernel of the B200 build (replaces the reference's
// cpu_search_avx2_sp inner loops, CPUsearch.c:553-956).
//
// Parallelisation (inter-task across groups, wavefront inside a group):
//   * a GROUP of G threads (G = 4, 8, 16 or 32 lanes of one warp) aligns the query against one PAIR of
//     database sequences; the two sequences live in the two 16-bit halves of every register (Lane16) and
//     are advanced by single DPX instructions (VIADDMNMX.S16x2, VIMNMX.S16x2.RELU, VIMNMX3.S16x2);
//   * thread t of the group owns query rows [t*K, (t+1)*K) of the current pass, H and E of those rows
//     stay in registers; database columns stream through the group as a systolic pipeline: at step s
//     thread t works on column s - t and hands (H, F) of its last row plus the column's profile
//     offsets to thread t + 1 with warp shuffles;
//   * queries longer than G*K rows take several passes; the last row of a pass is parked in a
//     per-warp global scratch line (L2 resident) and re-enters at thread 0 of the next pass.
// Per cell pair the recurrence costs 6 integer-pipe instructions plus one PRMT that packs the two
// substitution scores fetched from the shared-memory query profile.
//
// Exactness: lanes use wrapping 16-bit adds.  A lane whose running best reaches kOverflow16 is
// reported in resc_list and recomputed by the Lane32 instantiation, so every stored score is the
// exact int32 value, like the reference's 8 -> 16 -> 32 bit escalation (CPUsearch.c:678-956).
#pragma once

#include "swg_common.cuh"

namespace swg {

template <class L, int G, int K, bool GP>
__global__ void __launch_bounds__(kBlockThreads, 1) wavefront_kernel(const WfParams p)
{
    typedef typename L::reg reg;
    static_assert(G == 4 || G == 8 || G == 16 || G == 32, "group size");
    static_assert(K >= 1 && K <= kMaxRowsPerThread, "rows per thread");
    static_assert(L::kSeqs == 2 || G == 32, "the 32-bit kernel runs one sequence per warp");
    constexpr int GPW = 32 / G;                       // groups per warp
    constexpr int TPT = (L::kSeqs == 2) ? (kTilePairs / GPW) : 1;   // warp tasks per tile
    constexpr int KCH = (K + 15) / 16;                // 16-row profile chunks per thread

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
    const reg nge = L::splat(-p.gap_extend);
    const reg ngoe = L::splat(-p.gap_open_extend);

    uint32_t ntasks;
    if (L::kSeqs == 2) ntasks = p.tile_count * TPT;
    else ntasks = *p.resc_count;

    for (;;) {
        uint32_t task = 0;
        if (lane == 0) task = atomicAdd(p.task_counter, 1u);
        task = __shfl_sync(0xffffffffu, task, 0);
        if (task >= ntasks) break;

        uint32_t tile, pair, half = 0, lseq;
        if (L::kSeqs == 2) {
            tile = p.tile_first + p.tile_count - 1 - task / TPT;        // longest tiles first
            pair = (task % TPT) * GPW + g;
            lseq = tile * kTileSeqs + 2 * pair;
        } else {
            lseq = p.resc_list[task];
            tile = lseq / kTileSeqs;
            pair = (lseq % kTileSeqs) >> 1;
            half = lseq & 1u;
        }
        const uint32_t ncols = p.tile_cols[tile];
        const uint32_t *words = reinterpret_cast<const uint32_t *>(p.db + p.tile_off[tile] + pair);
        const uint32_t niter = (ncols + G) >> 1;        // two columns per iteration, ncols + G - 1 steps
        // PRMT selectors that turn a residue byte (code*4) into the profile offset code*1024
        const uint32_t selA = (L::kSeqs == 2) ? 0x1404u : (0x4404u | (half << 4));
        const uint32_t selB = (L::kSeqs == 2) ? 0x3424u : (0x4404u | ((2u + half) << 4));
        const uint32_t pad_pk = (L::kSeqs == 2) ? 0x60006000u : 0x00006000u;

        reg best = L::splat(0);

        for (uint32_t pass = 0; pass < p.passes; ++pass) {
            reg H[K], E[K];
#pragma unroll
            for (int x = 0; x < K; ++x) { H[x] = L::splat(0); E[x] = L::splat(0); }
            reg out_h = L::splat(0), out_f = L::splat(0), up_prev = L::splat(0);
            uint32_t pk = pad_pk;
            const uint8_t *pbase = prof + pass * kPassBytes + t * 16;
            const bool carry_in = pass > 0;
            const bool carry_out = (pass + 1 < p.passes) && (t == G - 1);

            // thread 0's inputs for columns (0, 1)
            uint32_t word = words[0];
            uint2 b0 = make_uint2(0u, 0u), b1 = make_uint2(0u, 0u);
            if (carry_in && t == 0) { b0 = __ldcg(bnd); b1 = __ldcg(bnd + 1); }

            auto column = [&](uint32_t in_pk, reg in_h, reg in_f, int col_out) {
                // receive from the thread above (it finished this column one step ago)
                uint32_t r_pk = __shfl_up_sync(0xffffffffu, pk, 1, G);
                reg r_h = (reg)__shfl_up_sync(0xffffffffu, out_h, 1, G);
                reg r_f = (reg)__shfl_up_sync(0xffffffffu, out_f, 1, G);
                if (t == 0) { r_pk = in_pk; r_h = in_h; r_f = in_f; }
                pk = r_pk;
                reg diag = up_prev;
                up_prev = r_h;
                reg f = r_f;

                uint32_t w1[KCH * 4], w2[KCH * 4];
                {
                    const uint4 *q1 = reinterpret_cast<const uint4 *>(pbase + (pk & 0xffffu));
#pragma unroll
                    for (int i = 0; i < KCH; ++i) {
                        const uint4 v = q1[i * G];
                        w1[4 * i] = v.x; w1[4 * i + 1] = v.y; w1[4 * i + 2] = v.z; w1[4 * i + 3] = v.w;
                    }
                    if (L::kSeqs == 2) {
                        const uint4 *q2 = reinterpret_cast<const uint4 *>(pbase + (pk >> 16));
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

                // One cell per row.  ds = H(i-1,j-1) + S is a plain packed add (VIADD.16x2 does not occupy the
                // ALU pipe) so that H = max(ds, E, F, 0) is ONE ALU instruction (VIMNMX3.RELU); the running best
                // is taken over ds instead of H -- the same maximum, because E and F only ever hold earlier H values
                // minus gap penalties -- which also keeps ptxas from fusing the add back into VIADDMNMX.
                reg dsprev = L::splat(0);
#pragma unroll
                for (int x = 0; x < K; ++x) {
                    reg s;
                    switch (x & 3) {
                        case 0: s = L::template score<0>(w1[x >> 2], w2[x >> 2]); break;
                        case 1: s = L::template score<1>(w1[x >> 2], w2[x >> 2]); break;
                        case 2: s = L::template score<2>(w1[x >> 2], w2[x >> 2]); break;
                        default: s = L::template score<3>(w1[x >> 2], w2[x >> 2]); break;
                    }
                    const reg ds = L::add(diag, s);                // H(i-1,j-1) + S
                    const reg h = L::max3_relu(ds, E[x], f);       // max(ds, E(i,j), F(i,j), 0)
                    diag = H[x];
                    H[x] = h;
                    const reg open = L::add(h, ngoe);              // H(i,j) - (go+ge)
                    E[x] = L::addmax(E[x], nge, open);             // E(i,j+1)
                    f = L::addmax(f, nge, open);                   // F(i+1,j)
                    if (x & 1) best = L::max3(best, dsprev, ds);
                    else if (x == K - 1) best = L::max2(best, ds);
                    dsprev = ds;
                }
                out_h = H[K - 1];
                out_f = f;
                if (carry_out && col_out >= 0 && col_out < (int)ncols)
                    bnd[col_out] = make_uint2((uint32_t)out_h, (uint32_t)out_f);
            };

            for (uint32_t it = 0; it < niter; ++it) {
                const uint32_t c0 = 2 * it;                 // thread 0's columns this iteration: c0, c0 + 1
                // prefetch thread 0's inputs for the next iteration
                uint32_t nword = kPadWord;
                uint2 nb0 = make_uint2(0u, 0u), nb1 = make_uint2(0u, 0u);
                const uint32_t cn = c0 + 2;
                if (cn < ncols) {
                    nword = words[(cn >> 3) * (kTilePairs * 4) + ((cn & 7u) >> 1)];
                    if (carry_in && t == 0) { nb0 = __ldcg(bnd + cn); nb1 = __ldcg(bnd + cn + 1); }
                }
                const uint32_t pkA = __byte_perm(word, 0u, selA);
                const uint32_t pkB = __byte_perm(word, 0u, selB);
                column(pkA, (reg)b0.x, (reg)b0.y, (int)c0 - (G - 1));
                column(pkB, (reg)b1.x, (reg)b1.y, (int)c0 + 1 - (G - 1));
                word = nword; b0 = nb0; b1 = nb1;
            }
            if (p.passes > 1) { __threadfence_block(); __syncwarp(); }
        }

        // best over the rows of the group
#pragma unroll
        for (int o = G >> 1; o > 0; o >>= 1)
            best = L::max2(best, (reg)__shfl_xor_sync(0xffffffffu, best, o, G));

        if (t == 0) {
            if (L::kSeqs == 2) {
                const uint32_t bb = (uint32_t)best;
                const int s_lo = (int)(short)(bb & 0xffffu), s_hi = (int)(short)(bb >> 16);
                if (global_seq_index(p, lseq) < p.n_total) {
                    p.scores[lseq] = s_lo;
                    if (s_lo >= kOverflow16) p.resc_list[atomicAdd(p.resc_count, 1u)] = lseq;
                }
                if (global_seq_index(p, lseq + 1) < p.n_total) {
                    p.scores[lseq + 1] = s_hi;
                    if (s_hi >= kOverflow16) p.resc_list[atomicAdd(p.resc_count, 1u)] = lseq + 1;
                }
            } else {
                p.scores[lseq] = (int32_t)best;
            }
        }
    }
}

// host-side launchers; one translation unit per (lane policy, group size) family (wavefront_inst_*.cu)
cudaError_t launch_wf_l16_g4(int K, int grid, size_t smem, cudaStream_t stream, const WfParams &p);
cudaError_t launch_wf_l16_g8(int K, int grid, size_t smem, cudaStream_t stream, const WfParams &p);
cudaError_t launch_wf_l16_g16(int K, int grid, size_t smem, cudaStream_t stream, const WfParams &p);
cudaError_t launch_wf_l16_g32(int K, int grid, size_t smem, cudaStream_t stream, const WfParams &p);
cudaError_t launch_wf_l32_g32(int K, int grid, size_t smem, cudaStream_t stream, const WfParams &p);
// the two "profile in global memory" variants (G = 32, K = 32) for queries of more than kMaxSmemPasses passes
cudaError_t launch_wf_l16_gp(int grid, cudaStream_t stream, const WfParams &p);
cudaError_t launch_wf_l32_gp(int grid, cudaStream_t stream, const WfParams &p);

template <class L, int G, int K, bool GP>
cudaError_t launch_one(int grid, size_t smem, cudaStream_t stream, const WfParams &p)
{
    static size_t configured[64] = {0};            // per device: the attribute lives in the device's context
    int dev = 0;
    cudaGetDevice(&dev);
    if (smem > configured[dev & 63]) {
        cudaError_t e = cudaFuncSetAttribute(wavefront_kernel<L, G, K, GP>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured[dev & 63] = smem;
    }
    wavefront_kernel<L, G, K, GP><<<grid, kBlockThreads, smem, stream>>>(p);
    return cudaGetLastError();
}

}  // namespace swg
