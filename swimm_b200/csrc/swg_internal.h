// swg_internal.h -- declarations shared by the translation units of libswimm_cuda.so (not part of the ABI).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "swg_common.cuh"

namespace swg {

// db_layout.cu: flat residues (this shard's sequences, concatenated) -> tiled pair-interleaved layout
cudaError_t launch_build_tiles(const int8_t *d_residues, const uint64_t *d_seq_off, const uint16_t *d_seq_len,
                               const uint64_t *d_tile_off, uint32_t ntiles, uint64_t total_units, uint4 *d_db,
                               cudaStream_t stream);

// profile.cu: query + substitution matrix -> [passes][25][1024] byte profile for a (G, K) configuration
cudaError_t launch_build_profile(const int8_t *d_query, uint32_t m, const int8_t *d_submat, int G, int K,
                                 uint32_t passes, uint8_t *d_profile, cudaStream_t stream);

// profile.cu: the query-pair profile of one launch, [25][4096] bytes (wavefront_q2.cuh): G*K rows of each lane's query
// starting at rowa0 / rowb0 (ma or mb = 0: the lane is idle)
cudaError_t launch_build_profile_q2(const int8_t *d_qa, uint32_t ma, const int8_t *d_qb, uint32_t mb, const int8_t *d_submat,
                                    int G, int K, uint32_t rowa0, uint32_t rowb0, uint8_t *d_profile, cudaStream_t stream);

// profile.cu: keep only the halves selected by keep_mask (0x0000ffff or 0xffff0000) in every pass-line entry
cudaError_t launch_clear_lane(uint2 *d_lines, uint64_t n_entries, uint32_t keep_mask, cudaStream_t stream);

// topk.cu: top-r selection on 64-bit keys (score << 32 | global index), descending, all queries of a batch
struct TopkPlan {
    uint64_t n_pad;        // local scores per query (multiple of 16)
    uint64_t top;          // keys wanted per query
    bool full_sort;        // top too large for the selection kernels: sort everything
    uint32_t slice;        // scores per stage-1 block
    uint32_t nslices;
    uint64_t scratch_keys; // keys the scratch buffer must hold
};
TopkPlan topk_plan(uint64_t n_pad, uint64_t top, uint64_t q_count);
// d_scores: [q_count][n_pad]; d_out: [q_count][top]; launches are added to *launches
cudaError_t launch_topk(const TopkPlan &plan, const int32_t *d_scores, uint64_t q_count, uint64_t n_total,
                        uint32_t shard, uint32_t num_shards, uint64_t *d_scratch, uint64_t *d_out, cudaStream_t stream,
                        uint64_t *launches);

// align_ends.cu: start / end coordinates of the alignments behind the hit keys d_keys[q_count][top] (opt-in pass);
// d_lines: [q_count * top][line_stride] scratch, d_coords: [q_count][top][4] = q_start, q_end, d_start, d_end (0-based)
cudaError_t launch_align_ends(const uint64_t *d_keys, uint64_t q_count, uint64_t top, const int8_t *d_queries,
                              const uint32_t *d_q_off, const int8_t *d_submat, int goe, int ge, const uint4 *d_db,
                              const uint64_t *d_tile_off, const uint32_t *d_tile_cols, uint32_t num_shards, int2 *d_lines,
                              uint32_t line_stride, int32_t *d_coords, cudaStream_t stream);

// pipebench.cu: measured issue rates of the kernel's instruction mix (the roofline denominator)
int pipebench_probe_count();
const char *pipebench_probe_name(int probe);
// rates[p] = 1e9 thread-instructions/s on the whole GPU, mhz[p] = SM clock seen by clock64() during probe p
cudaError_t run_pipebench_all(double *rates, double *mhz, int *sm_count, cudaStream_t stream);

}  // namespace swg
