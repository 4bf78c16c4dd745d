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

// topk.cu: per-query top-r selection on 64-bit keys (score << 32 | global index), descending
struct TopkPlan {
    uint64_t n_pad;        // local scores (multiple of 16)
    uint64_t top;          // keys wanted
    uint64_t scratch_keys; // keys each of the two scratch buffers must hold
};
TopkPlan topk_plan(uint64_t n_pad, uint64_t top);
// result: out[0 .. top) ; launches counted into *launches
cudaError_t launch_topk(const int32_t *d_scores, uint64_t n_pad, uint64_t n_total, uint32_t shard, uint32_t num_shards,
                        uint64_t top, uint64_t *d_scratch_a, uint64_t *d_scratch_b, uint64_t *d_out,
                        cudaStream_t stream, uint64_t *launches);

// pipebench.cu: measured issue rate of the kernel's integer instruction mix (the roofline denominator)
struct PipeRates {
    double mix6_ginstr;      // 1e9 thread-instructions/s of the 6-op s16x2 recurrence mix, whole GPU
    double mix7_ginstr;      // the same plus the PRMT score pack
    double viaddmnmx_ginstr; // single-opcode rates
    double vimnmx_ginstr;
    double vimnmx3_ginstr;
    double viadd_ginstr;
    double prmt_ginstr;
    double imad_ginstr;
    double idp_ginstr;
    double alu_fma_pair_ginstr;  // VIADDMNMX interleaved 1:1 with IMAD (do the two pipes overlap?)
    double sm_clock_mhz;     // clock derived from clock64() over the run
};
cudaError_t run_pipebench(PipeRates *out, cudaStream_t stream);

}  // namespace swg
