// swg_common.cuh -- shared constants, parameter blocks and lane policies of the sm_100a search kernels.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <map>
#include <mutex>
#include <utility>

namespace swg {

// cudaFuncAttributeMaxDynamicSharedMemorySize of a kernel, set once per (device, kernel) and raised when a launch needs
// more: the driver call is too slow to repeat on every launch (it can block for a millisecond, which a 1 ms search
// shows as GPU idle time), and the bookkeeping must be safe with one host thread per GPU.
inline cudaError_t ensure_dynamic_smem(const void *kernel, size_t bytes)
{
    if (bytes <= 48 * 1024) return cudaSuccess;
    static std::mutex mu;
    static std::map<std::pair<int, const void *>, size_t> done;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    std::lock_guard<std::mutex> lock(mu);
    size_t &have = done[std::make_pair(dev, kernel)];
    if (bytes <= have) return cudaSuccess;
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e == cudaSuccess) have = bytes;
    return e;
}

// ---- device database layout --------------------------------------------------------------------
// Sorted sequences are grouped 16 to a TILE (8 PAIRS).  All 16 are padded to the tile's column
// count (a multiple of 8).  Columns are stored in CHUNKS of 8: one 16-byte unit per (chunk, pair)
//     unit(tile, chunk, pair) = tile_off[tile] + chunk * 8 + pair            (16-byte units)
//     byte 2*c + 0 = residue of sequence 2*pair   at column 8*chunk + c,  byte 2*c + 1 = sequence 2*pair+1
// so one 32-bit word holds two columns of a pair and the 8 pairs of a chunk are 128 contiguous
// bytes (a warp's groups read neighbouring units).  Residues are stored as code*4 so that a byte
// placed in byte lane 1 of a register is directly the profile offset code*1024 (one PRMT).
constexpr int kTileSeqs = 16;
constexpr int kTilePairs = 8;
constexpr int kChunkCols = 8;
constexpr int kPadCode = 24;                 // reference sequences.h:17
constexpr uint32_t kPadWord = 0x60606060u;   // four pad residues (24*4)

// ---- query profile layout (global and shared memory, identical) -----------------------------------
// profile[pass][letter 0..24][1024 bytes]; inside a letter row, for a group of G threads with K rows each:
//     row x of thread t  ->  (x / 16) * (G * 16) + t * 16 + (x % 16)
// i.e. 16-row chunks, thread-interleaved: the quarter-warp of an LDS.128 covers 8 consecutive threads
// -> 8 distinct 16-byte bank groups, conflict-free for G >= 8 whatever letters the threads look up.
constexpr int kLetters = 25;
constexpr int kLetterStride = 1024;
constexpr int kPassBytes = 32768;             // slice of one pass: 25 letter rows, padded to a power of two so that
                                             // pass, letter and thread offsets occupy disjoint address bits
static_assert(kLetters * kLetterStride <= kPassBytes, "profile slice");
constexpr int kMaxRowsPerThread = 32;
constexpr int kMaxPassRows = 1024;           // 32 threads x 32 rows

#ifndef SWG_BLOCK_THREADS
#define SWG_BLOCK_THREADS 512
#endif
constexpr int kBlockThreads = SWG_BLOCK_THREADS;
// gap penalties with their own kernel instantiations (immediate operands): SWIMM's defaults, -g 10 -e 2
constexpr int kFastGapOpenExtend = 12;
constexpr int kFastGapExtend = 2;
constexpr int kTripCols = 4;                 // database columns per trip of the kernel's column loop
constexpr int kBoundarySlack = 64;           // scratch-line columns beyond the longest tile (short segments are padded)
constexpr int kOverflow16 = 32000;           // a 16-bit lane whose best reaches this is redone in 32 bits

// ---- query-pair kernel (wavefront_q2.cuh): 32-bit profile entries, per-sequence pass lines ----------
constexpr int kQ2LetterStride = 4096;                       // 1024 rows x 4 bytes
constexpr int kQ2ProfileBytes = kLetters * kQ2LetterStride; // one pass: 100 KB
constexpr int kQ2MinSegCols = 80;                           // every segment is padded to at least this many columns (G = 32: 2 x 32 + 16)
constexpr int kQ2LineSlack = 8;                             // columns a line extends past its segment (prefetch overrun)

struct WfParams {
    const uint4 *db;              // tiled database, 16-byte units
    const uint64_t *tile_off;     // [ntiles + 1] in 16-byte units
    const uint32_t *tile_cols;    // [ntiles] padded column count (multiple of 8)
    uint32_t ntiles;              // local tiles
    uint32_t tile_first, tile_count;   // sub-range of tiles this launch covers (longest-first inside it)
    uint64_t n_total;             // sequences in the whole database (all shards)
    uint32_t shard, num_shards;
    const uint8_t *profile;       // [passes][25][1024]
    uint32_t passes;
    int32_t *scores;              // [ntiles * 16] local scores of the current query
    uint2 *boundary;              // [resident warps][maxcols] last-row (H, F) between passes
    uint32_t maxcols;
    uint32_t *task_counter;
    uint32_t *resc_count;         // number of entries in resc_list
    uint32_t *resc_list;          // local sequence ids to redo in 32 bits
    int gap_open_extend;          // go + ge
    int gap_extend;               // ge
    // query-pair kernel (wavefront_q2.cuh) only: the second query's outputs and the per-sequence pass lines
    int32_t *scores2;             // [ntiles * 16] local scores of the query in the high halves
    uint32_t *resc_count2;
    uint32_t *resc_list2;
    const uint64_t *line_off;     // [ntiles] first line of a tile inside `boundary` (uint2 units); 16 lines per tile
    uint64_t line_dummy;          // offset of the per-group dummy lines inside `boundary`
    // the two 16-bit lanes are independent streams of queries: a lane may start a new query in a launch in which
    // the other lane continues one
    uint32_t lane_flags;          // kLaneActive / kLaneFirst bits, low lane in bits 0-1, high lane in bits 2-3
    // column chunks of long tiles (sequence-pair kernel, short queries): when vt != NULL the launch's tasks are the
    // entries {tile, first column, columns, 0} of this table instead of whole tiles, and scores are merged with atomicMax
    const uint4 *vt;
    uint32_t vt_count;
    // long-sequence kernel (wavefront_xw.cuh) only
    uint32_t xw_warps;            // W: warps (= concurrent passes) per sequence pair, 1, 2, 4, 8 or 16
    uint32_t xw_groups;           // sequence pairs a CTA works on at the same time (<= 16 / W)
};
constexpr uint32_t kLaneActive = 1u;   // the lane holds a query: its scores are stored
constexpr uint32_t kLaneFirst = 2u;    // ... and this launch is the query's first pass (nothing to merge with)

__device__ __forceinline__ uint64_t global_seq_index(const WfParams &p, uint32_t local_seq)
{
    const uint32_t ltile = local_seq / kTileSeqs;
    return ((uint64_t)ltile * p.num_shards + p.shard) * kTileSeqs + (local_seq % kTileSeqs);
}

// PRMT with the full 4-bit selectors: bit 3 of a nibble replicates the sign of the selected byte over the
// target byte (this is what sign-extends the 8-bit profile entries).  __byte_perm() masks that bit away.
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel)
{
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}

// ---- lane policies ---------------------------------------------------------------------------------
// Lane16: two database sequences per thread group in the halves of a 32-bit register (s16x2 DPX).
// Lane32: one sequence per group, plain int32 DPX -- the exact re-computation of overflowed lanes.
struct Lane16 {
    static constexpr int kSeqs = 2;
    typedef uint32_t reg;
    static __device__ __forceinline__ reg splat(int v) { return (uint32_t)(v & 0xffff) * 0x10001u; }
    static __device__ __forceinline__ reg addmax(reg a, reg b, reg c) { return __viaddmax_s16x2(a, b, c); }
    static __device__ __forceinline__ reg max_relu(reg a, reg b) { return __vimax_s16x2_relu(a, b); }
    static __device__ __forceinline__ reg add(reg a, reg b) { return __vadd2(a, b); }
    static __device__ __forceinline__ reg max2(reg a, reg b) { return __vmaxs2(a, b); }
    static __device__ __forceinline__ reg max3(reg a, reg b, reg c) { return __vimax3_s16x2(a, b, c); }
    static __device__ __forceinline__ reg max3_relu(reg a, reg b, reg c) { return __vimax3_s16x2_relu(a, b, c); }
    // substitution scores of row byte B for the two sequences: {sext16(w_hi.b[B]), sext16(w_lo.b[B])}
    template <int B>
    static __device__ __forceinline__ reg score(uint32_t w_lo, uint32_t w_hi)
    {
        constexpr uint32_t sel = (uint32_t)B | ((8u | B) << 4) | ((4u + B) << 8) | ((12u + B) << 12);
        return prmt(w_lo, w_hi, sel);
    }
};

struct Lane32 {
    static constexpr int kSeqs = 1;
    typedef int32_t reg;
    static __device__ __forceinline__ reg splat(int v) { return v; }
    static __device__ __forceinline__ reg addmax(reg a, reg b, reg c) { return __viaddmax_s32(a, b, c); }
    static __device__ __forceinline__ reg max_relu(reg a, reg b) { return __vimax_s32_relu(a, b); }
    static __device__ __forceinline__ reg add(reg a, reg b) { return a + b; }
    static __device__ __forceinline__ reg max2(reg a, reg b) { return max(a, b); }
    static __device__ __forceinline__ reg max3(reg a, reg b, reg c) { return __vimax3_s32(a, b, c); }
    static __device__ __forceinline__ reg max3_relu(reg a, reg b, reg c) { return __vimax3_s32_relu(a, b, c); }
    template <int B>
    static __device__ __forceinline__ reg score(uint32_t w_lo, uint32_t)
    {
        constexpr uint32_t sel = (uint32_t)B | ((8u | B) << 4) | ((8u | B) << 8) | ((8u | B) << 12);
        return (int32_t)prmt(w_lo, 0u, sel);
    }
};

}  // namespace swg
