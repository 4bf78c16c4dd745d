// align_ends.cu -- start / end coordinates of the alignments behind a batch's top-r hits (opt-in).
//
// The reference is score-only (CPUsearch.c:670-676 stores one int per sequence), so this pass has no counterpart
// there; its definition is oracle/sw_oracle.c:swo_align_ends.  Only the r hits of every query are aligned again, so
// the work is r alignments per query instead of n: one warp per (query, hit).
//
//   end   = the cell of the Gotoh recurrence that holds the hit's score; ties: smallest database position, then
//           smallest query position;
//   start = the same search on the reversed prefixes q[0..q_end], d[0..d_end], mapped back.
//
// One warp walks the query in strips of 32 rows (lane = row); inside a strip the columns move through the lanes as a
// systolic pipeline (lane l works on column s - l at step s, H and F of the row above arrive by shuffle); the last row
// of a strip is parked in a per-warp line in global memory and re-enters at lane 0 in the next strip.  Plain int32.
#include "swg_internal.h"

namespace swg {

namespace {

struct Cell {
    int32_t score, i, j;
};

// (score desc, j asc, i asc)
__device__ __forceinline__ bool better(const Cell &a, const Cell &b)
{
    if (a.score != b.score) return a.score > b.score;
    if (a.j != b.j) return a.j < b.j;
    return a.i < b.i;
}

struct SeqView {             // a database sequence inside the tiled layout (swg_common.cuh), residues stored as code * 4
    const uint8_t *base;     // first 16-byte unit of the sequence's tile + pair
    uint32_t half;
    __device__ __forceinline__ int code(uint32_t col) const
    {
        return base[(size_t)(col >> 3) * (kTilePairs * 16) + 2 * (col & 7u) + half] >> 2;
    }
};

// Best cell of query rows q[q0 + qstep * (0..m)) against database columns d0 + dstep * (0..n), whole warp.
__device__ Cell warp_best_cell(const int8_t *q, int32_t q0, int qstep, uint32_t m, const SeqView &d, int32_t d0, int dstep,
                               uint32_t n, const int8_t *submat /* shared, 24 x 32 */, int goe, int ge, int2 *line)
{
    const int lane = threadIdx.x & 31;
    Cell best = {0, -1, -1};
    const uint32_t strips = (m + 31) / 32;
    for (uint32_t strip = 0; strip < strips; ++strip) {
        const uint32_t r = strip * 32 + lane;
        const bool row_ok = r < m;
        int qc = row_ok ? q[q0 + qstep * (int32_t)r] : 23;
        if (qc < 0 || qc > 23) qc = 23;                      // 23 = dummy: table row 23 is all zero
        const int8_t *row = submat + 32 * qc;
        int32_t hleft = 0, e = 0, diag = 0, out_h = 0, out_f = 0;
        for (uint32_t s = 0; s < n + 31; ++s) {
            const int32_t col = (int32_t)s - lane;
            int32_t uh = __shfl_up_sync(0xffffffffu, out_h, 1);
            int32_t uf = __shfl_up_sync(0xffffffffu, out_f, 1);
            if (lane == 0) {
                uh = 0;
                uf = 0;
                if (strip > 0 && s < n) { const int2 v = line[s]; uh = v.x; uf = v.y; }
            }
            if (col >= 0 && col < (int32_t)n) {
                int32_t h = 0, f = 0;
                if (row_ok) {
                    const int dc = d.code((uint32_t)(d0 + dstep * col));
                    const int32_t sc = row[dc];
                    f = max(uf - ge, uh - goe);
                    e = max(e - ge, hleft - goe);
                    h = max(max(0, diag + sc), max(e, f));
                    if (h > best.score || (h == best.score && h > 0 && col < best.j)) { best.score = h; best.i = (int32_t)r; best.j = col; }
                    diag = uh;
                    hleft = h;
                }
                out_h = h;
                out_f = f;
                if (lane == 31 && strip + 1 < strips) line[col] = make_int2(h, f);       // col = s - 31: behind lane 0's reads
            }
        }
        __syncwarp();
    }
    // the warp's best under the tie rule
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        Cell other;
        other.score = __shfl_xor_sync(0xffffffffu, best.score, o);
        other.i = __shfl_xor_sync(0xffffffffu, best.i, o);
        other.j = __shfl_xor_sync(0xffffffffu, best.j, o);
        if (other.score > 0 && (best.score <= 0 || better(other, best))) best = other;
    }
    return best;
}

__global__ void __launch_bounds__(32) align_ends_kernel(const uint64_t *__restrict__ keys, uint32_t top, const int8_t *__restrict__ queries,
                                                        const uint32_t *__restrict__ q_off, const int8_t *__restrict__ submat_g,
                                                        int goe, int ge, const uint4 *__restrict__ db,
                                                        const uint64_t *__restrict__ tile_off, const uint32_t *__restrict__ tile_cols,
                                                        uint32_t num_shards, int2 *__restrict__ lines, uint32_t line_stride,
                                                        int32_t *__restrict__ coords)
{
    __shared__ int8_t submat[768];
    for (int i = threadIdx.x; i < 768; i += 32) submat[i] = submat_g[i];
    __syncwarp();
    const uint32_t qi = blockIdx.y, hit = blockIdx.x;
    const uint64_t key = keys[(size_t)qi * top + hit];
    int32_t *out = coords + ((size_t)qi * top + hit) * 4;
    const int32_t score = (int32_t)(key >> 32);
    if (score <= 0) {
        if (threadIdx.x < 4) out[threadIdx.x] = -1;
        return;
    }
    const uint32_t g = (uint32_t)(key & 0xffffffffu);
    const uint32_t ltile = (g / kTileSeqs) / num_shards, s = g % kTileSeqs;
    SeqView d;
    d.base = reinterpret_cast<const uint8_t *>(db + tile_off[ltile] + (s >> 1));
    d.half = s & 1u;
    const uint32_t n = tile_cols[ltile];                  // padded: pad residues score 0 and never win a tie
    const int8_t *q = queries + q_off[qi];
    const uint32_t m = q_off[qi + 1] - q_off[qi];
    int2 *line = lines + (size_t)(qi * top + hit) * line_stride;
    const Cell end = warp_best_cell(q, 0, 1, m, d, 0, 1, n, submat, goe, ge, line);
    if (end.score != score) {                              // cannot happen: the search kernels computed this score
        if (threadIdx.x < 4) out[threadIdx.x] = -2;
        return;
    }
    const Cell start = warp_best_cell(q, end.i, -1, (uint32_t)end.i + 1, d, end.j, -1, (uint32_t)end.j + 1, submat, goe, ge, line);
    if (threadIdx.x == 0) {
        out[0] = start.score == score ? end.i - start.i : -2;
        out[1] = end.i;
        out[2] = start.score == score ? end.j - start.j : -2;
        out[3] = end.j;
    }
}

}  // namespace

cudaError_t launch_align_ends(const uint64_t *d_keys, uint64_t q_count, uint64_t top, const int8_t *d_queries,
                              const uint32_t *d_q_off, const int8_t *d_submat, int goe, int ge, const uint4 *d_db,
                              const uint64_t *d_tile_off, const uint32_t *d_tile_cols, uint32_t num_shards, int2 *d_lines,
                              uint32_t line_stride, int32_t *d_coords, cudaStream_t stream)
{
    if (q_count == 0 || top == 0) return cudaSuccess;
    dim3 grid((unsigned)top, (unsigned)q_count);
    align_ends_kernel<<<grid, 32, 0, stream>>>(d_keys, (uint32_t)top, d_queries, d_q_off, d_submat, goe, ge, d_db, d_tile_off,
                                               d_tile_cols, num_shards, d_lines, line_stride, d_coords);
    return cudaGetLastError();
}

}  // namespace swg
