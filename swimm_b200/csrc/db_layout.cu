// db_layout.cu -- device-side build of the tiled database layout (role of the reference's lane
// interleave in assemble_single_chunk_db, sequences.c:704-723, done on the GPU instead of a scalar
// triple loop on the host).  Layout described in swg_common.cuh.
#include "swg_internal.h"

namespace swg {

__global__ void build_tiles_kernel(const int8_t *__restrict__ residues, const uint64_t *__restrict__ seq_off,
                                   const uint16_t *__restrict__ seq_len, const uint64_t *__restrict__ tile_off,
                                   uint32_t ntiles, uint64_t total_units, uint4 *__restrict__ db)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t u = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; u < total_units; u += stride) {
        // tile that owns 16-byte unit u: last tile with tile_off <= u
        uint32_t lo = 0, hi = ntiles;
        while (hi - lo > 1) {
            const uint32_t mid = (lo + hi) >> 1;
            if (tile_off[mid] <= u) lo = mid; else hi = mid;
        }
        const uint64_t in_tile = u - tile_off[lo];
        const uint32_t chunk = (uint32_t)(in_tile / kTilePairs);
        const uint32_t pair = (uint32_t)(in_tile % kTilePairs);
        const uint32_t sa = lo * kTileSeqs + 2 * pair, sb = sa + 1;
        const uint32_t la = seq_len[sa], lb = seq_len[sb];
        const int8_t *ra = residues + seq_off[sa];
        const int8_t *rb = residues + seq_off[sb];
        uint32_t w[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            uint32_t v = 0;
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const uint32_t col = chunk * kChunkCols + 2 * k + c;
                const uint32_t a = col < la ? min((uint32_t)(uint8_t)ra[col], (uint32_t)kPadCode) : (uint32_t)kPadCode;
                const uint32_t b = col < lb ? min((uint32_t)(uint8_t)rb[col], (uint32_t)kPadCode) : (uint32_t)kPadCode;
                v |= ((a * 4u) | ((b * 4u) << 8)) << (16 * c);
            }
            w[k] = v;
        }
        db[u] = make_uint4(w[0], w[1], w[2], w[3]);
    }
}

cudaError_t launch_build_tiles(const int8_t *d_residues, const uint64_t *d_seq_off, const uint16_t *d_seq_len,
                               const uint64_t *d_tile_off, uint32_t ntiles, uint64_t total_units, uint4 *d_db,
                               cudaStream_t stream)
{
    if (total_units == 0) return cudaSuccess;
    const int threads = 256;
    uint64_t blocks = (total_units + threads - 1) / threads;
    if (blocks > 148 * 64) blocks = 148 * 64;
    build_tiles_kernel<<<(unsigned)blocks, threads, 0, stream>>>(d_residues, d_seq_off, d_seq_len, d_tile_off, ntiles,
                                                                 total_units, d_db);
    return cudaGetLastError();
}

}  // namespace swg
