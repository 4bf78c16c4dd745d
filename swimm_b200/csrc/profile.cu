// profile.cu -- K0: query profile build.  Role of the reference's per-block score-profile build
// (CPUsearch.c:581-603) and of the Xeon-Phi query profile (MICsearch.c:124-219): substitution scores
// are gathered once per query into the byte layout the wavefront kernel reads with LDS.128.
#include "swg_internal.h"

namespace swg {

// profile[pass][letter][ (x/16)*(G*16) + t*16 + x%16 ] = submat[query[pass*G*K + t*K + x]][letter]
__global__ void build_profile_kernel(const int8_t *__restrict__ query, uint32_t m, const int8_t *__restrict__ submat,
                                     int G, int K, uint32_t passes, uint8_t *__restrict__ profile)
{
    const uint32_t rows_per_pass = (uint32_t)(G * K);
    const uint32_t total = passes * kLetters * rows_per_pass;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const uint32_t rr = i % rows_per_pass;
        const uint32_t letter = (i / rows_per_pass) % kLetters;
        const uint32_t pass = i / (rows_per_pass * kLetters);
        const uint32_t row = pass * rows_per_pass + rr;
        const uint32_t t = rr / K, x = rr % K;
        int8_t v = 0;
        if (row < m) {
            const uint32_t q = min((uint32_t)(uint8_t)query[row], 23u);   // 23 = dummy: table row 23 is all zero
            v = submat[q * 32 + letter];
        }
        profile[pass * kPassBytes + letter * kLetterStride + (x >> 4) * (G * 16) + t * 16 + (x & 15)] = (uint8_t)v;
    }
}

cudaError_t launch_build_profile(const int8_t *d_query, uint32_t m, const int8_t *d_submat, int G, int K,
                                 uint32_t passes, uint8_t *d_profile, cudaStream_t stream)
{
    cudaError_t e = cudaMemsetAsync(d_profile, 0, (size_t)passes * kPassBytes, stream);
    if (e != cudaSuccess) return e;
    const uint32_t total = passes * kLetters * (uint32_t)(G * K);
    const int threads = 256;
    const int blocks = (int)((total + threads - 1) / threads);
    build_profile_kernel<<<blocks, threads, 0, stream>>>(d_query, m, d_submat, G, K, passes, d_profile);
    return cudaGetLastError();
}

// Query-pair profile (wavefront_q2.cuh), one pass: a 32-bit entry per (row, letter) holding the scores of BOTH queries,
//   profile2[letter][ (x/4)*(G*16) + t*16 + (x%4)*4 ] = { s16 S[qB[row]][letter], s16 S[qA[row]][letter] }
// with row = row0 + t*K + x, row0 given per lane (the lanes are independent streams of queries); rows beyond a
// query's end score 0 (they can never raise its best).  Every launch has its own K, hence one call per launch.
__global__ void build_profile_q2_kernel(const int8_t *__restrict__ qa, uint32_t ma, const int8_t *__restrict__ qb,
                                        uint32_t mb, const int8_t *__restrict__ submat, int G, int K, uint32_t rowa0,
                                        uint32_t rowb0, uint32_t *__restrict__ profile)
{
    const uint32_t rows = (uint32_t)(G * K);
    const uint32_t total = kLetters * rows;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const uint32_t rr = i % rows;
        const uint32_t letter = i / rows;
        const uint32_t ra = rowa0 + rr, rb = rowb0 + rr;
        const uint32_t t = rr / K, x = rr % K;
        int va = 0, vb = 0;
        if (ra < ma) va = submat[min((uint32_t)(uint8_t)qa[ra], 23u) * 32 + letter];
        if (rb < mb) vb = submat[min((uint32_t)(uint8_t)qb[rb], 23u) * 32 + letter];
        profile[(letter * kQ2LetterStride + (x >> 2) * (G * 16) + t * 16 + (x & 3) * 4) / 4] =
            ((uint32_t)va & 0xffffu) | ((uint32_t)vb << 16);
    }
}

cudaError_t launch_build_profile_q2(const int8_t *d_qa, uint32_t ma, const int8_t *d_qb, uint32_t mb, const int8_t *d_submat,
                                    int G, int K, uint32_t rowa0, uint32_t rowb0, uint8_t *d_profile, cudaStream_t stream)
{
    const uint32_t total = kLetters * (uint32_t)(G * K);
    const int threads = 256;
    const int blocks = (int)((total + threads - 1) / threads);
    build_profile_q2_kernel<<<blocks, threads, 0, stream>>>(d_qa, ma, d_qb, mb, d_submat, G, K, rowa0, rowb0,
                                                            reinterpret_cast<uint32_t *>(d_profile));
    return cudaGetLastError();
}

// Pass lines between two launches of the query-pair kernel when ONE lane starts a new query and the other continues:
// the starting lane's half of every (H, F) entry is cleared, so that its first rows see H = F = 0 above them.
__global__ void clear_lane_kernel(uint4 *__restrict__ lines, uint64_t n16, uint32_t keep)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) {
        uint4 v = lines[i];
        v.x &= keep; v.y &= keep; v.z &= keep; v.w &= keep;
        lines[i] = v;
    }
}

cudaError_t launch_clear_lane(uint2 *d_lines, uint64_t n_entries, uint32_t keep_mask, cudaStream_t stream)
{
    const uint64_t n16 = (n_entries + 1) / 2;          // the buffer is allocated with a spare entry's worth of slack
    if (n16 == 0) return cudaSuccess;
    clear_lane_kernel<<<148 * 8, 512, 0, stream>>>(reinterpret_cast<uint4 *>(d_lines), n16, keep_mask);
    return cudaGetLastError();
}

}  // namespace swg
