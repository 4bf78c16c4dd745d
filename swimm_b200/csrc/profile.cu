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

}  // namespace swg
