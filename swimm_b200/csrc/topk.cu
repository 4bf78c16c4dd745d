// topk.cu -- K4: top-r selection of one query's scores on the GPU (replaces the reference's full
// descending merge sort of all n scores per query, utils.c:3-86, call site swimm.c:150-160).
//
// A hit is the 64-bit key (score << 32) | database index.  Keys are unique, and "largest key first" is
// exactly the order the reference's merge sort produces (score descending, index descending: utils.c:12
// takes the left element only when strictly greater, utils.c:52 swaps a pair on <=).
//
//   top <= kSelectMax : two launches for all queries of a batch together
//       stage 1  one block per (query, slice of the scores): radix-select the slice's `top` largest keys
//       stage 2  one block per query: radix-select `top` of the surviving candidates, bitonic-sort them
//   top >  kSelectMax : full bitonic sort of the query's keys (the `-r <n>` "print everything" case)
//
// Inside the kernels a key is stored +1 so that 0 can mean "no hit".
#include "swg_internal.h"

namespace swg {

namespace {

constexpr int kSelThreads = 1024;
constexpr uint32_t kSelectMax = 2048;

struct ScoreSource {             // keys made on the fly from one query's local scores
    const int32_t *scores;
    uint32_t first, count;       // slice [first, first + count) of local sequence ids
    uint64_t n_total;
    uint32_t shard, num_shards;
    __device__ __forceinline__ uint32_t size() const { return count; }
    __device__ __forceinline__ uint64_t get(uint32_t i) const
    {
        const uint32_t ls = first + i;
        const uint64_t g = ((uint64_t)(ls / kTileSeqs) * num_shards + shard) * kTileSeqs + (ls % kTileSeqs);
        if (g >= n_total) return 0;
        return (((uint64_t)(uint32_t)scores[ls] << 32) | g) + 1;
    }
};

struct KeySource {               // candidate keys (already +1, 0 = empty)
    const uint64_t *keys;
    uint32_t count;
    __device__ __forceinline__ uint32_t size() const { return count; }
    __device__ __forceinline__ uint64_t get(uint32_t i) const { return keys[i]; }
};

struct SelectSmem {
    uint32_t hist[256];
    uint64_t red[32];
    uint64_t prefix, mask;
    uint32_t need, out_count, done;
};

// The `r` largest non-zero keys of src go to dst[0..r) in arbitrary order (missing ones are 0).
// All threads of the block must call it.  dst may be shared or global memory.
template <class Src>
__device__ void block_select(const Src &src, uint32_t r, uint64_t *dst, SelectSmem &sm)
{
    const uint32_t n = src.size();
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    for (uint32_t i = tid; i < r; i += blockDim.x) dst[i] = 0;
    // largest key -> first digit that can differ
    uint64_t mx = 0;
    for (uint32_t i = tid; i < n; i += blockDim.x) { const uint64_t k = src.get(i); mx = k > mx ? k : mx; }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { const uint64_t v = __shfl_xor_sync(0xffffffffu, mx, o); mx = v > mx ? v : mx; }
    if (lane == 0) sm.red[wid] = mx;
    __syncthreads();
    if (tid == 0) {
        uint64_t m = 0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) m = sm.red[w] > m ? sm.red[w] : m;
        sm.red[0] = m;
        sm.prefix = 0; sm.mask = 0; sm.need = r; sm.out_count = 0; sm.done = 0;
    }
    __syncthreads();
    mx = sm.red[0];
    if (mx == 0 || r == 0) { __syncthreads(); return; }
    int shift = 56;
    while (shift > 0 && (mx >> shift) == 0) shift -= 8;

    for (; shift >= 0; shift -= 8) {
        for (int i = tid; i < 256; i += blockDim.x) sm.hist[i] = 0;
        __syncthreads();
        const uint64_t prefix = sm.prefix, mask = sm.mask;
        for (uint32_t base = 0; base < n; base += blockDim.x) {
            const uint32_t i = base + tid;
            const uint64_t k = i < n ? src.get(i) : 0;
            const bool cand = k != 0 && (k & mask) == prefix;
            const uint32_t digit = (uint32_t)(k >> shift) & 255u;
            // warp-aggregated histogram: one shared-memory atomic per distinct digit per warp
            const uint32_t peers = __match_any_sync(0xffffffffu, cand ? digit : 256u + lane);
            if (cand && lane == (__ffs(peers) - 1)) atomicAdd(&sm.hist[digit], __popc(peers));
        }
        __syncthreads();
        if (tid == 0) {
            uint32_t need = sm.need, b = 255;
            for (;; --b) {
                const uint32_t c = sm.hist[b];
                if (c >= need || b == 0) break;
                need -= c;
            }
            // fewer candidates than needed (n < r): take everything that is left
            if (sm.hist[b] < need) { sm.done = 2; }
            else {
                sm.prefix = prefix | ((uint64_t)b << shift);
                sm.mask = mask | (0xffull << shift);
                sm.need = need;
                if (sm.hist[b] == need) sm.done = 1;          // every candidate with this prefix is a hit
            }
        }
        __syncthreads();
        if (sm.done) break;
    }
    // keys >= threshold are the hits (exactly r of them, or all of them when n < r)
    const uint64_t thr = sm.done == 2 ? 1 : sm.prefix;
    for (uint32_t i = tid; i < n; i += blockDim.x) {
        const uint64_t k = src.get(i);
        if (k != 0 && k >= thr) {
            const uint32_t pos = atomicAdd(&sm.out_count, 1u);
            if (pos < r) dst[pos] = k;
        }
    }
    __syncthreads();
}

__global__ void __launch_bounds__(kSelThreads)
topk_stage1_kernel(const int32_t *scores, uint64_t n_pad, uint64_t n_total, uint32_t shard, uint32_t num_shards,
                   uint32_t slice, uint32_t nslices, uint32_t top, uint64_t *cand)
{
    __shared__ SelectSmem sm;
    const uint32_t q = blockIdx.y, s = blockIdx.x;
    ScoreSource src;
    src.scores = scores + (size_t)q * n_pad;
    src.first = s * slice;
    src.count = (uint32_t)min((uint64_t)slice, n_pad - (uint64_t)s * slice);
    src.n_total = n_total;
    src.shard = shard;
    src.num_shards = num_shards;
    block_select(src, top, cand + ((size_t)q * nslices + s) * top, sm);
}

__global__ void __launch_bounds__(kSelThreads)
topk_stage2_kernel(const uint64_t *cand, uint32_t ncand, uint32_t top, uint32_t top_pow2, uint64_t *out)
{
    __shared__ SelectSmem sm;
    __shared__ uint64_t best[kSelectMax];
    const uint32_t q = blockIdx.x;
    KeySource src;
    src.keys = cand + (size_t)q * ncand;
    src.count = ncand;
    for (uint32_t i = threadIdx.x; i < top_pow2; i += blockDim.x) best[i] = 0;
    __syncthreads();
    block_select(src, top, best, sm);
    // bitonic sort, descending
    for (uint32_t k = 2; k <= top_pow2; k <<= 1) {
        for (uint32_t j = k >> 1; j > 0; j >>= 1) {
            for (uint32_t i = threadIdx.x; i < top_pow2; i += blockDim.x) {
                const uint32_t p = i ^ j;
                if (p > i) {
                    const uint64_t a = best[i], b = best[p];
                    const bool desc = (i & k) == 0;
                    if (desc ? a < b : a > b) { best[i] = b; best[p] = a; }
                }
            }
            __syncthreads();
        }
    }
    for (uint32_t i = threadIdx.x; i < top; i += blockDim.x) {
        const uint64_t k = best[i];
        out[(size_t)q * top + i] = k ? k - 1 : 0;
    }
}

// ---- full sort (top > kSelectMax) ----------------------------------------------------------------
__global__ void make_keys_kernel(const int32_t *scores, uint64_t n_pad, uint64_t n_total, uint32_t shard,
                                 uint32_t num_shards, uint64_t npow2, uint64_t *keys)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npow2) return;
    uint64_t k = 0;
    if (i < n_pad) {
        ScoreSource src;
        src.scores = scores; src.first = 0; src.count = (uint32_t)n_pad;
        src.n_total = n_total; src.shard = shard; src.num_shards = num_shards;
        k = src.get((uint32_t)i);
    }
    keys[i] = k;
}

__global__ void bitonic_step_kernel(uint64_t *keys, uint64_t npow2, uint64_t k, uint64_t j)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= npow2) return;
    const uint64_t p = i ^ j;
    if (p > i) {
        const uint64_t a = keys[i], b = keys[p];
        const bool desc = (i & k) == 0;
        if (desc ? a < b : a > b) { keys[i] = b; keys[p] = a; }
    }
}

// all (k, j) steps with j < 1024 of one 2048-key block, in shared memory
__global__ void __launch_bounds__(1024) bitonic_local_kernel(uint64_t *keys, uint64_t npow2, uint64_t k_first, uint64_t k_last)
{
    __shared__ uint64_t s[2048];
    const uint64_t base = (uint64_t)blockIdx.x * 2048;
    for (uint32_t i = threadIdx.x; i < 2048; i += blockDim.x) s[i] = base + i < npow2 ? keys[base + i] : 0;
    __syncthreads();
    for (uint64_t k = k_first; k <= k_last; k <<= 1) {
        for (uint32_t j = (uint32_t)min((uint64_t)1024, k >> 1); j > 0; j >>= 1) {
            for (uint32_t i = threadIdx.x; i < 2048; i += blockDim.x) {
                const uint32_t p = i ^ j;
                if (p > i) {
                    const uint64_t a = s[i], b = s[p];
                    const bool desc = ((base + i) & k) == 0;
                    if (desc ? a < b : a > b) { s[i] = b; s[p] = a; }
                }
            }
            __syncthreads();
        }
    }
    for (uint32_t i = threadIdx.x; i < 2048; i += blockDim.x)
        if (base + i < npow2) keys[base + i] = s[i];
}

// the scratch holds npow2 keys of THIS shard; `top` is clamped to the whole database, so on a sharded context it can
// exceed them: the rest of the row is "no hit"
__global__ void emit_sorted_kernel(const uint64_t *keys, uint64_t npow2, uint64_t top, uint64_t *out)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < top) { const uint64_t k = i < npow2 ? keys[i] : 0; out[i] = k ? k - 1 : 0; }
}

uint64_t pow2_at_least(uint64_t v)
{
    uint64_t p = 1;
    while (p < v) p <<= 1;
    return p;
}

}  // namespace

TopkPlan topk_plan(uint64_t n_pad, uint64_t top, uint64_t q_count)
{
    TopkPlan pl;
    pl.n_pad = n_pad;
    pl.top = top;
    pl.full_sort = top > kSelectMax;
    if (pl.full_sort) {
        pl.slice = 0;
        pl.nslices = 0;
        pl.scratch_keys = pow2_at_least(n_pad < 2048 ? 2048 : n_pad);
    } else {
        // slices of 8192 scores; smaller ones (down to 1024) when the batch would otherwise not fill the GPU with blocks
        // (a single short search: the selection is latency-bound, more blocks in parallel shorten it)
        uint64_t slice = 8192;
        while (slice > 1024 && (n_pad + slice - 1) / slice * q_count < 296) slice /= 2;
        if (slice < 16 * top) slice = 16 * top;
        pl.slice = (uint32_t)slice;
        pl.nslices = (uint32_t)((n_pad + slice - 1) / slice);
        if (pl.nslices == 0) pl.nslices = 1;
        pl.scratch_keys = (uint64_t)pl.nslices * top * q_count;
    }
    if (pl.scratch_keys == 0) pl.scratch_keys = 1;
    return pl;
}

cudaError_t launch_topk(const TopkPlan &pl, const int32_t *d_scores, uint64_t q_count, uint64_t n_total, uint32_t shard,
                        uint32_t num_shards, uint64_t *d_scratch, uint64_t *d_out, cudaStream_t stream, uint64_t *launches)
{
    if (pl.top == 0 || q_count == 0) return cudaSuccess;
    if (!pl.full_sort) {
        const uint32_t top = (uint32_t)pl.top;
        dim3 g1(pl.nslices, (unsigned)q_count);
        topk_stage1_kernel<<<g1, kSelThreads, 0, stream>>>(d_scores, pl.n_pad, n_total, shard, num_shards, pl.slice,
                                                           pl.nslices, top, d_scratch);
        uint32_t p2 = 1;
        while (p2 < top) p2 <<= 1;
        topk_stage2_kernel<<<(unsigned)q_count, kSelThreads, 0, stream>>>(d_scratch, pl.nslices * top, top, p2, d_out);
        if (launches) *launches += 2;
        return cudaGetLastError();
    }
    const uint64_t np2 = pl.scratch_keys;
    const unsigned blocks = (unsigned)((np2 + 255) / 256);
    for (uint64_t q = 0; q < q_count; ++q) {
        make_keys_kernel<<<blocks, 256, 0, stream>>>(d_scores + q * pl.n_pad, pl.n_pad, n_total, shard, num_shards, np2,
                                                     d_scratch);
        bitonic_local_kernel<<<(unsigned)(np2 / 2048), 1024, 0, stream>>>(d_scratch, np2, 2, 2048);
        if (launches) *launches += 2;
        for (uint64_t k = 4096; k <= np2; k <<= 1) {
            for (uint64_t j = k >> 1; j >= 2048; j >>= 1) {
                bitonic_step_kernel<<<blocks, 256, 0, stream>>>(d_scratch, np2, k, j);
                if (launches) *launches += 1;
            }
            // the remaining strides 1024..1 of this k stay inside 2048-key blocks
            bitonic_local_kernel<<<(unsigned)(np2 / 2048), 1024, 0, stream>>>(d_scratch, np2, k, k);
            if (launches) *launches += 1;
        }
        emit_sorted_kernel<<<(unsigned)((pl.top + 255) / 256), 256, 0, stream>>>(d_scratch, np2, pl.top, d_out + q * pl.top);
        if (launches) *launches += 1;
    }
    return cudaGetLastError();
}

}  // namespace swg
