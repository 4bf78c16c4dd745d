// pipebench.cu -- measured issue rates of the integer/SIMD instructions the search kernel is made of.
//
// The roofline the search is reported against is NOT a datasheet number: it is
// the rate at which this chip issues the kernel's own instruction mix when nothing else is in the way
// (no memory, no shuffles, no dependencies shorter than the pipe latency).  Each probe below runs
// kChains independent dependency chains per thread, 16 warps per SM on every SM, and reports
// 1e9 thread-instructions per second for the whole GPU plus the SM clock seen by clock64().
//
// Stand-alone use on the GPU box:   ./pipebench            (prints one JSON object)
// Library use:                     swg::run_pipebench_all()   (exported through swg_gpu_pipebench)
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "swg_internal.h"

namespace swg {

namespace {

#ifndef SWG_PIPEBENCH_CHAINS
#define SWG_PIPEBENCH_CHAINS 8
#endif
constexpr int kChains = SWG_PIPEBENCH_CHAINS;   // independent chains per thread (latency 4-6 cycles, 4 warps per scheduler)
constexpr int kThreads = 512;
constexpr int kBlocksPerSm = 1;
constexpr int kUnroll = 16;      // cells per chain per loop trip (loop overhead < 2 %)

enum Probe {
    P_VIADDMNMX = 0, P_VIMNMX, P_VIMNMX_RELU, P_VIMNMX3, P_VIADD16, P_PRMT, P_IMAD, P_IADD3, P_LOP3,
    P_HFMA2, P_HMNMX2, P_IDP4A, P_MIX65, P_MIX55_IMAD, P_PAIR_DPX_IMAD, P_PAIR_DPX_HFMA2, P_PAIR_DPX_PRMT,
    P_PAIR_DPX_IDP, P_SHFL, P_LDS128, P_PAIR_DPX_SHFL, P_PAIR_DPX_LDS, P_MIX_V2, P_PAIR_DPX_VIADD, P_PAIR_DPX_HMNMX2,
    P_PAIR_DPX_FFMA, P_PAIR_VIADD_IMAD, P_MIX_A, P_MIX_B, P_MIX_C, P_PAIR_MAX3_VIADD, P_PAIR_PRMT_VIADD,
    P_MIX_V2_IMM, P_MIX_Q2_IMM, P_MIX_V2_GE_IMM, P_MIX_Q2_GE_IMM, P_COUNT
};

const char *const kProbeName[P_COUNT] = {
    "viaddmnmx_s16x2", "vimnmx_s16x2", "vimnmx3_s16x2_relu", "vimnmx3_s16x2", "viadd_16x2", "prmt", "imad",
    "iadd3_two_fused_adds", "lop3", "hfma2", "hmnmx2", "idp4a", "mix_6p5_dpx", "mix_5p5_dpx_plus_imad", "pair_viaddmnmx_imad",
    "pair_viaddmnmx_hfma2", "pair_viaddmnmx_prmt", "pair_viaddmnmx_idp4a", "shfl_up", "lds128",
    "pair_viaddmnmx_shfl", "pair_viaddmnmx_lds128", "mix_v2_4p5_alu_2_viadd", "pair_viaddmnmx_viadd16x2",
    "pair_viaddmnmx_hmnmx2", "pair_viaddmnmx_ffma", "pair_viadd16x2_imad", "mix_a_hef_3alu_2viadd",
    "mix_b_hef_prmt_4alu_2viadd", "mix_c_hef_best_3p5alu_2viadd", "pair_vimnmx3_viadd16x2", "pair_prmt_viadd16x2",
    "mix_v2_immediate_penalties", "mix_q2_3p5alu_2viadd_immediate", "mix_v2_extend_immediate_open_register",
    "mix_q2_extend_immediate_open_register"};

// thread-instructions counted per inner step of one chain
__host__ __device__ constexpr double probe_instr(int p)
{
    return p == P_MIX65 ? 6.5 : p == P_MIX_V2 ? 6.5 : p == P_MIX_V2_IMM ? 6.5 : p == P_MIX_V2_GE_IMM ? 6.5 : p == P_MIX_Q2_GE_IMM ? 5.5 : p == P_MIX_A ? 5.0 : p == P_MIX_B ? 6.0 : p == P_MIX_C ? 5.5 : p == P_MIX_Q2_IMM ? 5.5 :
           (p == P_PAIR_MAX3_VIADD || p == P_PAIR_PRMT_VIADD) ? 2.0 : (p >= P_PAIR_DPX_VIADD && p <= P_PAIR_VIADD_IMAD) ? 2.0 : p == P_MIX55_IMAD ? 6.5 : p == P_IADD3 ? 0.5 : (p >= P_PAIR_DPX_IMAD && p <= P_PAIR_DPX_IDP) ? 2.0
         : (p == P_PAIR_DPX_SHFL || p == P_PAIR_DPX_LDS) ? 2.0 : 1.0;
}

__device__ __forceinline__ uint32_t hfma2_u(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t d;
    asm volatile("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ uint32_t hmax2_u(uint32_t a, uint32_t b)
{
    uint32_t d;
    asm volatile("max.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}
__device__ __forceinline__ uint32_t iadd_u(uint32_t a, uint32_t b)
{
    uint32_t d;
    asm volatile("add.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}
__device__ __forceinline__ uint32_t lop3_u(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t d;
    asm volatile("lop3.b32 %0, %1, %2, %3, 0x6a;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ uint32_t ffma_u(uint32_t a, uint32_t b, uint32_t c)
{
    float d;
    asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(__uint_as_float(a)), "f"(__uint_as_float(b)), "f"(__uint_as_float(c)));
    return __float_as_uint(d);
}
__device__ __forceinline__ uint32_t hmin2_u(uint32_t a, uint32_t b)
{
    uint32_t d;
    asm volatile("min.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}
__device__ __forceinline__ uint32_t imad_u(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t d;
    asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

template <int P>
__global__ void __launch_bounds__(kThreads, kBlocksPerSm)
probe_kernel(uint32_t *out, uint32_t iters, const uint32_t *consts, long long *cycles)
{
    // operands come from memory so that they live in registers (kernel parameters would be re-read with LDC)
    const uint32_t k1 = consts[0], k2 = consts[1], k3 = consts[2];
    __shared__ uint4 sm[kThreads];
    uint32_t v[kChains], e[kChains], f[kChains], b[kChains], d[kChains], g[kChains];
#pragma unroll
    for (int c = 0; c < kChains; ++c) {
        v[c] = threadIdx.x * 0x10003u + c * k1;
        e[c] = v[c] ^ k2;
        f[c] = v[c] + k3;
        b[c] = 0;
        d[c] = v[c] >> 3;
        g[c] = v[c] >> 5;
    }
    sm[threadIdx.x] = make_uint4(v[0], v[1], v[2], v[3]);
    __syncthreads();
    const uint32_t lds_addr = (threadIdx.x ^ (k1 & 1)) % kThreads;
    const long long t0 = clock64();
    for (uint32_t it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
#pragma unroll
            for (int c = 0; c < kChains; ++c) {
                if (P == P_VIADDMNMX) v[c] = __viaddmax_s16x2(v[c], k1, k2);
                else if (P == P_VIMNMX) v[c] = (u & 1) ? __vmaxs2(v[c], e[c]) : __vmins2(v[c], f[c]);   // max/min alternate: two max would fuse into VIMNMX3
                else if (P == P_VIMNMX_RELU) v[c] = __vimax3_s16x2_relu(v[c], e[c], f[c]);
                else if (P == P_VIMNMX3) v[c] = __vimax3_s16x2(v[c], e[c], f[c]);
                else if (P == P_VIADD16) v[c] = __vadd2(v[c], e[c]);
                else if (P == P_PRMT) v[c] = __byte_perm(v[c], e[c], f[c]);
                else if (P == P_IMAD) v[c] = imad_u(v[c], k1, e[c]);
                else if (P == P_IADD3) v[c] = iadd_u(v[c], e[c]);
                else if (P == P_LOP3) v[c] = lop3_u(v[c], e[c], k1);
                else if (P == P_HFMA2) v[c] = hfma2_u(v[c], k1, e[c]);
                else if (P == P_HMNMX2) v[c] = (u & 1) ? hmax2_u(v[c], e[c]) : hmin2_u(v[c], f[c]);
                else if (P == P_IDP4A) v[c] = __dp4a((int)v[c], (int)k1, (int)e[c]);
                else if (P == P_MIX65 || P == P_MIX55_IMAD) {
                    // one cell pair of the kernel: score pack, H, E, F and (every other cell) the running best
                    const uint32_t s = prmt(e[c], f[c], 0xc480 + (u & 3) * 0x1111);
                    const uint32_t a = __viaddmax_s16x2(d[c], s, e[c]);
                    const uint32_t h = __vimax3_s16x2(a, f[c], k3);
                    uint32_t open;
                    if (P == P_MIX65) open = __vadd2(h, k1);
                    else open = imad_u(h, k2, k1);
                    e[c] = __viaddmax_s16x2(e[c], k2, open);
                    f[c] = __viaddmax_s16x2(f[c], k2, open);
                    if (u & 1) b[c] = __vimax3_s16x2(b[c], d[c], h);
                    d[c] = v[c];
                    v[c] = h;
                } else if (P == P_MIX_V2) {
                    // the kernel's cell: PRMT, VIADD (ds), VIMNMX3.RELU (H), VIADD (open), VIADDMNMX x2 (E, F), VIMNMX3 / 2 (best)
                    const uint32_t s = prmt(e[c], f[c], 0xc480 + (u & 3) * 0x1111);
                    const uint32_t ds = __vadd2(d[c], s);
                    const uint32_t h = __vimax3_s16x2_relu(ds, e[c], f[c]);
                    const uint32_t open = __vadd2(h, k1);
                    e[c] = __viaddmax_s16x2(e[c], k2, open);
                    f[c] = __viaddmax_s16x2(f[c], k2, open);
                    if (u & 1) b[c] = __vimax3_s16x2(b[c], g[c], ds);
                    g[c] = ds;
                    d[c] = v[c];
                    v[c] = h;
                } else if (P == P_MIX_A || P == P_MIX_B || P == P_MIX_C || P == P_MIX_V2_IMM || P == P_MIX_Q2_IMM || P == P_MIX_V2_GE_IMM ||
                           P == P_MIX_Q2_GE_IMM) {
                    uint32_t s = e[c] ^ 0u;
                    if (P == P_MIX_B || P == P_MIX_V2_IMM || P == P_MIX_V2_GE_IMM) s = prmt(e[c], f[c], 0xc480 + (u & 3) * 0x1111);
                    else s = g[c];
                    const uint32_t ds = __vadd2(d[c], s);
                    const uint32_t h = __vimax3_s16x2_relu(ds, e[c], f[c]);
                    constexpr bool IMM = (P == P_MIX_V2_IMM || P == P_MIX_Q2_IMM);
                    // only the gap-EXTEND penalty as an immediate (it is the third operand of the two VIADDMNMX), the
                    // gap-open term from a register (second operand of a VIADD.16x2)
                    constexpr bool GE_IMM = (P == P_MIX_V2_GE_IMM || P == P_MIX_Q2_GE_IMM);
                    const uint32_t open = IMM ? __vadd2(h, 0xfff4fff4u) : __vadd2(h, k1);
                    e[c] = (IMM || GE_IMM) ? __viaddmax_s16x2(e[c], 0xfffefffeu, open) : __viaddmax_s16x2(e[c], k2, open);
                    f[c] = (IMM || GE_IMM) ? __viaddmax_s16x2(f[c], 0xfffefffeu, open) : __viaddmax_s16x2(f[c], k2, open);
                    if ((P == P_MIX_C || IMM || GE_IMM) && (u & 1)) b[c] = __vimax3_s16x2(b[c], g[c], ds);
                    if (P != P_MIX_A && P != P_MIX_B) g[c] = ds;
                    d[c] = v[c];
                    v[c] = h;
                } else if (P == P_PAIR_MAX3_VIADD) { v[c] = __vimax3_s16x2(v[c], k1, k2); e[c] = __vadd2(e[c], k3); }
                else if (P == P_PAIR_PRMT_VIADD) { v[c] = prmt(v[c], k1, f[c]); e[c] = __vadd2(e[c], k3); }
                else if (P == P_PAIR_DPX_VIADD) { v[c] = __viaddmax_s16x2(v[c], k1, k2); e[c] = __vadd2(e[c], k3); }
                else if (P == P_PAIR_DPX_HMNMX2) { v[c] = __viaddmax_s16x2(v[c], k1, k2); e[c] = (u & 1) ? hmax2_u(e[c], k3) : hmin2_u(e[c], f[c]); }
                else if (P == P_PAIR_DPX_FFMA) { v[c] = __viaddmax_s16x2(v[c], k1, k2); e[c] = ffma_u(e[c], k1, k3); }
                else if (P == P_PAIR_VIADD_IMAD) { v[c] = __vadd2(v[c], k1); e[c] = imad_u(e[c], k1, k3); }
                else if (P == P_PAIR_DPX_IMAD) { v[c] = __viaddmax_s16x2(v[c], k1, k2); e[c] = imad_u(e[c], k1, k3); }
                else if (P == P_PAIR_DPX_HFMA2) { v[c] = __viaddmax_s16x2(v[c], k1, k2); e[c] = hfma2_u(e[c], k1, k3); }
                else if (P == P_PAIR_DPX_PRMT) { v[c] = __viaddmax_s16x2(v[c], k1, k2); e[c] = __byte_perm(e[c], k3, f[c]); }
                else if (P == P_PAIR_DPX_IDP) { v[c] = __viaddmax_s16x2(v[c], k1, k2); e[c] = __dp4a((int)e[c], (int)k1, (int)k3); }
                else if (P == P_SHFL) v[c] = __shfl_up_sync(0xffffffffu, v[c], 1, 8);
                else if (P == P_LDS128) {
                    const uint4 q = sm[(lds_addr + (v[c] & 1)) % kThreads];
                    v[c] = q.x ^ q.y ^ q.z ^ q.w;   // 3 LOP3 per load: the load rate is what is probed (counted as 1)
                } else if (P == P_PAIR_DPX_SHFL) { v[c] = __viaddmax_s16x2(v[c], k1, k2); e[c] = __shfl_up_sync(0xffffffffu, e[c], 1, 8); }
                else if (P == P_PAIR_DPX_LDS) {
                    v[c] = __viaddmax_s16x2(v[c], k1, k2);
                    if ((c & 3) == 0) { const uint4 q = sm[(lds_addr + (e[c] & 1)) % kThreads]; e[c] = q.x; f[c] ^= q.w; }
                }
            }
        }
    }
    const long long t1 = clock64();
    uint32_t acc = 0;
#pragma unroll
    for (int c = 0; c < kChains; ++c) acc ^= v[c] ^ e[c] ^ f[c] ^ b[c] ^ d[c] ^ g[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int P>
cudaError_t run_probe(int sms, uint32_t *d_out, const uint32_t *d_consts, long long *d_cyc, uint32_t iters, cudaStream_t stream, double *ginstr,
                      double *mhz)
{
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    const int grid = sms * kBlocksPerSm;
    probe_kernel<P><<<grid, kThreads, 0, stream>>>(d_out, iters / 8 + 1, d_consts, d_cyc);
    double best_ms = 1e30;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEventRecord(a, stream);
        probe_kernel<P><<<grid, kThreads, 0, stream>>>(d_out, iters, d_consts, d_cyc);
        cudaEventRecord(b, stream);
        cudaError_t e = cudaEventSynchronize(b);
        if (e != cudaSuccess) return e;
        float ms = 0.f;
        cudaEventElapsedTime(&ms, a, b);
        if (ms < best_ms) best_ms = ms;
    }
    long long cyc = 0;
    cudaMemcpy(&cyc, d_cyc, sizeof(cyc), cudaMemcpyDeviceToHost);
    const double instr = (double)grid * kThreads * (double)iters * kUnroll * kChains * probe_instr(P);
    *ginstr = instr / (best_ms * 1e-3) / 1e9;
    *mhz = (double)cyc / (best_ms * 1e-3) / 1e6;
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    return cudaGetLastError();
}

template <int P>
struct ProbeLoop {
    static cudaError_t run(int sms, uint32_t *d_out, const uint32_t *d_consts, long long *d_cyc, uint32_t iters,
                           cudaStream_t s, double *rates, double *mhz)
    {
        cudaError_t e = run_probe<P>(sms, d_out, d_consts, d_cyc, iters, s, &rates[P], &mhz[P]);
        if (e != cudaSuccess) return e;
        return ProbeLoop<P + 1>::run(sms, d_out, d_consts, d_cyc, iters, s, rates, mhz);
    }
};
template <>
struct ProbeLoop<P_COUNT> {
    static cudaError_t run(int, uint32_t *, const uint32_t *, long long *, uint32_t, cudaStream_t, double *, double *)
    {
        return cudaSuccess;
    }
};

}  // namespace

int pipebench_probe_count() { return P_COUNT; }
const char *pipebench_probe_name(int p) { return (p >= 0 && p < P_COUNT) ? kProbeName[p] : ""; }

cudaError_t run_pipebench_all(double *rates, double *mhz, int *sm_count, cudaStream_t stream)
{
    int dev = 0, sms = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return e;
    uint32_t *d_out = nullptr;
    long long *d_cyc = nullptr;
    e = cudaMalloc(&d_out, (size_t)sms * kBlocksPerSm * kThreads * sizeof(uint32_t));
    if (e != cudaSuccess) return e;
    e = cudaMalloc(&d_cyc, (size_t)sms * kBlocksPerSm * sizeof(long long));
    if (e != cudaSuccess) { cudaFree(d_out); return e; }
    uint32_t *d_consts = nullptr;
    const uint32_t h_consts[4] = {0xfffefffeu, 0x00010001u, 0x00400040u, 0u};
    e = cudaMalloc(&d_consts, sizeof(h_consts));
    if (e == cudaSuccess) e = cudaMemcpy(d_consts, h_consts, sizeof(h_consts), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = ProbeLoop<0>::run(sms, d_out, d_consts, d_cyc, 5000, stream, rates, mhz);
    cudaFree(d_consts);
    cudaFree(d_out);
    cudaFree(d_cyc);
    if (sm_count) *sm_count = sms;
    return e;
}

}  // namespace swg

#ifdef SWG_PIPEBENCH_MAIN
int main()
{
    double rates[64], mhz[64];
    int sms = 0;
    cudaError_t e = swg::run_pipebench_all(rates, mhz, &sms, 0);
    if (e != cudaSuccess) {
        fprintf(stderr, "pipebench: %s\n", cudaGetErrorString(e));
        return 1;
    }
    printf("{\"sms\": %d, \"threads_per_sm\": %d, \"chains\": %d, \"probes\": {\n", sms, swg::kThreads * swg::kBlocksPerSm,
           swg::kChains);
    const int n = swg::pipebench_probe_count();
    for (int p = 0; p < n; ++p) {
        const double per_clk_sm = rates[p] * 1e9 / (mhz[p] * 1e6) / sms;
        printf("  \"%s\": {\"ginstr_per_s\": %.1f, \"sm_mhz\": %.0f, \"thread_instr_per_clk_per_sm\": %.2f}%s\n",
               swg::pipebench_probe_name(p), rates[p], mhz[p], per_clk_sm, p + 1 < n ? "," : "");
    }
    printf("}}\n");
    return 0;
}
#endif
