// wavefront_inst_l16_g8.cu -- instantiations of wavefront_kernel<Lane16, 8, K, false> for K = 1..32 (one file per
// family so that the families compile in parallel).
#include "wavefront.cuh"

namespace swg {

cudaError_t launch_wf_l16_g8(int K, int grid, size_t smem, cudaStream_t stream, const WfParams &p)
{
    if (p.gap_open_extend == kFastGapOpenExtend && p.gap_extend == kFastGapExtend) {
        switch (K) {
#define SWG_CASE(k) case k: return launch_one<Lane16, 8, k, false, false, kFastGapOpenExtend, kFastGapExtend>(grid, smem, stream, p);
            SWG_CASE(1) SWG_CASE(2) SWG_CASE(3) SWG_CASE(4) SWG_CASE(5) SWG_CASE(6) SWG_CASE(7) SWG_CASE(8)
            SWG_CASE(9) SWG_CASE(10) SWG_CASE(11) SWG_CASE(12) SWG_CASE(13) SWG_CASE(14) SWG_CASE(15) SWG_CASE(16)
            SWG_CASE(17) SWG_CASE(18) SWG_CASE(19) SWG_CASE(20) SWG_CASE(21) SWG_CASE(22) SWG_CASE(23) SWG_CASE(24)
            SWG_CASE(25) SWG_CASE(26) SWG_CASE(27) SWG_CASE(28) SWG_CASE(29) SWG_CASE(30) SWG_CASE(31) SWG_CASE(32)
#undef SWG_CASE
            default: return cudaErrorInvalidValue;
        }
    }
    if (p.gap_extend == 1) {            // gap-extend penalty as an immediate, any gap-open penalty (see wavefront.cuh)
        switch (K) {
#define SWG_CASE(k) case k: return launch_one<Lane16, 8, k, false, false, 0, 1>(grid, smem, stream, p);
            SWG_CASE(1) SWG_CASE(2) SWG_CASE(3) SWG_CASE(4) SWG_CASE(5) SWG_CASE(6) SWG_CASE(7) SWG_CASE(8) SWG_CASE(9) SWG_CASE(10) SWG_CASE(11) SWG_CASE(12) SWG_CASE(13) SWG_CASE(14) SWG_CASE(15) SWG_CASE(16) SWG_CASE(17) SWG_CASE(18) SWG_CASE(19) SWG_CASE(20) SWG_CASE(21) SWG_CASE(22) SWG_CASE(23) SWG_CASE(24) SWG_CASE(25) SWG_CASE(26) SWG_CASE(27) SWG_CASE(28) SWG_CASE(29) SWG_CASE(30) SWG_CASE(31) SWG_CASE(32)
#undef SWG_CASE
            default: return cudaErrorInvalidValue;
        }
    }
    if (p.gap_extend == 2) {            // gap-extend penalty as an immediate, any gap-open penalty (see wavefront.cuh)
        switch (K) {
#define SWG_CASE(k) case k: return launch_one<Lane16, 8, k, false, false, 0, 2>(grid, smem, stream, p);
            SWG_CASE(1) SWG_CASE(2) SWG_CASE(3) SWG_CASE(4) SWG_CASE(5) SWG_CASE(6) SWG_CASE(7) SWG_CASE(8) SWG_CASE(9) SWG_CASE(10) SWG_CASE(11) SWG_CASE(12) SWG_CASE(13) SWG_CASE(14) SWG_CASE(15) SWG_CASE(16) SWG_CASE(17) SWG_CASE(18) SWG_CASE(19) SWG_CASE(20) SWG_CASE(21) SWG_CASE(22) SWG_CASE(23) SWG_CASE(24) SWG_CASE(25) SWG_CASE(26) SWG_CASE(27) SWG_CASE(28) SWG_CASE(29) SWG_CASE(30) SWG_CASE(31) SWG_CASE(32)
#undef SWG_CASE
            default: return cudaErrorInvalidValue;
        }
    }
    switch (K) {
#define SWG_CASE(k) case k: return launch_one<Lane16, 8, k, false, false>(grid, smem, stream, p);
        SWG_CASE(1) SWG_CASE(2) SWG_CASE(3) SWG_CASE(4) SWG_CASE(5) SWG_CASE(6) SWG_CASE(7) SWG_CASE(8)
        SWG_CASE(9) SWG_CASE(10) SWG_CASE(11) SWG_CASE(12) SWG_CASE(13) SWG_CASE(14) SWG_CASE(15) SWG_CASE(16)
        SWG_CASE(17) SWG_CASE(18) SWG_CASE(19) SWG_CASE(20) SWG_CASE(21) SWG_CASE(22) SWG_CASE(23) SWG_CASE(24)
        SWG_CASE(25) SWG_CASE(26) SWG_CASE(27) SWG_CASE(28) SWG_CASE(29) SWG_CASE(30) SWG_CASE(31) SWG_CASE(32)
#undef SWG_CASE
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace swg
