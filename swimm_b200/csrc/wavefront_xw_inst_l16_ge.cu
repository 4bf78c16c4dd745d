// wavefront_xw_inst_l16_ge.cu -- instantiations of wavefront_xw_kernel<Lane16, K, 0, 1> and <Lane16, K, 0, 2>: the gap-extend
// penalty (1 or 2) as an immediate, any gap-open penalty; K = 1..32.
#include "wavefront_xw.cuh"

namespace swg {

cudaError_t launch_xw_l16_ge(int K, int grid, cudaStream_t stream, const WfParams &p)
{
    if (p.gap_extend == 1) {
        switch (K) {
#define SWG_CASE(k) case k: return launch_xw_one<Lane16, k, 0, 1>(grid, stream, p);
            SWG_CASE(1) SWG_CASE(2) SWG_CASE(3) SWG_CASE(4) SWG_CASE(5) SWG_CASE(6) SWG_CASE(7) SWG_CASE(8) SWG_CASE(9) SWG_CASE(10) SWG_CASE(11) SWG_CASE(12) SWG_CASE(13) SWG_CASE(14) SWG_CASE(15) SWG_CASE(16) SWG_CASE(17) SWG_CASE(18) SWG_CASE(19) SWG_CASE(20) SWG_CASE(21) SWG_CASE(22) SWG_CASE(23) SWG_CASE(24) SWG_CASE(25) SWG_CASE(26) SWG_CASE(27) SWG_CASE(28) SWG_CASE(29) SWG_CASE(30) SWG_CASE(31) SWG_CASE(32)
#undef SWG_CASE
            default: return cudaErrorInvalidValue;
        }
    }
    switch (K) {
#define SWG_CASE(k) case k: return launch_xw_one<Lane16, k, 0, 2>(grid, stream, p);
        SWG_CASE(1) SWG_CASE(2) SWG_CASE(3) SWG_CASE(4) SWG_CASE(5) SWG_CASE(6) SWG_CASE(7) SWG_CASE(8) SWG_CASE(9) SWG_CASE(10) SWG_CASE(11) SWG_CASE(12) SWG_CASE(13) SWG_CASE(14) SWG_CASE(15) SWG_CASE(16) SWG_CASE(17) SWG_CASE(18) SWG_CASE(19) SWG_CASE(20) SWG_CASE(21) SWG_CASE(22) SWG_CASE(23) SWG_CASE(24) SWG_CASE(25) SWG_CASE(26) SWG_CASE(27) SWG_CASE(28) SWG_CASE(29) SWG_CASE(30) SWG_CASE(31) SWG_CASE(32)
#undef SWG_CASE
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace swg
