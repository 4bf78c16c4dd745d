// wavefront_q2_inst_g32_middle.cu -- instantiations of wavefront_q2_kernel<32, K, true, true> for K = 8, 10, ..., 32: the middle passes of a multi-pass pair
// (one file per family so that the families compile in parallel).
#include "wavefront_q2.cuh"

namespace swg {

cudaError_t launch_q2_g32_middle(int K, int grid, cudaStream_t stream, const WfParams &p)
{
    return launch_q2_family<32, true, true>(K, grid, stream, p);
}

}  // namespace swg
