// wavefront_q2_inst_g32_first.cu -- instantiations of wavefront_q2_kernel<32, K, false, true> for K = 8, 10, ..., 32: the first pass of a multi-pass pair
// (one file per family so that the families compile in parallel).
#include "wavefront_q2.cuh"

namespace swg {

cudaError_t launch_q2_g32_first(int K, int grid, cudaStream_t stream, const WfParams &p)
{
    return launch_q2_family<32, false, true>(K, grid, stream, p);
}

}  // namespace swg
