// wavefront_q2_inst_g16.cu -- instantiations of wavefront_q2_kernel<16, K, false, false> for K = 8, 10, ..., 32: single-pass pairs, 16-thread groups
// (one file per family so that the families compile in parallel).
#include "wavefront_q2.cuh"

namespace swg {

cudaError_t launch_q2_g16(int K, int grid, cudaStream_t stream, const WfParams &p)
{
    return launch_q2_family<16, false, false>(K, grid, stream, p);
}

}  // namespace swg
