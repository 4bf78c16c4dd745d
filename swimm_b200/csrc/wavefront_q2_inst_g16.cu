// wavefront_q2_inst_g16.cu -- instantiations of wavefront_q2_kernel<16, K, ...> for K = 8, 12, ..., 32 (one file per
// group size so that the families compile in parallel).
#include "wavefront_q2.cuh"

namespace swg {

cudaError_t launch_q2_g16(int K, bool cin, bool cout, int grid, cudaStream_t stream, const WfParams &p)
{
    return launch_q2_group<16>(K, cin, cout, grid, stream, p);
}

}  // namespace swg
