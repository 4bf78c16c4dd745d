// wavefront_q2_inst_g8.cu -- instantiations of wavefront_q2_kernel<8, K, ...> for K = 8, 12, ..., 32 (one file per
// group size so that the families compile in parallel).
#include "wavefront_q2.cuh"

namespace swg {

cudaError_t launch_q2_g8(int K, bool cin, bool cout, int grid, cudaStream_t stream, const WfParams &p)
{
    return launch_q2_group<8>(K, cin, cout, grid, stream, p);
}

}  // namespace swg
