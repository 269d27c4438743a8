#include "beam_kernel.cuh"
namespace mb {
void launch_beam_results(int ND, const BeamGroupDev& g, const StateDev& st, double* out, cudaStream_t s) {
    const unsigned nb = (unsigned)((g.nele + MB_RES_BLOCK - 1) / MB_RES_BLOCK);
    if (ND == 1) beam_results_kernel<1><<<nb, MB_RES_BLOCK, 0, s>>>(g, st, out);
    else if (ND == 2) beam_results_kernel<2><<<nb, MB_RES_BLOCK, 0, s>>>(g, st, out);
    else beam_results_kernel<3><<<nb, MB_RES_BLOCK, 0, s>>>(g, st, out);
}
}  // namespace mb
