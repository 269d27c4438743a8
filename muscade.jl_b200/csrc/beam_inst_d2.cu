#include "beam_kernel.cuh"
namespace mb { MB_INSTANTIATE_BEAM_DIRECT(2) }
