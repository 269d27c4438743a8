#include "beam_kernel.cuh"
namespace mb { MB_INSTANTIATE_BEAM(1, false) }
