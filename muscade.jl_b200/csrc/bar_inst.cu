#include "bar_kernel.cuh"
namespace mb {
void launch_bar(int ND, bool step, const BarGroupDev& g, const StateDev& st, const NewmarkDev& nm, double t, double* Ke, double* Re, double* Rp,
                unsigned long long* nanflag, unsigned long long nanbase, cudaStream_t s) {
    const unsigned nb = (unsigned)((g.nele + 127) / 128);
#define L_(ND_, ST_) bar_kernel<ND_, ST_><<<nb, 128, 0, s>>>(g, st, nm, t, Ke, Re, Rp, nanflag, nanbase)
    if (ND == 1) L_(1, false);
    else if (ND == 2) { if (step) L_(2, true); else L_(2, false); }
    else { if (step) L_(3, true); else L_(3, false); }
#undef L_
}
void launch_soil(int ND, bool step, const SoilGroupDev& g, const StateDev& st, const NewmarkDev& nm, double* Ke, double* Re, double* Rp,
                 unsigned long long* nanflag, unsigned long long nanbase, cudaStream_t s) {
    const unsigned nb = (unsigned)((g.nele + 255) / 256);
#define L_(ND_, ST_) soil_kernel<ND_, ST_><<<nb, 256, 0, s>>>(g, st, nm, Ke, Re, Rp, nanflag, nanbase)
    if (ND == 1) L_(1, false);
    else if (ND == 2) { if (step) L_(2, true); else L_(2, false); }
    else { if (step) L_(3, true); else L_(3, false); }
#undef L_
}
}  // namespace mb
