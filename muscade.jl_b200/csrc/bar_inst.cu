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
    const unsigned nb = (unsigned)((g.nele + 127) / 128);      // 128 threads: soil_kernel's staging tiles are sized for four warps
#define L_(ND_, ST_) soil_kernel<ND_, ST_><<<nb, 128, 0, s>>>(g, st, nm, Ke, Re, Rp, nanflag, nanbase)
    if (ND == 1) L_(1, false);
    else if (ND == 2) { if (step) L_(2, true); else L_(2, false); }
    else { if (step) L_(3, true); else L_(3, false); }
#undef L_
}
int launch_bar_direct(int ND, const BarGroupDev& g, const DirectStateDev& st, double t, double* dR, double* R, unsigned long long* nanflag,
                      unsigned long long nanbase, cudaStream_t s, const StepBatch& sb, int nbatch) {
    const int64_t nt = g.nele * (ND + (g.udof ? 1 : 0));
    const dim3 nb((unsigned)((nt + 127) / 128), nbatch);
    if (ND == 1) bar_direct_kernel<1><<<nb, 128, 0, s>>>(g, st, t, dR, R, nanflag, nanbase, sb);
    else if (ND == 2) bar_direct_kernel<2><<<nb, 128, 0, s>>>(g, st, t, dR, R, nanflag, nanbase, sb);
    else bar_direct_kernel<3><<<nb, 128, 0, s>>>(g, st, t, dR, R, nanflag, nanbase, sb);
    return 1;
}
int launch_soil_direct(int ND, const SoilGroupDev& g, const DirectStateDev& st, const double* Lam, double lamscale, double* dR, double* R, double* GX,
                       unsigned long long* nanflag, unsigned long long nanbase, cudaStream_t s, const StepBatch& sb, int nbatch) {
    const dim3 nb((unsigned)((g.nele + 255) / 256), nbatch);
    if (ND == 1) soil_direct_kernel<1><<<nb, 256, 0, s>>>(g, st, Lam, lamscale, dR, R, GX, nanflag, nanbase, sb);
    else if (ND == 2) soil_direct_kernel<2><<<nb, 256, 0, s>>>(g, st, Lam, lamscale, dR, R, GX, nanflag, nanbase, sb);
    else soil_direct_kernel<3><<<nb, 256, 0, s>>>(g, st, Lam, lamscale, dR, R, GX, nanflag, nanbase, sb);
    return 1;
}
}  // namespace mb
