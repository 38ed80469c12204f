// Explicit instantiation of the spectral kernels for field dtype c128.
#include "spectral.cuh"
namespace ies {
template int launch_zline<double, true>(Ctx*, const void*, const void*, void*, void*, int, int, int, int);
template int launch_xline<double, true>(Ctx*, const void*, const void*, void*, void*, int);
template int launch_yline_update<double, true>(Ctx*, const UpdParams&, int);
template int launch_shpf_fused<double, true>(Ctx*, const UpdParams&, int);
}  // namespace ies
