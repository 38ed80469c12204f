// Explicit instantiation of the spectral kernels for field dtype c64.
#include "spectral.cuh"
namespace ies {
template int launch_zline<float, true>(Ctx*, const void*, const void*, void*, void*, int, int, int, int);
template int launch_xline<float, true>(Ctx*, const void*, const void*, void*, void*, int);
template int launch_yline_update<float, true>(Ctx*, const UpdParams&, int);
template int launch_shpf_fused<float, true>(Ctx*, const UpdParams&, int);
}  // namespace ies
