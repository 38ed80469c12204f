// Explicit instantiation of the spectral kernels for field dtype f64.
#include "spectral.cuh"
namespace ies {
template int launch_zline<double, false>(Ctx*, const void*, const void*, void*, void*, int, int, int, int, cudaStream_t);
template int launch_sline<double, false>(Ctx*, const void*, const void*, void*, void*, int, int, int, int);
template int launch_yline_update<double, false>(Ctx*, const UpdParams&, int);
}  // namespace ies
