// Explicit instantiation of the spectral kernels for field dtype c128.
#include "spectral.cuh"
namespace ies {
template int launch_zline<double, true>(Ctx*, const void*, const void*, void*, void*, int, int, int, int, cudaStream_t);
template int launch_sline<double, true>(Ctx*, const void*, const void*, void*, void*, int, int, int, int);
template int launch_yline_update<double, true>(Ctx*, const UpdParams&, int);
}  // namespace ies
