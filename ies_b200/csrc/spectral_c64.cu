// Explicit instantiation of the spectral kernels for field dtype c64.
#include "spectral.cuh"
namespace ies {
template int launch_zline<float, true>(Ctx*, const void*, const void*, void*, void*, int, int, int, int, cudaStream_t);
template int launch_sline<float, true>(Ctx*, const void*, const void*, void*, void*, int, int, int, int);
template int launch_yline_update<float, true>(Ctx*, const UpdParams&, int);
}  // namespace ies
