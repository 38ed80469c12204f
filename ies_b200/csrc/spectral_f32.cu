// Explicit instantiation of the spectral kernels for field dtype f32.
#include "spectral.cuh"
namespace ies {
template int launch_zline<float, false>(Ctx*, const void*, const void*, void*, void*, int, int, int, int, cudaStream_t);
template int launch_sline<float, false>(Ctx*, const void*, const void*, void*, void*, int, int, int, int);
template int launch_yline_update<float, false>(Ctx*, const UpdParams&, int);
}  // namespace ies
