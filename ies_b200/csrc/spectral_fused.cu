// Explicit instantiation of the fused single-launch SHPF half-step (shpf_fused.cuh).
#include "shpf_fused.cuh"
namespace ies {
template int launch_shpf_fused<float, false>(Ctx*, const UpdParams&, int);
template int launch_shpf_fused<double, false>(Ctx*, const UpdParams&, int);
template int launch_shpf_fused<float, true>(Ctx*, const UpdParams&, int);
template int launch_shpf_fused<double, true>(Ctx*, const UpdParams&, int);
}  // namespace ies
