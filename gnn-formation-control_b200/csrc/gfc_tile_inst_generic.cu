// Instantiation unit of the fused tile kernels for TileCfg<0,0,0,0,256,4> (N, G, F, K, threads, n-tiles/task; 0 = runtime).
#include "gfc_tile_kernels.cuh"
namespace gfc {
using Cfg_generic = TileCfg<0,0,0,0,256,4>;
GFC_DEFINE_TILE_LAUNCHERS(generic, Cfg_generic)
}  // namespace gfc
