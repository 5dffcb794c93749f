// Instantiation unit of the fused tile kernels for TileCfg<0,128,128,3,256,4> (N, G, F, K, threads, n-tiles/task; 0 = runtime).
#include "gfc_tile_kernels.cuh"
namespace gfc {
using Cfg_128_128_3 = TileCfg<0,128,128,3,256,4>;
GFC_DEFINE_TILE_LAUNCHERS(128_128_3, Cfg_128_128_3)
}  // namespace gfc
