// Instantiation unit of the fused tile kernels for TileCfg<8,32,32,3,512,2> (N, G, F, K, threads, n-tiles/task; 0 = runtime).
#include "gfc_tile_kernels.cuh"
namespace gfc {
using Cfg_n8_32_32_3 = TileCfg<8,32,32,3,512,2>;
GFC_DEFINE_TILE_LAUNCHERS(n8_32_32_3, Cfg_n8_32_32_3)
}  // namespace gfc
