// Instantiation unit of the fused tile kernels for TileCfg<64,128,128,4,256,4> (N, G, F, K, threads, n-tiles/task; 0 = runtime).
#include "gfc_tile_kernels.cuh"
namespace gfc {
using Cfg_n64_128_128_4 = TileCfg<64,128,128,4,256,4>;
GFC_DEFINE_TILE_LAUNCHERS(n64_128_128_4, Cfg_n64_128_128_4)
}  // namespace gfc
