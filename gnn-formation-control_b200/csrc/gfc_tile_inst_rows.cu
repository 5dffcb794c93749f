// Instantiation unit of the fused tile kernels for TileCfg<0,0,0,0,384,4>: runtime shapes, 12 warps — used for the
// N = 1, K = 1 "rows" plans (the tensor-core tap contractions of the workspace / CSR pipeline): 12 warps hold
// the whole [F x C] dH accumulator of cfg5 (F = 32, C = 160: 10 warp tasks) in registers.
#include "gfc_tile_kernels.cuh"
namespace gfc {
using Cfg_rows = TileCfg<0,0,0,0,384,4>;
GFC_DEFINE_TILE_LAUNCHERS(rows, Cfg_rows)
}  // namespace gfc
