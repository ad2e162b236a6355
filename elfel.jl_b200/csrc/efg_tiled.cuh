// efg_tiled.cuh -- tiled fused path (placeholder while the two-pass path is brought up)
#pragma once
#include "efg_ctx.cuh"
inline void tiled_release(efg_ctx *) {}
template <class F> void tiled_symbolic(efg_ctx *) { efg_throw(EFG_ERR_INVALID, "tiled path not built yet"); }
template <class F> void tiled_numeric(efg_ctx *) { efg_throw(EFG_ERR_INVALID, "tiled path not built yet"); }
