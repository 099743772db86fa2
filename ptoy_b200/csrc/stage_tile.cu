// Instantiations of the tile-staged stage kernels for one (bonds, portals) combination, selected
// with -DPTOY_GROUP=0..3 so that the four groups compile in parallel (ptoy_b200/build.py).
#include "stage_tile.cuh"

#ifndef PTOY_GROUP
#error "compile with -DPTOY_GROUP=0..3"
#endif
#define PTOY_CAT2(a, b) a##b
#define PTOY_CAT(a, b) PTOY_CAT2(a, b)

namespace ptoy {

void PTOY_CAT(launch_stage_tile_g, PTOY_GROUP)(int stage, bool extras, const StageArgs& a, int n_bound, cudaStream_t s) {
  if (extras) launch_stage_tile_feat<PTOY_GROUP | kFeatExtras>(stage, a, n_bound, s);
  else launch_stage_tile_feat<PTOY_GROUP>(stage, a, n_bound, s);
}

}  // namespace ptoy
