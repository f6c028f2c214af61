// Host enumeration of the forward kernel's work schedule (3dfacerecon_b200/csrc/schedule.h): test infrastructure.
// fr_emul_schedule_cover: walks the static schedule of every CTA of a launch and every item of the dynamic pool and counts how
// often each (cluster, epilogue step) is visited.  Returns 0 when every pair is covered exactly once, else a code; stats out.
#include <cstdint>
#include <vector>

#include "../../3dfacerecon_b200/csrc/schedule.h"

extern "C" int fr_emul_schedule_cover(int nclusters, int g, int nsteps, int share_first, int pool, int* stats /* [4] */) {
  using namespace fr::f16;
  std::vector<int> hits((size_t)nclusters * nsteps, 0);
  int max_static = 0, min_static = 1 << 30, pool0 = nclusters, npool = 0;
  for (int j = 0; j < g; ++j) {
    TileWalk tw(nclusters, nsteps, share_first != 0, pool != 0, g, j);
    int steps = 0, guard = 0;
    while (tw.next()) {
      if (++guard > nclusters + 8) return 10;                         // runaway walk
      if (tw.tile < 0 || tw.tile >= nclusters || tw.step0 < 0 || tw.step1 > nsteps || tw.step0 >= tw.step1) return 11;
      if (tw.tile >= tw.pool0) return 12;                             // static item inside the pool
      for (int s = tw.step0; s < tw.step1; ++s) ++hits[(size_t)tw.tile * nsteps + s];
      steps += tw.step1 - tw.step0;
    }
    if (tw.next()) return 13;                                          // must stay exhausted
    if (steps > max_static) max_static = steps;
    if (steps < min_static) min_static = steps;
    if (j == 0) pool0 = tw.pool0, npool = tw.npool;
    else if (pool0 != tw.pool0 || npool != tw.npool) return 14;       // every CTA must see the same pool
  }
  if (!pool && npool != 0) return 15;
  if (pool0 + npool != nclusters) return 16;
  TileWalk tw0(nclusters, nsteps, share_first != 0, pool != 0, g, 0);
  const int nitems = pool_items(tw0);
  for (int i = 0; i < nitems; ++i) {
    int tile, s0, s1;
    pool_decode(tw0, i, &tile, &s0, &s1);
    if (tile < pool0 || tile >= nclusters || s0 < 0 || s1 > nsteps || s0 >= s1) return 17;
    for (int s = s0; s < s1; ++s) ++hits[(size_t)tile * nsteps + s];
  }
  for (size_t i = 0; i < hits.size(); ++i)
    if (hits[i] != 1) return hits[i] == 0 ? 1 : 2;                    // 1: a (cluster, step) nobody does, 2: done twice
  stats[0] = min_static, stats[1] = max_static, stats[2] = npool, stats[3] = nitems;
  return 0;
}
