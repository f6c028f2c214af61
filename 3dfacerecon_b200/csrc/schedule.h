// Work schedule of the tensor-core forward kernel (recon_f16.cuh): which clusters -- and which of their epilogue steps -- a CTA
// processes.  Pure integer arithmetic without CUDA types so that tests/host_emul/schedule_emul.cc can enumerate it on the host
// and check that every (cluster, step) of a launch is covered exactly once (tests/test_host_logic.py).
#ifndef FR_SCHEDULE_H_
#define FR_SCHEDULE_H_

#if defined(__CUDACC__)
#define FR_SCHED_HD __host__ __device__ __forceinline__
#else
#define FR_SCHED_HD inline
#endif

namespace fr {
namespace f16 {

FR_SCHED_HD int sched_min(int a, int b) { return a < b ? a : b; }

// Tile schedule of one CTA: the clusters are dealt round-robin to the gridDim.x CTAs of a batch tile.  When the last round
// is short (nclusters % gridDim.x != 0) its clusters are SPLIT by epilogue steps over several CTAs, each of which repeats the
// cluster's (cheap) tensor-core pass but projects / rasterizes only its share of the faces: the per-SM epilogue work, which
// bounds the raster flavour, then ends together instead of leaving most SMs idle for a whole cluster time.
// share_first (raster flavour): the FIRST round is shared by PAIRS of CTAs -- both stream cluster j / 2 (the second reader
// hits the lines the first one is fetching: half the HBM bytes) and each rasterizes half of its steps.  Nothing can be
// rasterized before a CTA's first cluster is complete, and with every SM pulling its own 360 KB that takes ~11 us of a
// 90 us step; halving the bytes of the first round halves that wait, and the half-size first job is still long enough to
// cover the stream of the next cluster.  Measured: 91.3 -> 89.0 us per step at 64 faces; with several batch tiles per launch the
// wait is amortised and the pairing only costs balance (4096 faces: 4.08 -> 4.14 ms), so it is used for single-tile launches only.
#ifndef FR_SHARE_FIRST_ROUND
#define FR_SHARE_FIRST_ROUND 1
#endif
struct TileWalk {
  int tile, step0, step1;          // current cluster and its epilogue steps [step0, step1)
  int g, j, nfullw, rem, split, nsteps, k, c0;
  int pool0, npool;                // dynamic pool: clusters [pool0, pool0 + npool) are not part of the static schedule
  // g CTAs share the clusters of one batch tile, this is CTA j of them (gridDim.x, blockIdx.x)
  FR_SCHED_HD TileWalk(int nclusters, int nsteps_, bool share_first, bool pool, int g_, int j_)
      : g(g_), j(j_), nsteps(nsteps_), k(-1), c0(0), pool0(nclusters), npool(0) {
    if (FR_SHARE_FIRST_ROUND && share_first && g >= 2 && nsteps >= 2 && nclusters >= g) c0 = (g + 1) / 2;   // clusters of the shared round
    const int n = nclusters - c0;
    nfullw = n / g;
    rem = n - nfullw * g;
    if (pool && nsteps >= 2 && nfullw >= 1) {
      // the last, partial round -- plus a full one when it is short -- is left to the pool (ItemWalk)
      if (rem < g / 2 && nfullw >= 2) --nfullw;
      pool0 = c0 + nfullw * g;
      npool = nclusters - pool0;
      rem = 0;
    }
    split = (rem > 0) ? sched_min(nsteps, g / rem) : 1;
    if (split < 1) split = 1;
    tile = step0 = step1 = 0;
  }
  FR_SCHED_HD bool next() {
    ++k;
    int kk = k;
    if (c0 > 0) {
      if (k == 0) {                                  // shared round: CTAs 2 t and 2 t + 1 take the two halves of cluster t
        tile = j >> 1;
        const bool alone = (j == g - 1) && (g & 1);  // odd CTA count: the last CTA has its cluster to itself
        const int half = nsteps >> 1;
        step0 = (alone || !(j & 1)) ? 0 : half;
        step1 = (alone || (j & 1)) ? nsteps : half;
        return true;
      }
      kk = k - 1;
    }
    if (kk < nfullw) {
      tile = c0 + j + kk * g;
      step0 = 0;
      step1 = nsteps;
      return true;
    }
    if (kk == nfullw && j < rem * split) {
      tile = c0 + nfullw * g + j / split;
      const int part = j - (j / split) * split;
      step0 = part * nsteps / split;
      step1 = (part + 1) * nsteps / split;
      return true;
    }
    return false;
  }
};

// Items of the dynamic pool (ItemWalk in recon_f16.cuh): pool cluster p is cut into `parts` ranges of epilogue steps; item
// i = p * parts + part, so that the parts of one cluster are drawn back to back (the second reader finds the tile in L2).
#ifndef FR_POOL_PARTS
#define FR_POOL_PARTS 2
#endif
FR_SCHED_HD int pool_parts(int nsteps) { return sched_min(FR_POOL_PARTS, nsteps); }
FR_SCHED_HD int pool_items(const TileWalk& tw) { return tw.npool * pool_parts(tw.nsteps); }
FR_SCHED_HD void pool_decode(const TileWalk& tw, int item, int* tile, int* step0, int* step1) {
  const int p = pool_parts(tw.nsteps), part = item % p;
  *tile = tw.pool0 + item / p;
  *step0 = part * tw.nsteps / p;
  *step1 = (part + 1) * tw.nsteps / p;
}

}  // namespace f16
}  // namespace fr

#endif  // FR_SCHEDULE_H_
