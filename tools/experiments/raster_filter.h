// EXPERIMENT (round 1, not used by the shipped kernels): float pre-filter of PointInTri with a rigorous margin, meant to keep
// FP64 off the common path of raster_keys_kernel.  It is exact whenever it answers (brute-forced on 5e7 triangles), and decides
// 99.9 % of the benchmark's pixel tests, but the kernel built on it (raster_keys_v4_float_filter.cuh) measured SLOWER than the
// plain FP64 phase B (render group 117 us vs 104 us at B=64): the kernel is bound by gathers, atomics and phase A, not by the
// FP64 pipe.  Kept for round 2.  Depends on the FR_* macros of 3dfacerecon_b200/csrc/raster_core.h.
// ---- float pre-filter of PointInTri -----------------------------------------------------------------
// In exact arithmetic PointInTri's u, v are ratios of signed areas: with v0 = P3-P1, v1 = P2-P1, v2 = P-P1 and
// cross(a,b) = ax*by - ay*bx:   u = cross(v2,v1)/cross(v0,v1),  v = cross(v0,v2)/cross(v0,v1).
// The filter evaluates the three areas A = cross(v0,v1), U = cross(v2,v1), V = cross(v0,v2) in float and only answers
// when every inequality (s*U > 0, s*V > 0, s*(A-U-V) > 0 with s = sign A) holds or fails by a margin of 2^-18 of the
// L1 magnitude scale T = |v0||v1| + |v2||v1| + |v0||v2| -- ~64x the worst-case float evaluation error (<= 6.2 ulp-units of
// T) -- and the triangle is not a sliver (|A| > |v0||v1|/16).  Under those two conditions the reference's double
// evaluation (dot-product form, which can cancel badly only for slivers) deviates from the exact value by < 2^-27
// relative, so its comparisons (:113-121) agree with the exact ones.  Everything else -- pixel centres on or near an
// edge, slivers, degenerate triangles, NaNs -- returns FR_FILTER_UNSURE and goes through the literal FP64 path.
// tests/test_raster_core_host.py checks "sure => identical to fr_point_in_tri" on hundreds of millions of cases.
#define FR_FILTER_OUTSIDE 0
#define FR_FILTER_INSIDE 1
#define FR_FILTER_UNSURE 2

struct FrTriFilter {
  float x1, y1;
  float v0x, v0y, v1x, v1y;
  float area, s;    // A and its sign as +-1 (0 => never sure)
  float t01;        // |v0|_1 * |v1|_1
  float n0, n1;     // L1 norms
};

FR_HD void fr_filter_setup(float x1, float y1, float x2, float y2, float x3, float y3, FrTriFilter* f) {
  f->x1 = x1;
  f->y1 = y1;
  f->v0x = FR_FSUB(x3, x1);
  f->v0y = FR_FSUB(y3, y1);
  f->v1x = FR_FSUB(x2, x1);
  f->v1y = FR_FSUB(y2, y1);
  f->n0 = fabsf(f->v0x) + fabsf(f->v0y);
  f->n1 = fabsf(f->v1x) + fabsf(f->v1y);
  f->t01 = f->n0 * f->n1;
  f->area = f->v0x * f->v1y - f->v0y * f->v1x;
  const bool solid = fabsf(f->area) > 0.0625f * f->t01;   // false for NaN / degenerate / sliver
  f->s = solid ? (f->area > 0.0f ? 1.0f : -1.0f) : 0.0f;
}

FR_HD int fr_filter_pixel(const FrTriFilter* f, int px, int py) {
  if (f->s == 0.0f) return FR_FILTER_UNSURE;
  const float v2x = FR_FSUB((float)px, f->x1), v2y = FR_FSUB((float)py, f->y1);
  const float n2 = fabsf(v2x) + fabsf(v2y);
  const float margin = 3.814697265625e-06f * (f->t01 + n2 * (f->n0 + f->n1));   // 2^-18 * T
  const float U = v2x * f->v1y - v2y * f->v1x;
  const float V = f->v0x * v2y - f->v0y * v2x;
  const float a = f->s * U, b = f->s * V, c = f->s * ((f->area - U) - V);
  if (a > margin && b > margin && c > margin) return FR_FILTER_INSIDE;
  if (a < -margin || b < -margin || c < -margin) return FR_FILTER_OUTSIDE;
  return FR_FILTER_UNSURE;
}

