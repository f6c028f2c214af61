// Per-triangle / per-pixel arithmetic of the z-buffer rasterizer, shared by the CUDA kernels
// (raster.cuh) and by a host build that tests/test_raster_core_host.py drives on the CPU to
// check this logic against the oracle without a GPU (test infrastructure; the product path is
// the CUDA build only).
//
// Semantics: /root/reference/rendering_layer/ops_src/render_depth_op.cc:76-122 (PointInTri),
// :201-246 (per-triangle setup), :263-316 (raster loop) -- see SURVEY.md App. A.3.
// Bit-exactness contract: every double operation below is a separately rounded IEEE operation
// in the reference's order (the reference is built without FMA contraction, ops.py:51), so on
// the device they are spelled with __dmul_rn/__dadd_rn/... which the compiler may not fuse.
#ifndef FR_RASTER_CORE_H_
#define FR_RASTER_CORE_H_

#include <limits.h>
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define FR_HD __host__ __device__ __forceinline__
#else
#define FR_HD static inline
#endif

#if defined(__CUDA_ARCH__)
#define FR_DMUL(a, b) __dmul_rn((a), (b))
#define FR_DADD(a, b) __dadd_rn((a), (b))
#define FR_DSUB(a, b) __dsub_rn((a), (b))
#define FR_DDIV(a, b) __ddiv_rn((a), (b))
#define FR_FADD(a, b) __fadd_rn((a), (b))
#define FR_FSUB(a, b) __fsub_rn((a), (b))
#define FR_FDIV(a, b) __fdiv_rn((a), (b))
#else  // host build: compiled with -ffp-contract=off
#define FR_DMUL(a, b) ((a) * (b))
#define FR_DADD(a, b) ((a) + (b))
#define FR_DSUB(a, b) ((a) - (b))
#define FR_DDIV(a, b) ((a) / (b))
#define FR_FADD(a, b) ((a) + (b))
#define FR_FSUB(a, b) ((a) - (b))
#define FR_FDIV(a, b) ((a) / (b))
#endif

// render_depth_op.h:15-16: the reference's min/max are these macros; keep their NaN behaviour.
#define FR_RMIN(a, b) ((a) < (b) ? (a) : (b))
#define FR_RMAX(a, b) ((a) > (b) ? (a) : (b))

// Bit pattern of (float)(-99999999999999), the reference's background depth (render_depth_op.cc:187).
#define FR_BACKGROUND_DEPTH_BITS 0xD6B5E621u

// (int)double as the reference binary does it (x86-64 cvttsd2si): NaN / out of range -> INT_MIN.
FR_HD int fr_cvt_int_x86(double d) {
  if (!(d > -2147483649.0 && d < 2147483648.0)) return INT_MIN;
  return (int)d;
}

struct FrBBox {
  int x_min, x_max, y_min, y_max;
};

// Integer bounding box (:276-280) and whole-triangle cull (:282).  Returns false when culled.
FR_HD bool fr_tri_bbox(float x1, float y1, float x2, float y2, float x3, float y3, int width, int height, FrBBox* bb) {
  const double ax = x1, ay = y1, bx = x2, by = y2, cx = x3, cy = y3;
  const double lox = FR_RMIN(FR_RMIN(ax, bx), cx);
  const double hix = FR_RMAX(FR_RMAX(ax, bx), cx);
  const double loy = FR_RMIN(FR_RMIN(ay, by), cy);
  const double hiy = FR_RMAX(FR_RMAX(ay, by), cy);
  bb->x_min = fr_cvt_int_x86(ceil(lox));
  bb->x_max = fr_cvt_int_x86(floor(hix));
  bb->y_min = fr_cvt_int_x86(ceil(loy));
  bb->y_max = fr_cvt_int_x86(floor(hiy));
  if (bb->x_max < bb->x_min || bb->y_max < bb->y_min || bb->x_max > width - 1 || bb->x_min < 0 ||
      bb->y_max > height - 1 || bb->y_min < 0)
    return false;
  return true;
}

// IEEE x / 3.0f without the division sequence (and without its slow-path call): q0 = x*c, q = fma(fma(-3,q0,x), c, q0) with
// c = RN(1/3) is the correctly rounded quotient for EVERY finite non-zero float, denormals included (exhaustively verified
// over all 2^32 bit patterns, on the host with a correctly rounded fmaf and on the device by tools/div3_check.cu); +-0,
// +-inf and NaN are their own quotients: x + x reproduces them (sign of zero kept, NaN quieted like the division does).
FR_HD float fr_div3(float x) {
#if defined(__CUDA_ARCH__)
  if ((__float_as_uint(x) & 0x7FFFFFFFu) - 1u < 0x7F7FFFFFu) {      // finite and non-zero
    const float c = 0.3333333432674407958984375f;
    const float q0 = __fmul_rn(x, c);
    return __fmaf_rn(__fmaf_rn(-3.0f, q0, x), c, q0);
  }
  return __fadd_rn(x, x);
#else
  return FR_FDIV(x, 3.0f);
#endif
}

// Flat depth of a triangle (:217): float adds left to right, IEEE float divide by 3.0f.
FR_HD float fr_tri_depth(float z1, float z2, float z3) { return fr_div3(FR_FADD(FR_FADD(z1, z2), z3)); }

// Mean of a per-vertex attribute (:223), same float arithmetic.
FR_HD float fr_tri_mean(float a, float b, float c) { return fr_div3(FR_FADD(FR_FADD(a, b), c)); }

// Edge-function state of one triangle for PointInTri (:76-109): everything that does not depend on the pixel.
struct FrTriEdge {
  double ax, ay;        // pt1
  double v0x, v0y;      // pt3 - pt1
  double v1x, v1y;      // pt2 - pt1
  double dot00, dot01, dot11;
  double inv;           // 0 when the triangle has zero area (:105-109)
};

FR_HD void fr_tri_edge_setup(float x1, float y1, float x2, float y2, float x3, float y3, FrTriEdge* e) {
  e->ax = x1;
  e->ay = y1;
  e->v0x = FR_DSUB((double)x3, e->ax);
  e->v0y = FR_DSUB((double)y3, e->ay);
  e->v1x = FR_DSUB((double)x2, e->ax);
  e->v1y = FR_DSUB((double)y2, e->ay);
  e->dot00 = FR_DADD(FR_DMUL(e->v0x, e->v0x), FR_DMUL(e->v0y, e->v0y));
  e->dot01 = FR_DADD(FR_DMUL(e->v0x, e->v1x), FR_DMUL(e->v0y, e->v1y));
  e->dot11 = FR_DADD(FR_DMUL(e->v1x, e->v1x), FR_DMUL(e->v1y, e->v1y));
  const double den = FR_DSUB(FR_DMUL(e->dot00, e->dot11), FR_DMUL(e->dot01, e->dot01));
  e->inv = (den == 0) ? 0.0 : FR_DDIV(1.0, den);
}

// PointInTri for pixel centre (px,py) (:96-121).  Edges through pt1 inclusive, edge pt2-pt3 exclusive.
FR_HD bool fr_point_in_tri(const FrTriEdge* e, int px, int py) {
  const double v2x = FR_DSUB((double)px, e->ax);
  const double v2y = FR_DSUB((double)py, e->ay);
  const double dot02 = FR_DADD(FR_DMUL(e->v0x, v2x), FR_DMUL(e->v0y, v2y));
  const double dot12 = FR_DADD(FR_DMUL(e->v1x, v2x), FR_DMUL(e->v1y, v2y));
  const double u = FR_DMUL(FR_DSUB(FR_DMUL(e->dot11, dot02), FR_DMUL(e->dot01, dot12)), e->inv);
  if (u < 0 || u > 1) return false;
  const double v = FR_DMUL(FR_DSUB(FR_DMUL(e->dot00, dot12), FR_DMUL(e->dot01, dot02)), e->inv);
  if (v < 0 || v > 1) return false;
  return FR_DADD(u, v) < 1;
}

// Face normal (:227-236): edge differences in FLOAT, cross product in double, rounded to float on write (:308).
FR_HD void fr_tri_normal(float x1, float y1, float z1, float x2, float y2, float z2, float x3, float y3, float z3,
                         float n[3]) {
  const double e12x = FR_FSUB(x1, x2), e12y = FR_FSUB(y1, y2), e12z = FR_FSUB(z1, z2);
  const double e13x = FR_FSUB(x1, x3), e13y = FR_FSUB(y1, y3), e13z = FR_FSUB(z1, z3);
  n[0] = (float)FR_DSUB(FR_DMUL(e12y, e13z), FR_DMUL(e12z, e13y));
  n[1] = (float)FR_DSUB(FR_DMUL(e12z, e13x), FR_DMUL(e12x, e13z));
  n[2] = (float)FR_DSUB(FR_DMUL(e12x, e13y), FR_DMUL(e12y, e13x));
}

// ---- visibility key ------------------------------------------------------------------------------
// The reference visits triangles in index order and overwrites a pixel iff depth < h (:295), so the
// winner is (max h, then min index).  Packed as one u64 so a single atomicMax resolves it in any order:
//   high 32 bits: order-preserving map of the float depth (with -0.0 folded onto +0.0, which compare equal)
//   low  32 bits: (0x7FFFFFFF - triangle index) << 1 | (h is -0.0)       smaller index = larger key
// The last bit keeps the sign of a zero depth, which the folded high word cannot (it never decides a comparison:
// two keys with equal high words and equal index are the same triangle), so depth AND index decode from the
// key alone.  0 is reserved for "background": a triangle only draws when h > background depth (NaN never
// draws), and every such h maps to a non-zero high word.
FR_HD uint32_t fr_float_order_bits(float h) {
  union { float f; uint32_t u; } c;
  c.f = h;
  if ((c.u & 0x7FFFFFFFu) == 0u) c.u = 0u;  // -0.0 -> +0.0
  return (c.u & 0x80000000u) ? ~c.u : (c.u | 0x80000000u);
}

FR_HD bool fr_depth_draws(float h) {
  union { uint32_t u; float f; } bg;
  bg.u = FR_BACKGROUND_DEPTH_BITS;
  return bg.f < h;  // false for NaN and for anything at or below the background depth
}

FR_HD unsigned long long fr_pack_key(float h, int tri_index) {
  union { float f; uint32_t u; } c;
  c.f = h;
  const uint32_t low = ((0x7FFFFFFFu - (uint32_t)tri_index) << 1) | ((c.u == 0x80000000u) ? 1u : 0u);
  return ((unsigned long long)fr_float_order_bits(h) << 32) | (unsigned long long)low;
}

FR_HD int fr_key_triangle(unsigned long long key) { return (int)(0x7FFFFFFFu - ((uint32_t)(key & 0xFFFFFFFFull) >> 1)); }

// Depth stored in a (non-zero) key, bit for bit (including the sign of a zero).
FR_HD float fr_key_depth(unsigned long long key) {
  const uint32_t o = (uint32_t)(key >> 32);
  union { uint32_t u; float f; } c;
  c.u = (o & 0x80000000u) ? (o & 0x7FFFFFFFu) : ~o;
  if (c.u == 0u && (key & 1ull)) c.u = 0x80000000u;
  return c.f;
}

// Same result as fr_tri_bbox for every input, but in float arithmetic when no coordinate is NaN: min/max/ceil/floor
// of float values are exact in float, and comparing the float bounds against [0, W-1] culls exactly the cases where
// the reference's (int) conversion leaves the image or overflows to INT_MIN.  NaNs take the literal path.
FR_HD bool fr_tri_bbox_fast(float x1, float y1, float x2, float y2, float x3, float y3, int width, int height, FrBBox* bb) {
  const float sx = (x1 + x2) + x3, sy = (y1 + y2) + y3;   // NaN iff a coordinate is NaN (or inf - inf: also slow path)
  if (sx != sx || sy != sy) return fr_tri_bbox(x1, y1, x2, y2, x3, y3, width, height, bb);
  const float lox = ceilf(fminf(fminf(x1, x2), x3)), hix = floorf(fmaxf(fmaxf(x1, x2), x3));
  const float loy = ceilf(fminf(fminf(y1, y2), y3)), hiy = floorf(fmaxf(fmaxf(y1, y2), y3));
  if (hix < lox || hiy < loy || hix > (float)(width - 1) || lox < 0.0f || hiy > (float)(height - 1) || loy < 0.0f) return false;
  bb->x_min = (int)lox;
  bb->x_max = (int)hix;
  bb->y_min = (int)loy;
  bb->y_max = (int)hiy;
  return true;
}

// ---- per-vertex pixel snapping -------------------------------------------------------------------
// ceil() and floor() are monotone, so the reference's integer bounding box (:276-280) can be formed from per-vertex
// values: x_min = min_i ceil(x_i), x_max = max_i floor(x_i).  A vertex is snapped ONCE per face to four small
// integers, biased by +1 and clamped to [0, W+1] (resp. H+1), packed as 16-bit fields:
//   lo = ceil-biased  (x | y << 16),   hi = floor-biased (x | y << 16)
// The clamp keeps the cull decision (:282) exact: a clamped value only occurs when some x_i <= -1 or x_i >= W, and
// then the triangle is culled either way.  A NaN coordinate snaps to ceil-biased 0, which forces the cull -- the
// reference never draws such a triangle either (every PointInTri comparison with NaN is false).
struct FrSnap {
  uint32_t lo, hi;
};

FR_HD uint32_t fr_snap_axis(float v, int extent, bool want_ceil) {
  float r = want_ceil ? ceilf(v) : floorf(v);
  r = fminf(fmaxf(r, -1.0f), (float)extent);  // NaN -> -1
  if (v != v) r = -1.0f;
  return (uint32_t)((int)r + 1);
}

FR_HD FrSnap fr_snap_vertex(float x, float y, int width, int height) {
  FrSnap s;
  s.lo = fr_snap_axis(x, width, true) | (fr_snap_axis(y, height, true) << 16);
  s.hi = fr_snap_axis(x, width, false) | (fr_snap_axis(y, height, false) << 16);
  return s;
}

FR_HD uint32_t fr_min3_u16x2(uint32_t a, uint32_t b, uint32_t c) {
#if defined(__CUDA_ARCH__)
  return __vminu2(__vminu2(a, b), c);
#else
  uint32_t l = a & 0xFFFFu, h = a >> 16;
  if ((b & 0xFFFFu) < l) l = b & 0xFFFFu;
  if ((c & 0xFFFFu) < l) l = c & 0xFFFFu;
  if ((b >> 16) < h) h = b >> 16;
  if ((c >> 16) < h) h = c >> 16;
  return l | (h << 16);
#endif
}
FR_HD uint32_t fr_max3_u16x2(uint32_t a, uint32_t b, uint32_t c) {
#if defined(__CUDA_ARCH__)
  return __vmaxu2(__vmaxu2(a, b), c);
#else
  uint32_t l = a & 0xFFFFu, h = a >> 16;
  if ((b & 0xFFFFu) > l) l = b & 0xFFFFu;
  if ((c & 0xFFFFu) > l) l = c & 0xFFFFu;
  if ((b >> 16) > h) h = b >> 16;
  if ((c >> 16) > h) h = c >> 16;
  return l | (h << 16);
#endif
}

// Cull test on snapped vertices; `limit` = width | height << 16.  On success *lo_min / *hi_max hold the biased
// bounding box (x_min+1 | y_min+1 << 16, x_max+1 | y_max+1 << 16).  Fields are < 2^15, so bit 15 / 31 serve as
// borrow guards: (a | G) - b keeps its guard bit iff a >= b, field by field, in one 32-bit subtraction.
FR_HD bool fr_snap_keep(FrSnap a, FrSnap b, FrSnap c, uint32_t limit, uint32_t* lo_min, uint32_t* hi_max) {
  const uint32_t G = 0x80008000u;
  const uint32_t lo = fr_min3_u16x2(a.lo, b.lo, c.lo);
  const uint32_t hi = fr_max3_u16x2(a.hi, b.hi, c.hi);
  const uint32_t nonempty = (hi | G) - lo;          // x_max >= x_min, y_max >= y_min
  const uint32_t inside_lo = (lo | G) - 0x00010001u; // x_min >= 0, y_min >= 0   (biased >= 1)
  const uint32_t inside_hi = (limit | G) - hi;       // x_max <= W-1, y_max <= H-1 (biased <= W)
  *lo_min = lo;
  *hi_max = hi;
  return (nonempty & inside_lo & inside_hi & G) == G;
}

// ---- one-word snap code --------------------------------------------------------------------------
// floor-biased is always ceil-biased or ceil-biased - 1 (also after the clamp), so one 16-bit field per axis carries
// both:  e = 2 * ceil_biased + (floor_biased == ceil_biased).  e is monotone in the pair (ceil, floor), hence
//   min_i ceil_i = (min_i e_i) >> 1      and      max_i floor_i = floor of (max_i e_i),
// so the cull needs one packed min3 and one packed max3 on the raw codes and a single decode per triangle.
// Fields stay below 2^16 because biased values are < 2^15.  The code is the 4th word of the 16-byte vertex record the
// rasterizer gathers (x, y, z, code).
FR_HD uint32_t fr_snap_code_literal(float x, float y, int width, int height) {   // the definition, spelled out
  const FrSnap s = fr_snap_vertex(x, y, width, height);
  const uint32_t same = ~(s.lo ^ s.hi);                                        // per field: all ones iff equal
  const uint32_t fx = ((same & 0xFFFFu) == 0xFFFFu) ? 1u : 0u, fy = ((same >> 16) == 0xFFFFu) ? 1u : 0u;
  return (s.lo << 1) | fx | (fy << 16);
}

// The same code in a handful of operations per axis.  Inside (-1, extent): floor_biased = floor(v) + 1 and
// ceil_biased = floor_biased + (v is not an integer), so e = 2 floor(v) + 4 - (v is an integer); at or below -1 (and for
// NaN) both clamp to 0 -> e = 1; at or above extent both clamp to extent + 1 -> e = 2 extent + 3.
// tests/host_emul checks it against fr_snap_code_literal over a sweep of float bit patterns.
FR_HD uint32_t fr_snap_code_axis(float v, int extent) {
  // clamp first: at -1 (also NaN: fmaxf returns the other operand) floor = -1 and "integer" give e = 1, at extent e = 2 extent + 3
#if defined(__CUDA_ARCH__)
  const float c = fminf(fmaxf(v, -1.0f), (float)extent);        // (CUDA's fmaxf returns the non-NaN operand)
#else
  const float lo = (v > -1.0f) ? v : -1.0f;                      // NaN -> -1, spelled out: host fmaxf may be lowered to maxss
  const float c = (lo < (float)extent) ? lo : (float)extent;
#endif
  const float t = floorf(c);
  return (uint32_t)(2 * (int)t + 3) + ((t != c) ? 1u : 0u);
}

FR_HD uint32_t fr_snap_code(float x, float y, int width, int height) {
  return fr_snap_code_axis(x, width) | (fr_snap_code_axis(y, height) << 16);
}

// Biased bounding box of a triangle from the snap codes of its three vertices
// (x_min+1 | y_min+1 << 16, x_max+1 | y_max+1 << 16); meaningful for triangles fr_code_keep keeps.
FR_HD void fr_code_box(uint32_t e1, uint32_t e2, uint32_t e3, uint32_t* lo_min, uint32_t* hi_max) {
  const uint32_t mn = fr_min3_u16x2(e1, e2, e3);
  const uint32_t mx = fr_max3_u16x2(e1, e2, e3);
  *lo_min = (mn >> 1) & 0x7FFF7FFFu;
  *hi_max = ((mx >> 1) & 0x7FFF7FFFu) - 0x00010001u + (mx & 0x00010001u);  // flag clear => ceil_biased >= 1: no borrow
}

// fr_snap_keep on snap codes: same decision, same biased bounding box.
FR_HD bool fr_code_keep(uint32_t e1, uint32_t e2, uint32_t e3, uint32_t limit, uint32_t* lo_min, uint32_t* hi_max) {
  const uint32_t G = 0x80008000u;
  uint32_t lo, hi;
  fr_code_box(e1, e2, e3, &lo, &hi);
  const uint32_t nonempty = (hi | G) - lo;
  const uint32_t inside_lo = (lo | G) - 0x00010001u;
  const uint32_t inside_hi = (limit | G) - hi;
  *lo_min = lo;
  *hi_max = hi;
  return (nonempty & inside_lo & inside_hi & G) == G;
}

// ---- cull on the raw packed min / max of the codes (tile rasterizer, raster_tile.cuh) ---------------
// mn / mx = per-field minimum / maximum of the three snap codes (one VIMNMX3.U16x2 each).  With e = 2 ceil_b + s
// (s = 1 when the coordinate is an integer): floor_b(max) + 1 = ((mx + 1) & ~1) / 2 and ceil_b(min) = mn / 2, so the
// integer box is non-empty on an axis iff ((mx + 1) & ~1) - mn >= 1.  Needs code fields < 2^15 (extent <= 16000): the
// field differences then stay below 2^15 and adding 0x7FFF turns ">= 1" into bit 15 of every field.
// The image-range half of the reference's cull (:282) is NOT part of this test (see fr_box_in_image).
FR_HD uint32_t fr_code_even_up(uint32_t mx) { return (mx + 0x00010001u) & 0xFFFEFFFEu; }   // per field 2 (floor_b(max) + 1)
FR_HD bool fr_code_nonempty(uint32_t mn, uint32_t mx) {
  return (((fr_code_even_up(mx) - mn) + 0x7FFF7FFFu) & 0x80008000u) == 0x80008000u;
}
// The box is exactly one pixel: floor_b(max) == ceil_b(min) on both axes (implies non-empty).
FR_HD bool fr_code_single(uint32_t mn, uint32_t mx) { return (((fr_code_even_up(mx) - 0x00020002u) ^ mn) & 0xFFFEFFFEu) == 0u; }
// Biased box corners (same values as fr_code_box): lo = x_min+1 | y_min+1 << 16, hi = x_max+1 | y_max+1 << 16.
FR_HD uint32_t fr_code_lo(uint32_t mn) { return (mn >> 1) & 0x7FFF7FFFu; }
FR_HD uint32_t fr_code_hi(uint32_t mx) { return ((fr_code_even_up(mx) >> 1) & 0x7FFF7FFFu) - 0x00010001u; }
// x_min >= 0, y_min >= 0, x_max <= W-1, y_max <= H-1 on the biased corners; limit = W | H << 16.
FR_HD bool fr_box_in_image(uint32_t lo, uint32_t hi, uint32_t limit) {
  const uint32_t G = 0x80008000u;
  return ((((lo | G) - 0x00010001u) & ((limit | G) - hi)) & G) == G;
}

// ---- certified fast inside test ------------------------------------------------------------------
// PointInTri (:76-121) decides  u >= 0, v >= 0, u + v < 1  on u = (dot11 dot02 - dot01 dot12) inv, v = ..., ten rounded
// double operations per pixel plus a division per triangle.  In exact arithmetic (Lagrange's identity)
//     u den = c10 cu,   v den = c01 cv,   den = c10^2      c10 = v1 x v0,  cu = v1 x v2,  cv = v0 x v2  (2-D cross products)
// so with s = sign(c10), C = |c10|:  u = s cu / C,  v = -s cv / C,  1 - u - v = (C - s cu + s cv) / C.
// The fast test evaluates the three numerators  cu' = s cu,  cv' = -s cv,  cw' = C - cu' - cv'  in FLOAT (a dozen
// operations per pixel, no conversion to double, no division) and only answers when the reference's rounded double
// computation provably gives the same answer; everything else is "undecided" and takes the literal fr_point_in_tri.
//
// Let Lb >= every |component| of v0, v1, v2: the integer box is w x h pixels and all three vertices lie within one pixel
// of it, so Lb = max(w, h) + 1 works; 2^e >= Lb^2 below.
//  (1) The reference.  Its numerators / denominator carry an absolute error below E = 64 eps Lb^4 (eps = 2^-53: two rounded
//      dot products of magnitude <= 2 Lb^2 with error <= 5 eps Lb^2 each, their rounded products, the rounded difference).
//      If C > 2^-20 Lb^2 then den_ref > 0.99 C^2 > 0, and
//        cu' >  2^-24 Lb^2  =>  num_u,ref >= C cu' - E > 0  =>  u_ref > 0        cu' < -2^-24 Lb^2  =>  u_ref < 0   (same for v)
//        cw' >  2^-24 Lb^2  =>  u_ref + v_ref < 1 - 2 eps  (the rounded sum is < 1, hence also u_ref, v_ref <= 1)
//        cw' < -2^-24 Lb^2  =>  u_ref + v_ref >= 1
//      (the last two need C cw' > 3 E + 12.2 eps C Lb^2, and 3 E / C < 3 2^-27 Lb^2).
//  (2) The float evaluation.  Every float difference (v0, v1, v2 components) has relative error <= 2^-24; a cross product
//      a b - c d (one rounded product, one fused or rounded multiply-subtract) then has absolute error < 7.1 2^-24 Lb^2
//      < 2^-21 Lb^2, and cw' (two more subtractions) < 34 2^-24 Lb^2 < 2^-18.9 Lb^2.
//  So  |cu'_float| > 2^-17 2^e,  |cv'_float| > ...,  |cw'_float| > 2^-17 2^e  and  C_float > 2^-16 2^e  certify (1) with room to
//  spare (thresholds are powers of two and compared on the float's bit pattern).  For sub-pixel triangles this leaves
//  pixel centres within ~3e-5 px of an edge, slivers of area < 3e-5 px^2 and degenerate triangles (den == 0 paints the
//  box, :105-109) to the literal path.  Coordinates are finite and inside (-1, 16000) here (the cull guarantees it), so
//  nothing overflows; float underflow only makes a numerator 0 = undecided.
struct FrTriFast {
  float ax, ay;         // pt1
  float v0x, v0y;       // pt3 - pt1  (the reference's v0)
  float v1x, v1y;       // pt2 - pt1
  float cabs;           // |c10|
  uint32_t sigma;       // sign bit of c10
  int32_t tol;          // bit pattern of the decision threshold 2^(e-17)
  bool ok;              // C > 2^(e-16)
};

FR_HD uint32_t fr_fbits(float f) {
#if defined(__CUDA_ARCH__)
  return __float_as_uint(f);
#else
  union { float f; uint32_t u; } c;
  c.f = f;
  return c.u;
#endif
}
FR_HD float fr_fxor(float f, uint32_t mask) {      // flips bits of the pattern (sign changes)
#if defined(__CUDA_ARCH__)
  return __uint_as_float(__float_as_uint(f) ^ mask);
#else
  union { float f; uint32_t u; } c;
  c.f = f;
  c.u ^= mask;
  return c.f;
#endif
}

// Threshold pattern for a box of box_extent = max(w, h) pixels: 2^(e-17) with 2^e >= (box_extent + 1)^2.
FR_HD int32_t fr_fast_tol(int box_extent) {
  const uint32_t n = (uint32_t)box_extent + 1u, n2m1 = n * n - 1u;   // e = bit length of Lb^2 - 1
#if defined(__CUDA_ARCH__)
  const int e = 32 - __clz((int)n2m1);
#else
  const int e = n2m1 ? 32 - __builtin_clz(n2m1) : 0;
#endif
  return (int32_t)((uint32_t)(127 - 17 + e) << 23);
}

FR_HD void fr_fast_setup(float x1, float y1, float x2, float y2, float x3, float y3, int32_t tol, FrTriFast* f) {
  f->ax = x1;
  f->ay = y1;
  f->v0x = x3 - x1;
  f->v0y = y3 - y1;
  f->v1x = x2 - x1;
  f->v1y = y2 - y1;
  const float c10 = f->v1x * f->v0y - f->v1y * f->v0x;      // (rounded or fused: both covered by the margins)
  const uint32_t bits = fr_fbits(c10);
  f->sigma = bits & 0x80000000u;
  f->cabs = fr_fxor(c10, f->sigma);
  f->tol = tol;
  f->ok = (int32_t)(bits & 0x7FFFFFFFu) > tol + (1 << 23);
}

// 1 = inside, 0 = outside, -1 = undecided (run fr_point_in_tri).
FR_HD int fr_fast_classify(const FrTriFast* f, int px, int py) {
  const float v2x = (float)px - f->ax, v2y = (float)py - f->ay;
  const float cu = f->v1x * v2y - f->v1y * v2x;
  const float cv = f->v0x * v2y - f->v0y * v2x;
  const float cup = fr_fxor(cu, f->sigma), cvp = fr_fxor(cv, f->sigma ^ 0x80000000u);
  const float cwp = (f->cabs - cup) - cvp;
  const int32_t hu = (int32_t)fr_fbits(cup), hv = (int32_t)fr_fbits(cvp), hw = (int32_t)fr_fbits(cwp);
#if defined(__CUDA_ARCH__)
  const int32_t mn = __vimin3_s32(hu, hv, hw);
  const uint32_t mx = __vimax3_u32((uint32_t)hu, (uint32_t)hv, (uint32_t)hw);
#else
  const int32_t mn = hu < hv ? (hu < hw ? hu : hw) : (hv < hw ? hv : hw);
  const uint32_t a = (uint32_t)hu, b = (uint32_t)hv, c = (uint32_t)hw;
  const uint32_t mx = a > b ? (a > c ? a : c) : (b > c ? b : c);
#endif
  if (!f->ok) return -1;
  if (mn > f->tol) return 1;                                   // all three numerators positive and above the threshold
  if (mx > (0x80000000u | (uint32_t)f->tol)) return 0;         // one of them negative and below minus the threshold
  return -1;
}

// The same test for the pixels of a larger box: the three numerators are linear in the pixel, so they are evaluated as
// plane equations around the box origin (x0, y0):  c(x0 + dx, y0 + dy) = c0 + dx cx + dy cy  (two fused multiply-adds each,
// dx / dy small exact integers).  Error budget: c0 as in fr_fast_classify (< 2^-21 Lb^2, cw' < 2^-18.9 Lb^2), the slopes are
// the edge components themselves (cw's: one float addition, 2^-24 relative), each fma rounds once (<= 2^-23 Lb^2): the
// totals stay below 2^-20.4 Lb^2 resp. 2^-18.6 Lb^2, inside the same thresholds.
struct FrTriPlanes {
  float u0, ux, uy;     // cu'
  float v0, vx, vy;     // cv'
  float w0, wx, wy;     // cw' = C - cu' - cv'
  int32_t tol;
  bool ok;
};

FR_HD void fr_planes_setup(float x1, float y1, float x2, float y2, float x3, float y3, int x0, int y0, int32_t tol, FrTriPlanes* p) {
  const float v0x = x3 - x1, v0y = y3 - y1, v1x = x2 - x1, v1y = y2 - y1;
  const float c10 = v1x * v0y - v1y * v0x;
  const uint32_t bits = fr_fbits(c10), su = bits & 0x80000000u, sv = su ^ 0x80000000u;
  const float v2x = (float)x0 - x1, v2y = (float)y0 - y1;
  p->u0 = fr_fxor(v1x * v2y - v1y * v2x, su);      // cu = v1x (y - ay) - v1y (x - ax):  d/dx = -v1y,  d/dy = v1x
  p->ux = fr_fxor(v1y, sv);
  p->uy = fr_fxor(v1x, su);
  p->v0 = fr_fxor(v0x * v2y - v0y * v2x, sv);      // cv' = -s cv
  p->vx = fr_fxor(v0y, su);
  p->vy = fr_fxor(v0x, sv);
  p->w0 = (fr_fxor(c10, su) - p->u0) - p->v0;
  p->wx = -(p->ux + p->vx);
  p->wy = -(p->uy + p->vy);
  p->tol = tol;
  p->ok = (int32_t)(bits & 0x7FFFFFFFu) > tol + (1 << 23);
}

// dx, dy: pixel offsets from the box origin as floats.  1 = inside, 0 = outside, -1 = undecided.
FR_HD int fr_planes_classify(const FrTriPlanes* p, float dx, float dy) {
#if defined(__CUDA_ARCH__)
  const float cu = __fmaf_rn(dy, p->uy, __fmaf_rn(dx, p->ux, p->u0));
  const float cv = __fmaf_rn(dy, p->vy, __fmaf_rn(dx, p->vx, p->v0));
  const float cw = __fmaf_rn(dy, p->wy, __fmaf_rn(dx, p->wx, p->w0));
#else
  const float cu = fmaf(dy, p->uy, fmaf(dx, p->ux, p->u0));
  const float cv = fmaf(dy, p->vy, fmaf(dx, p->vx, p->v0));
  const float cw = fmaf(dy, p->wy, fmaf(dx, p->wx, p->w0));
#endif
  const int32_t hu = (int32_t)fr_fbits(cu), hv = (int32_t)fr_fbits(cv), hw = (int32_t)fr_fbits(cw);
#if defined(__CUDA_ARCH__)
  const int32_t mn = __vimin3_s32(hu, hv, hw);
  const uint32_t mx = __vimax3_u32((uint32_t)hu, (uint32_t)hv, (uint32_t)hw);
#else
  const int32_t mn = hu < hv ? (hu < hw ? hu : hw) : (hv < hw ? hv : hw);
  const uint32_t a = (uint32_t)hu, b = (uint32_t)hv, c = (uint32_t)hw;
  const uint32_t mx = a > b ? (a > c ? a : c) : (b > c ? b : c);
#endif
  if (!p->ok) return -1;
  if (mn > p->tol) return 1;
  if (mx > (0x80000000u | (uint32_t)p->tol)) return 0;
  return -1;
}

FR_HD void fr_snap_bbox(uint32_t lo_min, uint32_t hi_max, FrBBox* bb) {
  bb->x_min = (int)(lo_min & 0xFFFFu) - 1;
  bb->y_min = (int)(lo_min >> 16) - 1;
  bb->x_max = (int)(hi_max & 0xFFFFu) - 1;
  bb->y_max = (int)(hi_max >> 16) - 1;
}

#endif  // FR_RASTER_CORE_H_
