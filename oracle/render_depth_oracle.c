/* TEST INFRASTRUCTURE ONLY -- not part of the product path.
 *
 * CPU restatement (plain C99) of the reference z-buffer op
 *   /root/reference/rendering_layer/ops_src/render_depth_op.cc
 * used by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg as the checker
 * for the CUDA path.  It is pinned against the reference itself: tests/test_oracle_vs_ref.py
 * runs it side by side with oracle/_ref/libref_render_depth.so (the unmodified reference
 * TU compiled in place) and tests/golden/ holds outputs of that reference build.
 *
 * Build: gcc -O2 -std=c99 -ffp-contract=off -fPIC -shared   (NO -march / -ffast-math: an
 * FMA-contracted build changes tri_ind on shared-edge pixels).
 *
 * Written from the semantics, not transcribed: one fused
 * per-triangle pass without the reference's static scratch arrays; every arithmetic step
 * cites the reference line whose rounding behaviour it reproduces.
 */
#include <limits.h>
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <string.h>

/* render_depth_op.h:15-16 -- the reference's function-like macros; their NaN / signed-zero
 * behaviour (second operand wins unless strictly ordered) is part of the semantics. */
#define RD_MIN(a, b) ((a) < (b) ? (a) : (b))
#define RD_MAX(a, b) ((a) > (b) ? (a) : (b))

/* (int) of a double as x86-64 cvttsd2si does it: NaN and out-of-range give INT_MIN.  The
 * reference's `(int)ceil(...)` (render_depth_op.cc:276-280) compiles to exactly that. */
static int rd_cvt_int(double d) {
  if (!(d > -2147483649.0 && d < 2147483648.0)) return INT_MIN;
  return (int)d;
}

/* render_depth_op.cc:76-122 PointInTri: IEEE double, left-to-right, no contraction. */
static int rd_point_in_tri(double px, double py, double ax, double ay, double bx, double by, double cx, double cy) {
  /* a = pt1, b = pt2, c = pt3 */
  const double v0x = cx - ax, v0y = cy - ay; /* :90-91  pt3 - pt1 */
  const double v1x = bx - ax, v1y = by - ay; /* :93-94  pt2 - pt1 */
  const double v2x = px - ax, v2y = py - ay; /* :96-97  point - pt1 */
  const double dot00 = v0x * v0x + v0y * v0y; /* :99 */
  const double dot01 = v0x * v1x + v0y * v1y; /* :100 */
  const double dot02 = v0x * v2x + v0y * v2y; /* :101 */
  const double dot11 = v1x * v1x + v1y * v1y; /* :102 */
  const double dot12 = v1x * v2x + v1y * v2y; /* :103 */
  const double den = dot00 * dot11 - dot01 * dot01; /* :106 */
  const double inv = (den == 0) ? 0 : 1 / den;      /* :105-109 zero area => inv = 0 */
  const double u = (dot11 * dot02 - dot01 * dot12) * inv; /* :111 */
  if (u < 0 || u > 1) return 0;                            /* :113-114 */
  const double v = (dot00 * dot12 - dot01 * dot02) * inv; /* :116 */
  if (v < 0 || v > 1) return 0;                            /* :118-119 */
  return u + v < 1;                                        /* :121 strict */
}

/* Forward for `batch` faces.
 *   vertex  [batch,3,nver]   tri [3,ntri] (float indices, truncated :204-206)
 *   texture [batch,tex_ch,nver], addressed with texture_batch_stride floats between faces
 *           (0 = one texture shared by all faces, as network.py:179 tiles it)
 *   depth [batch,H,W,1]  texture_image [batch,H,W,tex_ch]  normal [batch,H,W,3]  tri_ind [batch,H,W,1]
 * Returns 0, or 1 when ntri >= 10M (the reference prints and returns, :161-166). */
int fr_oracle_render_depth_forward(const float* vertex, const float* tri, const float* texture,
                                   long texture_batch_stride, int batch, int nver, int ntri, int height, int width,
                                   int tex_ch, float* depth, float* texture_image, float* normal, float* tri_ind) {
  if (ntri >= 10 * 1000 * 1000) return 1;
  const size_t npix = (size_t)height * (size_t)width;
  for (int b = 0; b < batch; ++b) {
    const float* vx = vertex + (size_t)b * 3 * (size_t)nver;
    const float* vy = vx + nver;
    const float* vz = vy + nver;
    const float* tex = texture + (size_t)b * (size_t)texture_batch_stride;
    float* d_b = depth + (size_t)b * npix;
    float* t_b = texture_image + (size_t)b * npix * (size_t)tex_ch;
    float* n_b = normal + (size_t)b * npix * 3;
    float* i_b = tri_ind + (size_t)b * npix;

    /* :182-192, :255-261 background */
    for (size_t p = 0; p < npix; ++p) {
      d_b[p] = -99999999999999; /* int64 literal -> float: -1.00000000376832e14 */
      i_b[p] = -1;
      n_b[3 * p + 0] = 0;
      n_b[3 * p + 1] = 0;
      n_b[3 * p + 2] = 0;
      for (int c = 0; c < tex_ch; ++c) t_b[(size_t)tex_ch * p + (size_t)c] = 0;
    }

    for (int i = 0; i < ntri; ++i) {
      const int p1 = (int)tri[i];                     /* :204 */
      const int p2 = (int)tri[(size_t)ntri + i];      /* :205 */
      const int p3 = (int)tri[2 * (size_t)ntri + i];  /* :206 */

      /* :208-213 xy widened to double */
      const double ax = vx[p1], ay = vy[p1];
      const double bx = vx[p2], by = vy[p2];
      const double cx = vx[p3], cy = vy[p3];

      /* :276-280 integer bounding box */
      const int x_min = rd_cvt_int(ceil(RD_MIN(RD_MIN(ax, bx), cx)));
      const int x_max = rd_cvt_int(floor(RD_MAX(RD_MAX(ax, bx), cx)));
      const int y_min = rd_cvt_int(ceil(RD_MIN(RD_MIN(ay, by), cy)));
      const int y_max = rd_cvt_int(floor(RD_MAX(RD_MAX(ay, by), cy)));
      /* :282 whole-triangle cull, no clipping */
      if (x_max < x_min || y_max < y_min || x_max > width - 1 || x_min < 0 || y_max > height - 1 || y_min < 0) continue;

      /* :217 flat depth: float adds left-to-right, float divide, then widened */
      const float hf = (vz[p1] + vz[p2] + vz[p3]) / 3.0f;
      const double h = (double)hf;

      int setup_done = 0;
      float tritex[8];
      float nrm[3];

      for (int x = x_min; x <= x_max; ++x) {
        for (int y = y_min; y <= y_max; ++y) {
          const size_t p = (size_t)y * (size_t)width + (size_t)x;
          /* :295 strict "<" in index order => max depth wins, lowest index wins ties */
          if ((double)d_b[p] < h && rd_point_in_tri((double)x, (double)y, ax, ay, bx, by, cx, cy)) {
            if (!setup_done) {
              /* :223 mean texture in float */
              for (int c = 0; c < tex_ch && c < 8; ++c)
                tritex[c] = (tex[(size_t)c * nver + p1] + tex[(size_t)c * nver + p2] + tex[(size_t)c * nver + p3]) / 3.0f;
              /* :227-236 edge differences in FLOAT, cross product in double, stored as float on write (:308) */
              const double e12x = vx[p1] - vx[p2], e12y = vy[p1] - vy[p2], e12z = vz[p1] - vz[p2];
              const double e13x = vx[p1] - vx[p3], e13y = vy[p1] - vy[p3], e13z = vz[p1] - vz[p3];
              nrm[0] = (float)(e12y * e13z - e12z * e13y);
              nrm[1] = (float)(e12z * e13x - e12x * e13z);
              nrm[2] = (float)(e12x * e13y - e12y * e13x);
              setup_done = 1;
            }
            d_b[p] = hf;                                                                      /* :297 */
            for (int c = 0; c < tex_ch && c < 8; ++c) t_b[(size_t)tex_ch * p + (size_t)c] = tritex[c]; /* :301 */
            n_b[3 * p + 0] = nrm[0];                                                          /* :308 */
            n_b[3 * p + 1] = nrm[1];
            n_b[3 * p + 2] = nrm[2];
            i_b[p] = (float)i;                                                                /* :310 */
          }
        }
      }
    }
  }
  return 0;
}

/* Backward (render_depth_op.cc:325-368): every covered pixel adds (g * 1.0f) / 3.0f to the z
 * row of its triangle's three vertices, pixels visited row-major (:346-348).
 * Two documented departures from the reference, shared with the
 * CUDA path: vertex_grad is zero-filled first (the reference accumulates into uninitialised
 * memory, :514-516) and pixels with tri_ind < 0 are skipped (the reference reads tri(k,-1), :350-353). */
int fr_oracle_render_depth_backward(const float* depth_grad, const float* tri, const float* tri_ind, int batch,
                                    int nver, int ntri, int height, int width, float* vertex_grad) {
  const size_t npix = (size_t)height * (size_t)width;
  memset(vertex_grad, 0, sizeof(float) * (size_t)batch * 3 * (size_t)nver);
  for (int b = 0; b < batch; ++b) {
    float* gz = vertex_grad + ((size_t)b * 3 + 2) * (size_t)nver;
    for (size_t p = 0; p < npix; ++p) {
      const int t = (int)tri_ind[(size_t)b * npix + p]; /* :350 */
      if (t < 0 || t >= ntri) continue;
      const float g = depth_grad[(size_t)b * npix + p]; /* :349 */
      const int p1 = (int)tri[t];
      const int p2 = (int)tri[(size_t)ntri + t];
      const int p3 = (int)tri[2 * (size_t)ntri + t];
      const float share = g * 1.0f / 3.0f; /* :361 parses as (g * 1.0f) / 3.0f */
      gz[p1] += share;                     /* :361-363 */
      gz[p2] += share;
      gz[p3] += share;
    }
  }
  return 0;
}
