// TEST INFRASTRUCTURE ONLY.  Host re-enactment of the CUDA rasterizer's data flow (raster.cuh): the same
// raster_core.h functions the kernels call, driven by plain loops in a deliberately scrambled triangle order, so the
// key packing / atomicMax visibility / resolve logic can be checked against the oracle on a machine without a GPU.
// Built by tests/test_raster_core_host.py with g++ -O2 -ffp-contract=off.
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../3dfacerecon_b200/csrc/mesh_table.h"
#include "../../3dfacerecon_b200/csrc/raster_core.h"

static bool vidx(float f, int nver, int* out) {
  if (!(f > -1.0f && f < (float)nver)) return false;
  *out = (int)f;
  return true;
}

extern "C" int fr_emul_render_forward(const float* vertex, const float* tri, const float* texture, long long tex_stride,
                                      int batch, int nver, int ntri, int height, int width, unsigned order_seed,
                                      float* depth, float* teximg, float* normal, float* tri_ind) {
  const size_t npix = (size_t)height * width;
  std::vector<unsigned long long> keys(npix);
  for (int b = 0; b < batch; ++b) {
    const float* vx = vertex + (size_t)b * 3 * nver;
    const float* vy = vx + nver;
    const float* vz = vy + nver;
    for (auto& k : keys) k = 0ull;
    // "raster_keys_kernel": any order must give the same result -> visit triangles in a scrambled order
    const unsigned long long stride = 2654435761ull % (ntri > 0 ? ntri : 1) | 1ull;
    for (int i = 0; i < ntri; ++i) {
      int t = order_seed ? (int)(((unsigned long long)i * stride + order_seed) % ntri) : i;
      // a scrambled permutation needs gcd(stride, ntri) == 1; fall back to reversed order otherwise
      if (order_seed && ntri % 2 == 0) t = ntri - 1 - i;
      int p1, p2, p3;
      if (!vidx(tri[t], nver, &p1) || !vidx(tri[ntri + t], nver, &p2) || !vidx(tri[2 * (size_t)ntri + t], nver, &p3)) continue;
      FrBBox bb;
      FrBBox bb_ref;
      const bool keep_ref = fr_tri_bbox(vx[p1], vy[p1], vx[p2], vy[p2], vx[p3], vy[p3], width, height, &bb_ref);
      const bool keep = fr_tri_bbox_fast(vx[p1], vy[p1], vx[p2], vy[p2], vx[p3], vy[p3], width, height, &bb);
      if (keep != keep_ref) return 100;  // the float fast path must agree with the literal path
      if (keep && (bb.x_min != bb_ref.x_min || bb.x_max != bb_ref.x_max || bb.y_min != bb_ref.y_min || bb.y_max != bb_ref.y_max)) return 101;
      {  // the per-vertex snapped cull (what raster_keys_kernel runs) must agree as well, except that a triangle with a
         // NaN coordinate may be culled early: it can never pass the inside test
        const uint32_t limit = (uint32_t)width | ((uint32_t)height << 16);
        uint32_t lo_min, hi_max;
        const bool keep_snap = fr_snap_keep(fr_snap_vertex(vx[p1], vy[p1], width, height), fr_snap_vertex(vx[p2], vy[p2], width, height),
                                            fr_snap_vertex(vx[p3], vy[p3], width, height), limit, &lo_min, &hi_max);
        {  // the one-word snap codes (4th word of the vertex record) must reproduce fr_snap_keep exactly
          uint32_t lo2, hi2;
          const bool keep_code = fr_code_keep(fr_snap_code(vx[p1], vy[p1], width, height), fr_snap_code(vx[p2], vy[p2], width, height),
                                              fr_snap_code(vx[p3], vy[p3], width, height), limit, &lo2, &hi2);
          if (keep_code != keep_snap) return 107;
          if (keep_code && (lo2 != lo_min || hi2 != hi_max)) return 108;
        }
        const bool has_nan = vx[p1] != vx[p1] || vx[p2] != vx[p2] || vx[p3] != vx[p3] || vy[p1] != vy[p1] || vy[p2] != vy[p2] || vy[p3] != vy[p3];
        if (has_nan) {
          if (keep_snap) return 103;
          if (keep_ref) {  // literal path kept it: make sure it really draws nothing
            FrTriEdge en;
            fr_tri_edge_setup(vx[p1], vy[p1], vx[p2], vy[p2], vx[p3], vy[p3], &en);
            for (int y = bb_ref.y_min; y <= bb_ref.y_max; ++y)
              for (int x = bb_ref.x_min; x <= bb_ref.x_max; ++x)
                if (fr_point_in_tri(&en, x, y)) return 104;
          }
          continue;
        }
        if (keep_snap != keep_ref) return 105;
        if (keep_snap) {
          FrBBox bs;
          fr_snap_bbox(lo_min, hi_max, &bs);
          if (bs.x_min != bb_ref.x_min || bs.x_max != bb_ref.x_max || bs.y_min != bb_ref.y_min || bs.y_max != bb_ref.y_max) return 106;
        }
      }
      if (!keep) continue;
      const float h = fr_tri_depth(vz[p1], vz[p2], vz[p3]);
      if (!fr_depth_draws(h)) continue;
      FrTriEdge e;
      fr_tri_edge_setup(vx[p1], vy[p1], vx[p2], vy[p2], vx[p3], vy[p3], &e);
      const unsigned long long key = fr_pack_key(h, t);
      for (int y = bb.y_min; y <= bb.y_max; ++y)
        for (int x = bb.x_min; x <= bb.x_max; ++x)
          if (fr_point_in_tri(&e, x, y)) {
            unsigned long long& k = keys[(size_t)y * width + x];
            if (key > k) k = key;  // atomicMax
          }
    }
    // "raster_resolve_kernel"
    for (size_t p = 0; p < npix; ++p) {
      const size_t o = (size_t)b * npix + p;
      union { uint32_t u; float f; } bg;
      bg.u = FR_BACKGROUND_DEPTH_BITS;
      float d = bg.f, ti = -1.0f, n[3] = {0, 0, 0}, tx[3] = {0, 0, 0};
      if (keys[p] != 0ull) {
        const int t = fr_key_triangle(keys[p]);
        const int p1 = (int)tri[t], p2 = (int)tri[ntri + t], p3 = (int)tri[2 * (size_t)ntri + t];
        d = fr_key_depth(keys[p]);
        {  // the decoded depth must equal the recomputed one bit for bit (including the sign of a zero)
          union { float f; uint32_t u; } a, r;
          a.f = d;
          r.f = fr_tri_depth(vz[p1], vz[p2], vz[p3]);
          if (a.u != r.u) return 102;
        }
        ti = (float)t;
        fr_tri_normal(vx[p1], vy[p1], vz[p1], vx[p2], vy[p2], vz[p2], vx[p3], vy[p3], vz[p3], n);
        const float* tex = texture + (size_t)b * tex_stride;
        for (int c = 0; c < 3; ++c) tx[c] = fr_tri_mean(tex[(size_t)c * nver + p1], tex[(size_t)c * nver + p2], tex[(size_t)c * nver + p3]);
      }
      depth[o] = d;
      tri_ind[o] = ti;
      for (int c = 0; c < 3; ++c) {
        normal[3 * o + c] = n[c];
        teximg[3 * o + c] = tx[c];
      }
    }
  }
  return 0;
}

// Data flow of the TILE rasterizer (raster_tile.cuh, stand-alone and inside the fused reconstruction epilogue): walk the mesh
// table cluster by cluster, "stage" the cluster's vertices with their snap codes, cull every triangle on the raw packed
// min / max of the codes of its three LOCAL slots (fr_code_nonempty), check the image range on the survivors, run the
// certified fast inside test (direct form for one-pixel boxes, plane equations for larger ones) with the literal PointInTri for
// the pixels it leaves undecided, take the packed-key maximum, then resolve depth and index from the keys alone.
// Outputs: depth and tri_ind.
extern "C" int fr_emul_render_forward_clustered(const float* vertex, const unsigned char* table, int batch, int nver, int height,
                                                int width, float* depth, float* tri_ind) {
  fr::MeshTableHeader h;
  std::memcpy(&h, table, sizeof(h));
  if (h.magic != fr::kMeshMagic || h.nver != nver) return 200;
  const int32_t* cv = reinterpret_cast<const int32_t*>(table + h.off_vert);
  const int32_t* tb = reinterpret_cast<const int32_t*>(table + h.off_tri_begin);
  const uint32_t* te = reinterpret_cast<const uint32_t*>(table + h.off_tri);
  const size_t npix = (size_t)height * width;
  const uint32_t limit = (uint32_t)width | ((uint32_t)height << 16);
  std::vector<unsigned long long> keys(npix);
  for (int b = 0; b < batch; ++b) {
    const float* vx = vertex + (size_t)b * 3 * nver;
    const float* vy = vx + nver;
    const float* vz = vy + nver;
    for (auto& k : keys) k = 0ull;
    for (int c = h.nclusters - 1; c >= 0; --c) {             // any cluster order gives the same keys
      float sx[fr::kClusterVerts], sy[fr::kClusterVerts], sz[fr::kClusterVerts];
      uint32_t code[fr::kClusterVerts];
      for (int v = 0; v < fr::kClusterVerts; ++v) {
        const int32_t raw = cv[(size_t)c * fr::kClusterVerts + v];
        const int n = raw < 0 ? -1 : (int)((uint32_t)raw & fr::kVertIdMask);
        sx[v] = n < 0 ? 0.0f : vx[n];
        sy[v] = n < 0 ? 0.0f : vy[n];
        sz[v] = n < 0 ? 0.0f : vz[n];
        code[v] = fr_snap_code(sx[v], sy[v], width, height);
      }
      for (int i = tb[c]; i < tb[c + 1]; ++i) {
        const uint32_t w = te[2 * (size_t)i];
        const int t = (int)te[2 * (size_t)i + 1];
        const unsigned l1 = w & 0xFFu, l2 = (w >> 8) & 0xFFu, l3 = (w >> 16) & 0xFFu;
        const uint32_t mn = fr_min3_u16x2(code[l1], code[l2], code[l3]), mx = fr_max3_u16x2(code[l1], code[l2], code[l3]);
        if (!fr_code_nonempty(mn, mx)) continue;                           // cull phase
        const bool single = fr_code_single(mn, mx);
        const uint32_t lo = fr_code_lo(mn), hi = single ? lo : fr_code_hi(mx);
        if (!fr_box_in_image(lo, hi, limit)) continue;                     // draw phase: image range (:282)
        const float hgt = fr_tri_depth(sz[l1], sz[l2], sz[l3]);
        if (!fr_depth_draws(hgt)) continue;
        FrTriEdge e;
        fr_tri_edge_setup(sx[l1], sy[l1], sx[l2], sy[l2], sx[l3], sy[l3], &e);   // (only used for undecided pixels)
        const unsigned long long key = fr_pack_key(hgt, t);
        FrBBox bb;
        fr_snap_bbox(lo, hi, &bb);
        const int ext = (bb.x_max - bb.x_min > bb.y_max - bb.y_min ? bb.x_max - bb.x_min : bb.y_max - bb.y_min) + 1;
        FrTriFast ff;
        FrTriPlanes pl;
        fr_fast_setup(sx[l1], sy[l1], sx[l2], sy[l2], sx[l3], sy[l3], fr_fast_tol(1), &ff);
        fr_planes_setup(sx[l1], sy[l1], sx[l2], sy[l2], sx[l3], sy[l3], bb.x_min, bb.y_min, fr_fast_tol(ext), &pl);
        for (int y = bb.y_min; y <= bb.y_max; ++y)
          for (int x = bb.x_min; x <= bb.x_max; ++x) {
            int in = single ? fr_fast_classify(&ff, x, y) : fr_planes_classify(&pl, (float)(x - bb.x_min), (float)(y - bb.y_min));
            if (in < 0) in = fr_point_in_tri(&e, x, y) ? 1 : 0;
            if (in) {
              unsigned long long& k = keys[(size_t)y * width + x];
              if (key > k) k = key;
            }
          }
      }
    }
    for (size_t p = 0; p < npix; ++p) {
      union { uint32_t u; float f; } bg;
      bg.u = FR_BACKGROUND_DEPTH_BITS;
      const size_t o = (size_t)b * npix + p;
      depth[o] = keys[p] ? fr_key_depth(keys[p]) : bg.f;
      tri_ind[o] = keys[p] ? (float)fr_key_triangle(keys[p]) : -1.0f;
    }
  }
  return 0;
}

// fr_snap_code (the short form the kernels run) against its literal definition: every `stride`-th float bit pattern
// plus the neighbourhood of every integer in [-3, extent + 3].  Returns the number of mismatches.
extern "C" long long fr_emul_check_snap_code(int width, int height, unsigned stride) {
  long long bad = 0;
  auto one = [&](float v) {
    if (fr_snap_code(v, v, width, height) != fr_snap_code_literal(v, v, width, height)) ++bad;
    if (fr_snap_code(v, 0.5f, width, height) != fr_snap_code_literal(v, 0.5f, width, height)) ++bad;
  };
  for (unsigned long long u = 0; u < (1ull << 32); u += stride) {
    union { uint32_t u; float f; } c;
    c.u = (uint32_t)u;
    one(c.f);
  }
  const int top = (width > height ? width : height) + 3;
  for (int i = -3; i <= top; ++i) {
    union { uint32_t u; float f; } c;
    c.f = (float)i;
    for (int d = -2; d <= 2; ++d) {
      union { uint32_t u; float f; } e;
      e.u = c.u + (uint32_t)d;
      one(e.f);
    }
    one((float)i + 0.5f);
  }
  return bad;
}

// Tile rasterizer arithmetic (raster_tile.cuh): the raw-min/max cull must agree with fr_code_keep / fr_code_box, and the
// certified fast inside test may only answer what the literal PointInTri answers.  Random sub-pixel triangles plus the
// adversarial families: vertices on integer / half-integer coordinates, pixel centres exactly on edges and vertices,
// slivers, degenerate and near-degenerate triangles, tiny and huge coordinates, large boxes.
// Returns a negative code on a mismatch, else the number of pixel tests the fast test decided (out[0] = all tests).
static uint64_t emul_rng(uint64_t* s) {
  *s ^= *s << 13; *s ^= *s >> 7; *s ^= *s << 17;
  return *s;
}
static float emul_unit(uint64_t* s) { return (float)(emul_rng(s) >> 40) * (1.0f / 16777216.0f); }

extern "C" long long fr_emul_check_tile_arith(int width, int height, int ntrials, unsigned seed, long long* out) {
  uint64_t st = 0x9E3779B97F4A7C15ull ^ ((uint64_t)seed << 1 | 1ull);
  const uint32_t limit = (uint32_t)width | ((uint32_t)height << 16);
  long long decided = 0, total = 0;
  for (int it = 0; it < ntrials; ++it) {
    float x[3], y[3];
    const int family = (int)(emul_rng(&st) % 10);
    const float cx = emul_unit(&st) * (float)(width + 4) - 2.0f, cy = emul_unit(&st) * (float)(height + 4) - 2.0f;
    const float scale = (family == 9) ? 40.0f : ((family == 8) ? 6.0f : 1.5f);
    for (int k = 0; k < 3; ++k) {
      x[k] = cx + (emul_unit(&st) - 0.5f) * scale;
      y[k] = cy + (emul_unit(&st) - 0.5f) * scale;
    }
    switch (family) {
      case 0: for (int k = 0; k < 3; ++k) { x[k] = floorf(x[k]); y[k] = floorf(y[k]); } break;            // integer vertices
      case 1: for (int k = 0; k < 3; ++k) { x[k] = floorf(x[k] * 2.0f) * 0.5f; y[k] = floorf(y[k] * 2.0f) * 0.5f; } break;
      case 2: x[2] = x[0] + (x[1] - x[0]) * 0.5f; y[2] = y[0] + (y[1] - y[0]) * 0.5f; break;                // (nearly) collinear
      case 3: x[1] = x[0]; y[1] = y[0]; break;                                                             // two equal vertices
      case 4: x[0] = floorf(x[0]); y[0] = floorf(y[0]); break;                                             // pt1 on a pixel centre
      case 5: y[0] = floorf(cy); y[1] = y[0]; break;                                                       // an edge through a pixel row
      case 6: x[2] = x[0] + (x[1] - x[0]) * 0.5f + 1e-6f; y[2] = y[0] + (y[1] - y[0]) * 0.5f; break;        // sliver
      default: break;
    }
    const uint32_t e1 = fr_snap_code(x[0], y[0], width, height), e2 = fr_snap_code(x[1], y[1], width, height),
                   e3 = fr_snap_code(x[2], y[2], width, height);
    uint32_t lo, hi;
    const bool keep = fr_code_keep(e1, e2, e3, limit, &lo, &hi);
    const uint32_t mn = fr_min3_u16x2(e1, e2, e3), mx = fr_max3_u16x2(e1, e2, e3);
    uint32_t lo_ref, hi_ref;
    fr_code_box(e1, e2, e3, &lo_ref, &hi_ref);
    const bool nonempty_ref = (lo_ref & 0xFFFFu) <= (hi_ref & 0xFFFFu) && (lo_ref >> 16) <= (hi_ref >> 16);
    if (fr_code_nonempty(mn, mx) != nonempty_ref) return -1;
    if (nonempty_ref) {
      if (fr_code_lo(mn) != lo_ref || fr_code_hi(mx) != hi_ref) return -2;
      if (fr_code_single(mn, mx) != (lo_ref == hi_ref)) return -3;
      if ((fr_code_nonempty(mn, mx) && fr_box_in_image(fr_code_lo(mn), fr_code_hi(mx), limit)) != keep) return -4;
    } else if (keep) {
      return -5;
    }
    if (!keep) continue;
    FrBBox bb;
    fr_snap_bbox(lo, hi, &bb);
    FrTriEdge e;
    fr_tri_edge_setup(x[0], y[0], x[1], y[1], x[2], y[2], &e);
    const int ext = (bb.x_max - bb.x_min > bb.y_max - bb.y_min ? bb.x_max - bb.x_min : bb.y_max - bb.y_min) + 1;
    FrTriFast ff;
    fr_fast_setup(x[0], y[0], x[1], y[1], x[2], y[2], fr_fast_tol(ext), &ff);
    FrTriPlanes pl;                                     // the plane-equation form the multi-pixel path runs
    fr_planes_setup(x[0], y[0], x[1], y[1], x[2], y[2], bb.x_min, bb.y_min, fr_fast_tol(ext), &pl);
    for (int py = bb.y_min; py <= bb.y_max; ++py)
      for (int px = bb.x_min; px <= bb.x_max; ++px) {
        const int fast = fr_fast_classify(&ff, px, py);
        const int plane = fr_planes_classify(&pl, (float)(px - bb.x_min), (float)(py - bb.y_min));
        const bool exact = fr_point_in_tri(&e, px, py);
        ++total;
        if (fast >= 0) {
          ++decided;
          if ((fast == 1) != exact) return -10 - family;
        }
        if (plane >= 0 && (plane == 1) != exact) return -30 - family;
        if (family >= 7 && plane < 0 && fast >= 0 && ext <= 2) return -50;   // generic small boxes: both forms decide alike
      }
  }
  if (out) out[0] = total;
  return decided;
}
