// Per-mesh side table of the rasterizer: the triangle list partitioned into CLUSTERS of at most 128 unique vertices
// and at most 256 triangles, with pre-validated 8-bit local vertex indices.
//
// The reference rasterizes triangle by triangle with three float -> int index conversions and nine scattered
// vertex loads each (render_depth_op.cc:204-213, render_depth_op.cu.cc:84-92).  Here a cluster's vertices are staged
// in shared memory ONCE per face (by the reconstruction epilogue that has just computed them, or by one coalesced
// gather from the vertex tensor) and every triangle of the cluster is culled / tested from there: a vertex is
// fetched once per face instead of once per incident triangle, and the inner loop has no index validation.
//
// A cluster doubles as the row tile of the tensor-core reconstruction (recon_f16.cuh: M = 128 TMEM lanes), so the
// packed basis stores its forward operand tiles in cluster order; a vertex on the border of several clusters is a
// member of each of them (its basis rows are stored once per member cluster), and exactly one of them OWNS it
// (writes it to the planar vertex_proj tensor).
//
// This header is host-only C++ (the builder runs once per mesh, at model load) plus the plain-old-data table format
// both sides agree on.  Nothing here is taken from the reference: it has no such structure.
#ifndef FR_MESH_TABLE_H_
#define FR_MESH_TABLE_H_

#include <stdint.h>

namespace fr {

constexpr int kClusterVerts = 128;   // == kTileVerts (TMEM lanes of one reconstruction tile)
constexpr int kClusterTris = 256;    // local triangle ids fit 8 bits
constexpr uint32_t kMeshMagic = 0x544D5246u;   // "FRMT"
constexpr uint32_t kMeshVersion = 4u;
constexpr uint32_t kVertOwner = 0x40000000u;   // cluster_vert flag: this cluster writes the vertex to planar outputs
constexpr uint32_t kVertIdMask = 0x00FFFFFFu;  // nver <= 2^24 (float triangle indices are exact up to there)

// Table layout (one contiguous blob, identical on host and device; all offsets in bytes from the start):
//   MeshTableHeader
//   cluster_vert  int32 [nclusters][128]   vertex id | kVertOwner, -1 = unused slot
//   tri_begin     int32 [nclusters + 1]    first triangle entry of each cluster
//   tri_entry     uint2 [ntri_slots]       { l1 | l2 << 8 | l3 << 16 (slots within the cluster), original triangle index }
//   tri_vid       uint4 [ntri_slots]       { r1, r2, r3 (vertex RANKS, pre-validated), original triangle index }: the same
//                                          triangles in the same (cluster) order for kernels that gather vertex records
//   rank_vert     int32 [ceil(nver/128)*128]  vertex id | kVertOwner of every rank, -1 behind the last one
//   vert_rank     int32 [nver]             rank of every vertex
//   cluster_rank  int32 [nclusters][128]   rank of the vertex in every cluster slot (-1 = unused slot): where the tile
//                                          rasterizer (raster_tile.cuh) finds the slot's 16-byte record; starts at
//                                          mesh_off_cluster_rank(header) (derived: the 64-byte header is full)
//   tri_rank4     uint4 [ntri]             by ORIGINAL triangle index: { r1, r2, r3 (vertex ranks), 1 } or all zero for a
//                                          triangle that was dropped: one 16-byte gather resolves a winner's vertices in the
//                                          resolve pass; starts at mesh_off_tri_rank4(header)
// RANK = the position of a vertex in cluster order (clusters in table order, within a cluster its owned vertices in slot
// order).  Kernels that keep per-vertex data in global memory (the 16-byte vertex records of raster.cuh, the row order of
// the packed basis) use ranks instead of the mesh's own numbering: the vertices a block of triangles touches are then
// neighbours in memory whatever the numbering of the mesh file -- locality comes from this table, not from the generator.
struct MeshTableHeader {
  uint32_t magic, version;
  int32_t nver, ntri;
  int32_t nclusters;
  int32_t ntri_slots;          // valid triangles (entries); triangles with an index outside [0, nver) are dropped
  int32_t max_cluster_tris;
  int32_t nvert_slots;         // used cluster_vert slots (vertices counted once per member cluster)
  uint32_t off_vert, off_tri_begin, off_tri, total_bytes;
  uint32_t hash;               // FNV-1a of everything behind the header
  uint32_t off_tri_vid, off_rank_vert, off_vert_rank;
};
static_assert(sizeof(MeshTableHeader) == 64, "header is 64 bytes");
#if defined(__CUDACC__)
__host__ __device__
#endif
inline uint32_t mesh_off_cluster_rank(const MeshTableHeader& h) { return (h.off_vert_rank + (uint32_t)h.nver * 4u + 15u) / 16u * 16u; }
#if defined(__CUDACC__)
__host__ __device__
#endif
inline uint32_t mesh_off_tri_rank4(const MeshTableHeader& h) { return mesh_off_cluster_rank(h) + (uint32_t)h.nclusters * 128u * 4u; }

}  // namespace fr

// ---- host-side builder
#include <algorithm>
#include <cmath>
#include <cstring>
#include <utility>
#include <vector>

namespace fr {

// float triangle index -> int the way the reference does ((int)tri(k,i), render_depth_op.cc:204-206), rejecting
// anything that would index outside [0, nver)  (same rule as tri_vertex_index in raster.cuh).
inline bool mesh_index(float f, int nver, int* out) {
  if (!(f > -1.0f && f < (float)nver)) return false;
  *out = (int)f;
  return true;
}

class MeshTableBuilder {
 public:
  // tri [3][ntri] float; pos: [3][nver] planar (x block, y block, z block), or interleaved [nver][3], or null
  MeshTableBuilder(const float* tri, int ntri, int nver, const float* pos, bool pos_interleaved)
      : nver_(nver), ntri_(ntri) {
    tv_.reserve((size_t)3 * ntri);
    for (int t = 0; t < ntri; ++t) {
      int a, b, c;
      if (mesh_index(tri[t], nver, &a) && mesh_index(tri[(size_t)ntri + t], nver, &b) &&
          mesh_index(tri[2 * (size_t)ntri + t], nver, &c)) {
        tv_.push_back(a);
        tv_.push_back(b);
        tv_.push_back(c);
        orig_.push_back(t);
      }
    }
    nvalid_ = (int)orig_.size();
    coord_.assign((size_t)3 * nver, 0.0f);
    if (pos != nullptr) {
      for (int n = 0; n < nver; ++n)
        for (int c = 0; c < 3; ++c) {
          const float v = pos_interleaved ? pos[(size_t)3 * n + c] : pos[(size_t)c * nver + n];
          coord_[(size_t)3 * n + c] = std::isfinite(v) ? v : 0.0f;
        }
    } else {
      landmark_coords();
    }
    cen_.resize((size_t)3 * nvalid_);
    for (int i = 0; i < nvalid_; ++i)
      for (int c = 0; c < 3; ++c)
        cen_[(size_t)3 * i + c] = (coord_[(size_t)3 * tv_[3 * i] + c] + coord_[(size_t)3 * tv_[3 * i + 1] + c] +
                                   coord_[(size_t)3 * tv_[3 * i + 2] + c]) * (1.0f / 3.0f);
    stamp_.assign(nver, -1);
  }

  // Partition and serialise.  Tries a few granularities of the bisection / slicing hand-over and keeps the
  // partition with the fewest clusters (every cluster costs one 128-row pass over the basis).
  std::vector<unsigned char> build() {
    std::vector<int> best_perm, best_cuts;
    const int pmax_candidates[] = {8, 32, 128, 1 << 20};
    for (int pmax : pmax_candidates) {
      perm_.resize(nvalid_);
      for (int i = 0; i < nvalid_; ++i) perm_[i] = i;
      cuts_.clear();
      cuts_.push_back(0);
      pmax_ = pmax;
      if (nvalid_ > 0) split(0, nvalid_);
      if (best_cuts.empty() || cuts_.size() < best_cuts.size()) {
        best_perm = perm_;
        best_cuts = cuts_;
      }
      if (nvalid_ <= kClusterTris) break;   // tiny meshes: nothing to tune
    }
    perm_.swap(best_perm);
    cuts_.swap(best_cuts);
    return serialise();
  }

 private:
  int nver_, ntri_, nvalid_ = 0, pmax_ = 4, stamp_gen_ = 0;
  std::vector<int> tv_, orig_, perm_, cuts_, stamp_;
  std::vector<float> coord_, cen_;

  // No positions given: breadth-first distance fields over the vertex graph serve as coordinates (from a peripheral
  // vertex a, from the vertex b farthest from a, and from the vertex farthest from both) -- enough for compact parts.
  void landmark_coords() {
    std::vector<int> first(nver_ + 1, 0);
    for (size_t i = 0; i < tv_.size(); ++i) first[tv_[i] + 1] += 2;
    for (int n = 0; n < nver_; ++n) first[n + 1] += first[n];
    std::vector<int> adj(first[nver_]), fill(first.begin(), first.end() - 1);
    for (int t = 0; t < nvalid_; ++t)
      for (int k = 0; k < 3; ++k) {
        const int a = tv_[3 * t + k];
        adj[fill[a]++] = tv_[3 * t + (k + 1) % 3];
        adj[fill[a]++] = tv_[3 * t + (k + 2) % 3];
      }
    std::vector<int> dist(nver_), queue(nver_);
    // distances from `start`; further connected components are appended behind it (offset, so that they stay apart).
    // Returns the vertex of start's own component reached last.
    auto bfs = [&](int start) {
      std::fill(dist.begin(), dist.end(), -1);
      int head = 0, tail = 0, last = start, offset = 0;
      for (int seed = -1; seed < nver_; ++seed) {
        const int s = seed < 0 ? start : seed;
        if (dist[s] >= 0) continue;
        dist[s] = offset;
        queue[tail++] = s;
        while (head < tail) {
          const int a = queue[head++];
          for (int e = first[a]; e < first[a + 1]; ++e)
            if (dist[adj[e]] < 0) {
              dist[adj[e]] = dist[a] + 1;
              queue[tail++] = adj[e];
            }
        }
        if (seed < 0) last = queue[tail - 1];
        offset = dist[queue[tail - 1]] + 64;
      }
      return last;
    };
    if (nver_ == 0) return;
    const int a = bfs(nvalid_ > 0 ? tv_[0] : 0);
    const int b = bfs(a);
    for (int n = 0; n < nver_; ++n) coord_[(size_t)3 * n] = (float)dist[n];
    bfs(b);
    int c = b;
    float best = -1.0f;
    for (int n = 0; n < nver_; ++n) {
      coord_[(size_t)3 * n + 1] = (float)dist[n];
      const float s = coord_[(size_t)3 * n] + coord_[(size_t)3 * n + 1];
      if (s > best) {
        best = s;
        c = n;
      }
    }
    bfs(c);
    for (int n = 0; n < nver_; ++n) coord_[(size_t)3 * n + 2] = (float)dist[n];
  }

  int count_unique(int lo, int hi) {
    ++stamp_gen_;
    int nv = 0;
    for (int i = lo; i < hi; ++i)
      for (int k = 0; k < 3; ++k) {
        const int v = tv_[3 * perm_[i] + k];
        if (stamp_[v] != stamp_gen_) {
          stamp_[v] = stamp_gen_;
          ++nv;
        }
      }
    return nv;
  }
  bool fits(int lo, int hi) { return hi - lo <= kClusterTris && count_unique(lo, hi) <= kClusterVerts; }

  // Z-order (Morton) key of a point within the bounding box of a vertex set, over the two axes of largest extent.
  struct ZOrder {
    float lo[2], scale[2];
    int axis[2];
    ZOrder(const float* coord, const std::vector<int>& verts) {
      float mn[3] = {1e30f, 1e30f, 1e30f}, mx[3] = {-1e30f, -1e30f, -1e30f};
      for (int n : verts)
        for (int c = 0; c < 3; ++c) {
          mn[c] = std::min(mn[c], coord[(size_t)3 * n + c]);
          mx[c] = std::max(mx[c], coord[(size_t)3 * n + c]);
        }
      int order[3] = {0, 1, 2};
      std::sort(order, order + 3, [&](int a, int b) { return mx[a] - mn[a] > mx[b] - mn[b]; });
      for (int k = 0; k < 2; ++k) {
        axis[k] = order[k];
        lo[k] = verts.empty() ? 0.0f : mn[order[k]];
        const float ext = verts.empty() ? 0.0f : mx[order[k]] - mn[order[k]];
        scale[k] = ext > 0.0f ? 255.0f / ext : 0.0f;
      }
    }
    uint32_t key(const float* p) const {
      uint32_t q[2];
      for (int k = 0; k < 2; ++k) {
        const float f = (p[axis[k]] - lo[k]) * scale[k];
        q[k] = (uint32_t)std::min(255.0f, std::max(0.0f, f));
      }
      uint32_t m = 0;
      for (int b = 0; b < 8; ++b) m |= ((q[0] >> b) & 1u) << (2 * b) | ((q[1] >> b) & 1u) << (2 * b + 1);
      return m;
    }
  };

  struct AxisLess {
    const float* cen;
    int axis;
    bool operator()(int a, int b) const {
      const float ca = cen[(size_t)3 * a + axis], cb = cen[(size_t)3 * b + axis];
      return ca < cb || (ca == cb && a < b);
    }
  };

  // Cut position near `k` in perm_[lo, hi) (sorted along `axis`): if the sorted centroid coordinates have a pronounced
  // gap close by (structured meshes: the boundary between two rows of triangles), cut there, so that the boundary
  // between the two sides is a straight line of vertices instead of a saw-tooth.
  int snap_cut(int lo, int hi, int k, int axis, int window) const {
    const int a = std::max(lo + 1, k - window), b = std::min(hi - 1, k + window);
    if (b <= a) return k;
    float max_gap = 0.0f, sum = 0.0f;
    for (int j = a; j <= b; ++j) {
      const float gap = cen_[(size_t)3 * perm_[j] + axis] - cen_[(size_t)3 * perm_[j - 1] + axis];
      sum += gap;
      max_gap = std::max(max_gap, gap);
    }
    const float mean = sum / (float)(b - a + 1);
    if (!(max_gap > 3.0f * mean)) return k;     // no structure (irregular mesh): keep the balanced cut
    const float thresh = std::max(3.0f * mean, 0.5f * max_gap);
    int best = k, best_d = 1 << 30;
    for (int j = a; j <= b; ++j) {               // the nearest pronounced gap
      const float gap = cen_[(size_t)3 * perm_[j] + axis] - cen_[(size_t)3 * perm_[j - 1] + axis];
      if (gap >= thresh && std::abs(j - k) < best_d) {
        best_d = std::abs(j - k);
        best = j;
      }
    }
    return best;
  }

  // Cuts perm_[lo, hi) (sorted along `axis`) into m slices of (nearly) equal triangle counts, snapped to gaps.
  void slice_positions(int lo, int hi, int m, int axis, std::vector<int>* pos) const {
    const int nt = hi - lo;
    pos->clear();
    pos->push_back(lo);
    for (int s = 1; s < m; ++s) {
      int k = lo + (int)((long long)nt * s / m);
      k = snap_cut(lo, hi, k, axis, std::max(2, nt / (5 * m)));
      if (k > pos->back()) pos->push_back(k);
    }
    pos->push_back(hi);
  }

  // Smallest number of slices along `axis` (>= m0) such that every slice fits a cluster; perm_[lo, hi) gets sorted.
  // Returns 0 if more than `mmax` slices would be needed.
  int best_slicing(int lo, int hi, int axis, int m0, int mmax, std::vector<int>* pos) {
    std::sort(perm_.begin() + lo, perm_.begin() + hi, AxisLess{cen_.data(), axis});
    for (int m = std::max(1, m0); m <= mmax; ++m) {
      slice_positions(lo, hi, m, axis, pos);
      bool ok = true;
      for (size_t s = 0; s + 1 < pos->size() && ok; ++s) ok = fits((*pos)[s], (*pos)[s + 1]);
      if (ok) return (int)pos->size() - 1;
    }
    return 0;
  }

  void axes_by_extent(int lo, int hi, int axes[3]) const {
    float mn[3] = {1e30f, 1e30f, 1e30f}, mx[3] = {-1e30f, -1e30f, -1e30f};
    for (int i = lo; i < hi; ++i)
      for (int c = 0; c < 3; ++c) {
        const float v = cen_[(size_t)3 * perm_[i] + c];
        mn[c] = std::min(mn[c], v);
        mx[c] = std::max(mx[c], v);
      }
    axes[0] = 0; axes[1] = 1; axes[2] = 2;
    std::sort(axes, axes + 3, [&](int a, int b) { return mx[a] - mn[a] > mx[b] - mn[b]; });
  }

  // Recursive coordinate bisection in proportion to the estimated number of clusters on each side, down to parts of
  // at most pmax_ clusters; such a part is tiled in two dimensions: bands across its longest axis, every band cut
  // into the smallest number of equal-count slices along the second axis that all fit a cluster.  The number of bands
  // is the one that gives the fewest clusters.  (No half-empty remainder clusters, and the clusters of one band carry
  // equal numbers of triangles.)
  void split(int lo, int hi) {
    const int nt = hi - lo;
    const int nv = count_unique(lo, hi);
    if (nt <= kClusterTris && nv <= kClusterVerts) {
      cuts_.push_back(hi);
      return;
    }
    const int m_est = std::max(2, std::max((nt + 209) / 210, (nv + kClusterVerts - 1) / kClusterVerts));
    int axes[3];
    axes_by_extent(lo, hi, axes);
    if (m_est > pmax_) {
      const int ktarget = lo + (int)((long long)nt * (m_est / 2) / m_est);
      std::sort(perm_.begin() + lo, perm_.begin() + hi, AxisLess{cen_.data(), axes[0]});
      const int k = snap_cut(lo, hi, ktarget, axes[0], std::max(2, nt / (4 * m_est)));
      split(lo, k);
      split(k, hi);
      return;
    }
    // 2-D tiling: try 1 .. bands_max bands
    const int bands_max = std::max(1, (int)std::ceil(std::sqrt(2.0 * m_est)) + 1);
    const int mmax = 3 * m_est + 8;
    int best_total = 0, best_bands = 0;
    std::vector<int> bpos, spos;
    for (int nb = 1; nb <= bands_max; ++nb) {
      std::sort(perm_.begin() + lo, perm_.begin() + hi, AxisLess{cen_.data(), axes[0]});
      slice_positions(lo, hi, nb, axes[0], &bpos);
      int total = 0;
      for (size_t b = 0; b + 1 < bpos.size(); ++b) {
        const int m = best_slicing(bpos[b], bpos[b + 1], axes[1], 1, mmax, &spos);
        if (m == 0) { total = 0; break; }
        total += m;
      }
      if (total > 0 && (best_total == 0 || total < best_total)) {
        best_total = total;
        best_bands = nb;
      }
    }
    if (best_total == 0) {                       // badly shaped part (e.g. a triangle soup): bisect by count and retry
      std::sort(perm_.begin() + lo, perm_.begin() + hi, AxisLess{cen_.data(), axes[0]});
      const int k = lo + nt / 2;
      if (k == lo || k == hi) {                  // cannot happen: a single triangle always fits
        cuts_.push_back(hi);
        return;
      }
      split(lo, k);
      split(k, hi);
      return;
    }
    std::sort(perm_.begin() + lo, perm_.begin() + hi, AxisLess{cen_.data(), axes[0]});
    slice_positions(lo, hi, best_bands, axes[0], &bpos);
    for (size_t b = 0; b + 1 < bpos.size(); ++b) {
      best_slicing(bpos[b], bpos[b + 1], axes[1], 1, mmax, &spos);
      for (size_t s = 1; s < spos.size(); ++s) cuts_.push_back(spos[s]);
    }
  }

  std::vector<unsigned char> serialise() {
    const int ncl_tri = (int)cuts_.size() - 1;
    // cluster vertex lists: in vertex-id order when the mesh numbering is already local (the cluster's ids span a narrow
    // range: planar tensors in the mesh's numbering are then written / read in runs), else in Z order of the vertex
    // positions within the cluster -- neighbouring slots / ranks are neighbouring vertices whatever the numbering
    std::vector<std::vector<int>> cverts(ncl_tri);
    std::vector<char> owned(nver_, 0);
    std::vector<std::pair<uint32_t, int>> keyed;
    for (int c = 0; c < ncl_tri; ++c) {
      std::vector<int>& v = cverts[c];
      for (int i = cuts_[c]; i < cuts_[c + 1]; ++i)
        for (int k = 0; k < 3; ++k) v.push_back(tv_[3 * perm_[i] + k]);
      std::sort(v.begin(), v.end());
      v.erase(std::unique(v.begin(), v.end()), v.end());
      if (!v.empty() && (long long)(v.back() - v.front()) > 64ll * (long long)v.size()) {
        const ZOrder z(coord_.data(), v);
        keyed.clear();
        for (int n : v) keyed.emplace_back(z.key(&coord_[(size_t)3 * n]), n);
        std::sort(keyed.begin(), keyed.end());
        for (size_t i = 0; i < v.size(); ++i) v[i] = keyed[i].second;
      }
    }
    // vertices no triangle references still have to be reconstructed (planar vertex_proj): they fill free slots
    std::vector<char> referenced(nver_, 0);
    for (size_t i = 0; i < tv_.size(); ++i) referenced[tv_[i]] = 1;
    std::vector<int> loose;
    for (int n = 0; n < nver_; ++n)
      if (!referenced[n]) loose.push_back(n);
    std::vector<std::vector<int>> extra(ncl_tri);
    size_t li = 0;
    for (int c = 0; c < ncl_tri && li < loose.size(); ++c)
      while ((int)(cverts[c].size() + extra[c].size()) < kClusterVerts && li < loose.size()) extra[c].push_back(loose[li++]);
    const int ncl_extra = (int)((loose.size() - li + kClusterVerts - 1) / kClusterVerts);
    const int ncl = ncl_tri + ncl_extra;

    MeshTableHeader h;
    std::memset(&h, 0, sizeof(h));
    h.magic = kMeshMagic;
    h.version = kMeshVersion;
    h.nver = nver_;
    h.ntri = ntri_;
    h.nclusters = ncl;
    h.ntri_slots = nvalid_;
    h.off_vert = sizeof(MeshTableHeader);
    h.off_tri_begin = h.off_vert + (uint32_t)ncl * kClusterVerts * 4u;
    h.off_tri = (h.off_tri_begin + (uint32_t)(ncl + 1) * 4u + 15u) / 16u * 16u;
    h.off_tri_vid = (h.off_tri + (uint32_t)nvalid_ * 8u + 15u) / 16u * 16u;
    const uint32_t nrank = (uint32_t)((nver_ + kClusterVerts - 1) / kClusterVerts * kClusterVerts);
    h.off_rank_vert = h.off_tri_vid + (uint32_t)nvalid_ * 16u;
    h.off_vert_rank = h.off_rank_vert + nrank * 4u;
    h.total_bytes = (mesh_off_tri_rank4(h) + (uint32_t)ntri_ * 16u + 255u) / 256u * 256u;
    std::vector<unsigned char> blob(h.total_bytes, 0);
    int32_t* cv = reinterpret_cast<int32_t*>(blob.data() + h.off_vert);
    int32_t* tb = reinterpret_cast<int32_t*>(blob.data() + h.off_tri_begin);
    uint32_t* te = reinterpret_cast<uint32_t*>(blob.data() + h.off_tri);
    uint32_t* tq = reinterpret_cast<uint32_t*>(blob.data() + h.off_tri_vid);
    std::fill(cv, cv + (size_t)ncl * kClusterVerts, -1);
    std::vector<int> slot(nver_, -1);
    int nslots = 0, max_tris = 0;
    for (int c = 0; c < ncl_tri; ++c) {
      int32_t* row = cv + (size_t)c * kClusterVerts;
      int s = 0;
      for (int v : cverts[c]) {
        slot[v] = s;
        row[s++] = v | (owned[v] ? 0 : (int32_t)kVertOwner);
        owned[v] = 1;
      }
      for (int v : extra[c]) {
        row[s++] = v | (int32_t)kVertOwner;
        owned[v] = 1;
      }
      nslots += s;
      tb[c] = cuts_[c];
      // triangles of a cluster in the order of their slots (lowest slot first; the order is irrelevant for the result:
      // visibility is resolved by (depth, index) keys; it keeps neighbouring lanes on neighbouring vertices)
      std::vector<int> ts(perm_.begin() + cuts_[c], perm_.begin() + cuts_[c + 1]);
      {
        keyed.clear();
        for (int t : ts) {
          const int a = slot[tv_[3 * t]], b = slot[tv_[3 * t + 1]], d = slot[tv_[3 * t + 2]];
          keyed.emplace_back((uint32_t)std::min(a, std::min(b, d)) * 1024u + (uint32_t)(a + b + d), t);
        }
        std::sort(keyed.begin(), keyed.end(), [&](const std::pair<uint32_t, int>& a, const std::pair<uint32_t, int>& b) {
          return a.first < b.first || (a.first == b.first && orig_[a.second] < orig_[b.second]);
        });
        for (size_t i = 0; i < ts.size(); ++i) ts[i] = keyed[i].second;
      }
      for (size_t i = 0; i < ts.size(); ++i) {
        const int t = ts[i];
        te[2 * ((size_t)cuts_[c] + i)] = (uint32_t)slot[tv_[3 * t]] | ((uint32_t)slot[tv_[3 * t + 1]] << 8) | ((uint32_t)slot[tv_[3 * t + 2]] << 16);
        te[2 * ((size_t)cuts_[c] + i) + 1] = (uint32_t)orig_[t];
        uint32_t* q4 = tq + 4 * ((size_t)cuts_[c] + i);      // vertex ids for now, ranks once they are known (below)
        q4[0] = (uint32_t)tv_[3 * t];
        q4[1] = (uint32_t)tv_[3 * t + 1];
        q4[2] = (uint32_t)tv_[3 * t + 2];
        q4[3] = (uint32_t)orig_[t];
      }
      max_tris = std::max(max_tris, cuts_[c + 1] - cuts_[c]);
    }
    for (int c = ncl_tri; c < ncl; ++c) {
      int32_t* row = cv + (size_t)c * kClusterVerts;
      int s = 0;
      while (s < kClusterVerts && li < loose.size()) row[s++] = loose[li++] | (int32_t)kVertOwner;
      nslots += s;
      tb[c] = nvalid_;
    }
    tb[ncl] = nvalid_;
    // ranks: owned vertices in cluster / slot order
    int32_t* rank_vert = reinterpret_cast<int32_t*>(blob.data() + h.off_rank_vert);
    int32_t* vert_rank = reinterpret_cast<int32_t*>(blob.data() + h.off_vert_rank);
    std::fill(rank_vert, rank_vert + nrank, -1);
    int next = 0;
    for (size_t i = 0; i < (size_t)ncl * kClusterVerts; ++i)
      if (cv[i] >= 0 && ((uint32_t)cv[i] & kVertOwner)) {
        const int v = (int)((uint32_t)cv[i] & kVertIdMask);
        vert_rank[v] = next;
        rank_vert[next++] = v | (int32_t)kVertOwner;
      }
    for (size_t i = 0; i < (size_t)nvalid_; ++i)
      for (int k = 0; k < 3; ++k) tq[4 * i + k] = (uint32_t)vert_rank[tq[4 * i + k]];
    int32_t* cluster_rank = reinterpret_cast<int32_t*>(blob.data() + mesh_off_cluster_rank(h));
    for (size_t i = 0; i < (size_t)ncl * kClusterVerts; ++i)
      cluster_rank[i] = cv[i] >= 0 ? vert_rank[(uint32_t)cv[i] & kVertIdMask] : -1;
    uint32_t* tr4 = reinterpret_cast<uint32_t*>(blob.data() + mesh_off_tri_rank4(h));       // (zero-initialised: dropped triangles)
    for (int i = 0; i < nvalid_; ++i) {
      uint32_t* e = tr4 + 4 * (size_t)orig_[i];
      for (int k = 0; k < 3; ++k) e[k] = (uint32_t)vert_rank[tv_[3 * i + k]];
      e[3] = 1u;
    }
    h.max_cluster_tris = max_tris;
    h.nvert_slots = nslots;
    uint32_t hash = 2166136261u;
    for (size_t i = sizeof(MeshTableHeader); i < blob.size(); ++i) hash = (hash ^ blob[i]) * 16777619u;
    h.hash = hash;
    std::memcpy(blob.data(), &h, sizeof(h));
    return blob;
  }
};

// Trivial table for a model without triangles: consecutive 128-vertex tiles, every vertex owned, no triangle entries.
// (fr_pack_basis uses the same vertex order when it is given no table.)

}  // namespace fr
#endif  // FR_MESH_TABLE_H_
