// Host build of the mesh-table builder (3dfacerecon_b200/csrc/mesh_table.h) for CPU experiments and tests
// (test infrastructure; the product reaches the same builder through fr_mesh_table_create in the CUDA library).
#include <cstring>
#include "../../3dfacerecon_b200/csrc/mesh_table.h"

extern "C" long long mesh_emul_build(const float* tri, int ntri, int nver, const float* pos, int interleaved,
                                     unsigned char* out, long long cap) {
  fr::MeshTableBuilder b(tri, ntri, nver, pos, interleaved != 0);
  std::vector<unsigned char> blob = b.build();
  if ((long long)blob.size() > cap) return -(long long)blob.size();
  std::memcpy(out, blob.data(), blob.size());
  return (long long)blob.size();
}
