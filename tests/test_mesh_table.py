"""CPU: the mesh table (3dfacerecon_b200/csrc/mesh_table.h, built through fr_mesh_table_create of the real library):
every valid triangle lands in exactly one cluster with local indices that point back at its own vertices, clusters
respect the 128-vertex / 256-triangle limits, every vertex has exactly one owner, blobs round-trip."""
import numpy as np
import pytest

from conftest import fr


def _check_table(tri, nver, table):
    p = table.parsed()
    tri = np.asarray(tri, np.float32)
    ntri = tri.shape[1]
    assert p["nver"] == nver and p["ntri"] == ntri
    cv, tb, te = p["cluster_vert"], p["tri_begin"], p["tri_entry"]
    ncl = p["nclusters"]
    assert cv.shape == (ncl, 128) and tb.shape == (ncl + 1,) and tb[0] == 0 and tb[-1] == p["ntri_slots"]
    assert (np.diff(tb) >= 0).all() and np.diff(tb).max(initial=0) <= 256 and np.diff(tb).max(initial=0) == p["max_cluster_tris"]
    used = cv >= 0
    ids = np.where(used, cv & 0x00FFFFFF, -1)
    owner = used & ((cv & 0x40000000) != 0)
    assert ids[used].max(initial=0) < nver
    assert int(used.sum()) == p["nvert_slots"]
    # exactly one owner per vertex, every vertex present
    counts = np.bincount(ids[owner], minlength=nver)
    assert (counts == 1).all()
    for c in range(ncl):                                   # no vertex twice in one cluster
        row = ids[c][used[c]]
        assert len(np.unique(row)) == len(row)
    # triangles: valid ones exactly once, local indices resolve to the original vertex ids
    valid = np.ones(ntri, bool)
    for k in range(3):
        valid &= (tri[k] > -1.0) & (tri[k] < nver)
    seen = np.zeros(ntri, int)
    cluster_of = np.repeat(np.arange(ncl), np.diff(tb))
    local, orig = te[:, 0], te[:, 1].astype(np.int64)
    np.add.at(seen, orig, 1)
    assert (seen[valid] == 1).all() and (seen[~valid] == 0).all()
    for k in range(3):
        lk = (local >> (8 * k)) & 0xFF
        got = ids[cluster_of, lk]
        assert (got == tri[k, orig].astype(np.int64)).all(), k
    assert (local >> 24 == 0).all()
    # the vertex-id flavour of the same list
    # ranks: a permutation of the vertices in cluster order; the rank flavour of the triangle list
    rv, vr = p["rank_vert"], p["vert_rank"]
    assert (rv[:nver] >= 0).all() and (rv[nver:] == -1).all() and ((rv[:nver] & 0x40000000) != 0).all()
    assert (np.sort(rv[:nver] & 0x00FFFFFF) == np.arange(nver)).all()
    assert (vr[rv[:nver] & 0x00FFFFFF] == np.arange(nver)).all()
    first_owner = np.full(nver, -1)
    for c in range(ncl):
        row = ids[c][owner[c]]
        first_owner[row] = c
    assert (np.diff(first_owner[rv[:nver] & 0x00FFFFFF]) >= 0).all()       # ranks follow the cluster order
    cr = p["cluster_rank"]                                                   # where the tile rasterizer finds a slot's record
    assert cr.shape == cv.shape and (cr[~used] == -1).all() and (cr[used] == vr[ids[used]]).all()
    t4 = p["tri_rank4"]                                                      # by original index: ranks of the three vertices, or zeros
    assert t4.shape == (ntri, 4) and (t4[:, 3] == valid.astype(np.uint32)).all() and not t4[~valid].any()
    for k in range(3):
        assert (t4[valid, k].astype(np.int64) == vr[tri[k, valid].astype(np.int64)]).all(), k
    tq = p["tri_vid"]
    assert (tq[:, 3].astype(np.int64) == orig).all()
    for k in range(3):
        assert (tq[:, k].astype(np.int64) == vr[tri[k, orig].astype(np.int64)]).all(), k
    return p


def test_mesh_table_grid_model():
    synth, mesh = fr("synth"), fr("mesh")
    m = synth.make_synthetic_model(grid=(61, 75), ndim_shape=2, ndim_exp=1, seed=1, jitter=0.2)
    nver = 61 * 75
    t = mesh.MeshTable(m["tri"], nver, m["mu"].reshape(3, nver))
    p = _check_table(m["tri"], nver, t)
    # the partition is good enough to be useful: < 1.35 slots per vertex, >= 150 triangles per cluster on average
    assert p["nvert_slots"] <= 1.35 * nver
    assert p["ntri_slots"] / p["nclusters"] >= 150
    # blob round trip
    t2 = mesh.MeshTable(blob=t.blob())
    assert t2.nclusters == t.nclusters and t2.blob().tobytes() == t.blob().tobytes()
    bad = t.blob()
    bad[200] ^= 0xFF
    with pytest.raises(ValueError, match="corrupt|inconsistent"):
        mesh.MeshTable(blob=bad)


def test_mesh_table_without_positions_and_permuted():
    synth, mesh = fr("synth"), fr("mesh")
    m = synth.make_synthetic_model(grid=(23, 31), ndim_shape=2, ndim_exp=1, seed=2)
    nver = 23 * 31
    rng = np.random.default_rng(0)
    perm = rng.permutation(nver)
    tri = perm[m["tri"].astype(np.int64)][:, rng.permutation(m["tri"].shape[1])].astype(np.float32)
    _check_table(tri, nver, mesh.MeshTable(tri, nver, None))
    pos = np.empty((3, nver), np.float32)
    pos[:, perm] = m["mu"].reshape(3, nver)
    p_with = _check_table(tri, nver, mesh.MeshTable(tri, nver, pos))
    p_grid = _check_table(m["tri"], nver, mesh.MeshTable(m["tri"], nver, m["mu"].reshape(3, nver)))
    assert p_with["nclusters"] == p_grid["nclusters"]          # locality comes from the table, not from the vertex order
    # interleaved positions give the same partition as planar ones
    p_int = mesh.MeshTable(tri, nver, np.ascontiguousarray(pos.T), interleaved=True).parsed()
    assert p_int["nclusters"] == p_with["nclusters"]


def test_mesh_table_soups_and_edge_cases():
    mesh = fr("mesh")
    rng = np.random.default_rng(5)
    # random soup with invalid / degenerate / duplicate triangles and unreferenced vertices
    nver = 700
    tri = rng.integers(0, 500, (3, 3000)).astype(np.float32)
    tri[:, 10] = [-1.0, 5.0, 6.0]
    tri[:, 11] = [3.0, 700.0, 6.0]
    tri[:, 12] = [np.nan, 1.0, 2.0]
    tri[:, 13] = [7.0, 7.0, 7.0]
    tri[:, 14] = tri[:, 15]
    tri[:, 16] = [2.9, 3.2, 4.99]                             # fractional values truncate like (int)tri(k,i)
    _check_table(tri, nver, mesh.MeshTable(tri, nver, rng.normal(size=(3, nver)).astype(np.float32)))
    _check_table(tri, nver, mesh.MeshTable(tri, nver, None))
    # no triangles at all: only loose vertices
    p = _check_table(np.zeros((3, 0), np.float32), 300, mesh.MeshTable(np.zeros((3, 0), np.float32), 300, None))
    assert p["nclusters"] == 3 and p["ntri_slots"] == 0
    # one triangle, three vertices; non-finite positions are tolerated
    pos = np.full((3, 3), np.nan, np.float32)
    p = _check_table(np.array([[0], [1], [2]], np.float32), 3, mesh.MeshTable(np.array([[0], [1], [2]], np.float32), 3, pos))
    assert p["nclusters"] == 1
    # a fan around one vertex with more than 128 distinct neighbours must be split
    k = 400
    fan = np.stack([np.zeros(k), 1 + np.arange(k), 1 + (np.arange(k) + 1) % k]).astype(np.float32)
    p = _check_table(fan, k + 1, mesh.MeshTable(fan, k + 1, None))
    assert p["nclusters"] >= 4
    with pytest.raises(ValueError):
        mesh.MeshTable(np.zeros((2, 5), np.float32), 10)


@pytest.mark.parametrize("flavour", ["grid", "permuted"])
def test_mesh_table_bfm_size(flavour):
    """The BASELINE mesh (53 215 vertices, 105 840 triangles): validity, and the partition quality the fused kernel's byte
    count depends on (every cluster is one 128-row pass over the basis)."""
    synth, mesh = fr("synth"), fr("mesh")
    m = synth.make_synthetic_model(ndim_shape=1, ndim_exp=1, seed=0, jitter=0.2, permute=(flavour == "permuted"))
    nver = m["mu"].size // 3
    t = mesh.MeshTable(m["tri"], nver, m["mu"].reshape(3, nver))
    p = _check_table(m["tri"], nver, t)
    assert p["nclusters"] <= 540, p["nclusters"]              # 416 tiles without duplication; ideal for this grid ~ 505
    assert p["nvert_slots"] <= 1.25 * nver
