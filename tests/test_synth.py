"""CPU: the windowed table builders reproduce the reference's tables; synthetic worlds are
well formed."""
import numpy as np
import pytest

from genetic_coverage_planning_b200 import synth, worlds


def test_windowed_frustums_equal_reference_tables(world_big):
    """All 1,618 edges of the reference's big map: polygons, headings and lengths from
    the windowed CSR builder are identical to Mappy.computeFrustums' (golden world)."""
    w = world_big
    poly, theta, dist = synth.compute_frustums(w.img, w.xy, w.row_ptr, w.col, w.map_params)
    assert np.array_equal(poly, w.edge_poly)
    assert np.array_equal(theta, w.edge_theta)
    assert np.array_equal(dist, w.edge_dist)


def test_windowed_frustums_small_map(world_small):
    w = world_small
    poly, theta, dist = synth.compute_frustums(w.img, w.xy, w.row_ptr, w.col, w.map_params)
    assert np.array_equal(poly, w.edge_poly)
    assert np.array_equal(theta, w.edge_theta) and np.array_equal(dist, w.edge_dist)


@pytest.fixture(scope="module")
def synth_world():
    return synth.make_world(768, seed=3)


def test_synth_world_is_binary_and_connected(synth_world):
    import scipy.sparse as sp
    import scipy.sparse.csgraph as cg
    w = synth_world
    assert set(np.unique(w.img)) <= {0.0, 1.0}
    assert w.N > 300 and w.E > w.N
    g = sp.csr_matrix((np.ones(w.E), w.col, w.row_ptr), shape=(w.N, w.N))
    nc, lab = cg.connected_components(g, directed=False)
    assert np.bincount(lab).max() > 0.75 * w.N
    assert (w.img[w.xy[:, 0], w.xy[:, 1]] == 0).all()          # waypoints on free pixels
    # every edge within max_traverse_dist
    src = np.repeat(np.arange(w.N), np.diff(w.row_ptr))
    d = np.linalg.norm(w.xy[src] - w.xy[w.col], axis=1) * w.map_params['scale']
    assert (d < w.path_params['max_traverse_dist']).all()


def test_random_walks_follow_edges(synth_world):
    w = synth_world
    starts = synth.spread_starts(w.xy, 4, 768)
    dna = synth.random_walk_population(w.xy, w.row_ptr, w.col, starts, 16, 60, 80, seed=2)
    assert dna.shape == (16, 4, 80)
    assert (dna[:, :, 0] == np.array(starts)[None, :]).all()
    assert (dna[:, :, 61:] == -1).all() and (dna[:, :, :61] >= 0).all()
    edges = set(zip(np.repeat(np.arange(w.N), np.diff(w.row_ptr)).tolist(), w.col.tolist()))
    for row in dna.reshape(-1, 80):
        v = row[row >= 0]
        for a, b in zip(v[:-1], v[1:]):
            assert (int(a), int(b)) in edges or a == b
    assert np.array_equal(worlds.executed_steps(dna), np.full(16, 4 * 61))


def test_graph_row_restatement_equals_reference(world_small, golden_dir):
    """graphs._row_host (the exactness fallback of the device graph builder) reproduces rows of the
    graph the unmodified reference built (tests/golden/graph_ref.npz)."""
    import os
    from genetic_coverage_planning_b200 import graphs
    w = world_small
    z = np.load(os.path.join(golden_dir, "graph_ref.npz"))
    rp, col = z["small_row_ptr"], z["small_col"]
    nd = int(((w.map_params['narrowest_hall'] / 2) / 3) / w.map_params['scale'])
    obstacles = np.array(np.nonzero(graphs._dilated(w.img, nd))).T
    for i in range(0, len(w.xy), 7):
        got = graphs._row_host(i, w.xy, w.map_params['scale'], 4.0, obstacles)
        assert sorted(int(v) for v in got if v >= 0) == [int(v) for v in col[rp[i]:rp[i + 1]]]


@pytest.mark.parametrize("which", ["small", "big", "synth"])
def test_edge_headings_equal_reference_tables(which, world_small, world_big, golden_dir):
    """frustums.edge_headings (the host half of gcp_compute_frustums: heading, length, cosine and
    sine per edge from scalar numpy calls per distinct pixel offset) reproduces the reference's
    _traverse_angles / _traverse_dists bit for bit on all three reference-made worlds."""
    import os
    from world import World
    from genetic_coverage_planning_b200 import frustums
    w = {"small": world_small, "big": world_big}.get(which) or World.load_npz(os.path.join(golden_dir, "world_synth.npz"))
    theta, dist, co, si = frustums.edge_headings(w.xy, w.row_ptr, w.col)
    assert np.array_equal(theta, w.edge_theta) and np.array_equal(dist, w.edge_dist)
    assert np.array_equal(co, np.cos(w.edge_theta)) and np.array_equal(si, np.sin(w.edge_theta))
