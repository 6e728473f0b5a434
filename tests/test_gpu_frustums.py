"""GPU: footprint polygons from the device kernel (gcp_compute_frustums) are
bit-identical to the tables the unmodified reference built (golden worlds) and to the windowed
numpy restatement on a synthetic world."""
import time

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _device(w):
    from genetic_coverage_planning_b200 import frustums
    return frustums.compute_frustums_device(w.img, w.xy, w.row_ptr, w.col, w.map_params, return_info=True)


@pytest.mark.parametrize("which", ["small", "big"])
def test_device_frustums_equal_reference_tables(which, world_small, world_big):
    """Every edge of the reference's two shipped maps (C1: 1,724 edges; C2: 1,618 edges, gray map):
    polygons, headings and lengths equal Mappy.computeFrustums' own output."""
    w = world_small if which == "small" else world_big
    poly, theta, dist, info = _device(w)
    assert poly.dtype == np.int32 and poly.shape == w.edge_poly.shape
    assert np.array_equal(poly, w.edge_poly)
    assert np.array_equal(theta, w.edge_theta)
    assert np.array_equal(dist, w.edge_dist)
    assert info["ambiguous_edges"] <= 2          # the exactness guard is an exception, not the path


def test_device_frustums_equal_host_restatement_on_synthetic_world():
    from genetic_coverage_planning_b200 import synth
    w = synth.make_world(1024, seed=5)           # host restatement builds edge_poly
    t0 = time.time()
    poly, theta, dist, info = _device(w)
    t_dev = time.time() - t0
    assert np.array_equal(poly, w.edge_poly)
    assert np.array_equal(theta, w.edge_theta) and np.array_equal(dist, w.edge_dist)
    assert info["ambiguous_edges"] <= max(2, w.E // 10000)
    print("device frustums: E=%d kernel %.3f ms, call %.2f s, ambiguous %d" %
          (w.E, info["kernel_ms"], t_dev, info["ambiguous_edges"]))


def test_ambiguous_edges_are_recomputed(world_big, monkeypatch):
    """With absurdly wide guard bands every edge is flagged and redone on the host path: same result."""
    from genetic_coverage_planning_b200 import frustums, _lib
    real = _lib.GcpFrustumParams

    class Wide(real):
        def __setattr__(self, k, v):
            if k == "eps_angle":
                v = 10.0
            real.__setattr__(self, k, v)
    monkeypatch.setattr(_lib, "GcpFrustumParams", Wide)
    w = world_big
    poly, theta, dist, info = frustums.compute_frustums_device(w.img, w.xy, w.row_ptr, w.col, w.map_params,
                                                               return_info=True)
    assert info["ambiguous_edges"] > 0.9 * len(w.col)
    assert np.array_equal(poly, w.edge_poly)


def test_dropin_fills_the_reference_attributes(world_small):
    """frustums.computeFrustums(mappy, graph) leaves the dense tables Mappy.computeFrustums would."""
    from genetic_coverage_planning_b200 import frustums
    w = world_small
    N = len(w.xy)

    class M(object):
        pass
    m = M()
    m._img, m._all_waypoints = w.img, w.xy
    mp = w.map_params
    m._scale, m._num_rays, m._min_view = mp['scale'], mp['num_rays'], mp['min_view']
    m._max_view, m._view_angle = mp['max_view'], mp['view_angle']
    graph = np.zeros((N, N))
    src = np.repeat(np.arange(N), np.diff(w.row_ptr))
    graph[src, w.col] = 1
    frustums.computeFrustums(m, graph)
    assert m._traverse_frustum.shape == (N, N, int(mp['num_rays']) + 2, 2)
    assert m._traverse_frustum.dtype == np.int32
    assert np.array_equal(m._traverse_frustum[src, w.col], w.edge_poly)
    assert np.array_equal(m._traverse_angles[src, w.col], w.edge_theta)
    assert np.array_equal(m._traverse_dists[src, w.col], w.edge_dist)
    assert m._traverse_frustum.sum(axis=(2, 3))[graph == 0].sum() == 0
