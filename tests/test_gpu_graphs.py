"""GPU: the traversability graph from the device kernels (gcp_build_graph) equals
what the unmodified reference's PathMaker.computeTraversableGraph builds (tests/golden/graph_ref.npz,
legacy_population.npz: made by make_golden_graph.py / make_golden_legacy.py)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _params(w):
    return w.map_params['scale'], 4.0, w.map_params['narrowest_hall']      # max_traverse_dist: pather_driver.py:186


@pytest.mark.parametrize("which", ["small", "big"])
def test_device_graph_equals_reference(which, world_small, world_big, golden_dir):
    from genetic_coverage_planning_b200 import graphs
    w = world_small if which == "small" else world_big
    z = np.load(os.path.join(golden_dir, "graph_ref.npz"))
    scale, max_dist, hall = _params(w)
    assert float(z[which + "_max_traverse_dist"]) == max_dist
    row_ptr, col, info = graphs.compute_traversable_graph_device(w.img, w.xy, scale, max_dist, hall,
                                                                 return_info=True)
    assert np.array_equal(row_ptr, z[which + "_row_ptr"])
    assert np.array_equal(col, z[which + "_col"])
    assert info["ambiguous_rows"] == 0
    print("%s: N=%d E=%d kernels %.3f ms" % (which, len(w.xy), len(col), info["kernel_ms"]))


def test_device_graph_on_the_cropped_world_and_dropin(golden_dir):
    """Waypoints 8 px from the border of a cropped map; and the drop-in fills `_graph` /
    `_crossover_cand` like PathMaker.computeTraversableGraph + findCrossoverCand."""
    from genetic_coverage_planning_b200 import graphs
    z = np.load(os.path.join(golden_dir, "legacy_population.npz"))

    class Bag(object):
        pass
    m, p = Bag(), Bag()
    m._img = z['img']
    p._mappy, p._XY, p._scale = m, z['xy'], float(z['scale'])
    p._max_dist, p._hall_width = float(z['max_traverse_dist']), float(z['narrowest_hall'])
    graphs.computeTraversableGraph(p)
    assert p._graph.dtype == np.float64 and np.array_equal(p._graph, z['graph'].astype(np.float64))
    want = np.where(z['graph'].sum(axis=1) > 2)
    assert np.array_equal(p._crossover_cand[0], want[0])


def test_ambiguous_rows_are_recomputed_on_the_host(world_small, golden_dir, monkeypatch):
    from genetic_coverage_planning_b200 import graphs, _lib
    real = _lib.GcpGraphParams

    class Wide(real):
        def __setattr__(self, k, v):
            if k == "eps_angle":
                v = 10.0
            real.__setattr__(self, k, v)
    monkeypatch.setattr(_lib, "GcpGraphParams", Wide)
    w = world_small
    z = np.load(os.path.join(golden_dir, "graph_ref.npz"))
    scale, max_dist, hall = _params(w)
    row_ptr, col, info = graphs.compute_traversable_graph_device(w.img, w.xy, scale, max_dist, hall,
                                                                 return_info=True)
    assert info["ambiguous_rows"] > 0.9 * len(w.xy)
    assert np.array_equal(row_ptr, z["small_row_ptr"]) and np.array_equal(col, z["small_col"])
