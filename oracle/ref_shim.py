"""Import the UNMODIFIED reference (/root/reference/src) in this container.

TEST INFRASTRUCTURE ONLY.  Used by tests/golden/make_golden.py (fixture
generation) and by tests that are skipped when /root/reference is absent (it
does not exist on the GPU box).  Nothing in the product package imports this.

The reference modules import IPython, matplotlib and `cv2.cv2`, none of which
exist here; they are only used for GUI / debugging, so they are
replaced by attribute-swallowing stubs.  No reference file is modified or copied.
"""
import os
import sys
import types

REF_ROOT = os.environ.get("GCP_REFERENCE_ROOT", "/root/reference")
REF_SRC = os.path.join(REF_ROOT, "src")
REF_DATA = os.path.join(REF_ROOT, "data")


def available():
    return os.path.isfile(os.path.join(REF_SRC, "geneticalgorithm.py"))


class _Dummy(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        d = _Dummy(self.__name__ + "." + name)
        setattr(self, name, d)
        return d

    def __call__(self, *a, **k):
        return _Dummy("call")


def _install_stubs():
    import cv2  # real OpenCV; the reference wants the legacy alias cv2.cv2

    sys.modules.setdefault("cv2.cv2", cv2)
    for name in ("IPython", "IPython.core", "IPython.core.debugger",
                 "matplotlib", "matplotlib.pyplot", "matplotlib.gridspec",
                 "matplotlib.animation"):
        if name not in sys.modules:
            sys.modules[name] = _Dummy(name)
    dbg = sys.modules["IPython.core.debugger"]

    def set_trace(*a, **k):
        raise RuntimeError("reference called set_trace()")

    dbg.set_trace = set_trace


_mods = None


def load():
    """Return (geneticalgorithm, mappy, pathmaker, gori_tools) reference modules."""
    global _mods
    if _mods is not None:
        return _mods
    if not available():
        raise ImportError("reference not present at %s" % REF_SRC)
    os.environ.setdefault("TQDM_DISABLE", "1")
    _install_stubs()
    if REF_SRC not in sys.path:
        sys.path.insert(0, REF_SRC)
    # import order matters: the modules reload() their siblings
    import geneticalgorithm  # noqa
    import mappy  # noqa
    import pathmaker  # noqa
    import gori_tools  # noqa
    _mods = (sys.modules["geneticalgorithm"], sys.modules["mappy"],
             sys.modules["pathmaker"], sys.modules["gori_tools"])
    return _mods


def driver_params(num_agents=6):
    """The four parameter dicts of pather_driver.py:66-200 (values restated, the
    driver itself cannot be imported: it runs the whole experiment)."""
    import numpy as np
    solo_sep_thresh = [30, 30, 20, 10, 10, 10]
    starting_path_len = [200, 150, 100, 100, 75, 150]
    start_idx = [[70], [70, 70], [93, 93, 93], [99, 0, 15, 184],
                 [99, 99, 0, 15, 184], [0, 68, 536, 450, 78, 97]]
    max_dna_len = [250, 200, 150, 125, 125, 200]
    min_dna_len = [70, 60, 50, 30, 30, 30]
    num_muterpolations = [20, 20, 15, 15, 10, 20]
    min_solo_lcs = [5, 5, 5, 2, 1, 2]
    min_comb_lcs = [0, 10, 5, 5, 5, 5]
    k = num_agents - 1
    map_params = {'scale': 0.15, 'narrowest_hall': 2.5, 'num_rays': 15,
                  'min_view': 0.2, 'max_view': 7,
                  'view_angle': 69.4 * np.pi / 180, 'rho': 4.0,
                  'solo_sep_thresh': solo_sep_thresh[k], 'bw_thresh': 90,
                  'scale_px2m': 1 / 0.44 * 0.0254 / 0.709,
                  'coverage_blend': 1.0}
    path_params = {'path_memory': 10, 'max_traverse_dist': 4.0,
                   'waypoint_dist_factor': 0.9, 'wall_waypoint_factor': 2.5,
                   'log_scale_weight': 0.7}
    org_params = {'start_idx': start_idx[k], 'max_dna_len': max_dna_len[k],
                  'min_dna_len': min_dna_len[k], 'crossover_prob': 0.7,
                  'crossover_time_thresh': 70, 'mutation_prob': 0.3,
                  'muterpolate_prob': 0.2,
                  'num_muterpolations': num_muterpolations[k],
                  'muterpolation_srch_dist': 5, 'muterpolation_sub_prob': 0.8,
                  'min_solo_lcs': min_solo_lcs[k], 'min_comb_lcs': min_comb_lcs[k],
                  'flight_time_scale': 0.0001}
    gen_params = {'gen_size': 100, 'starting_path_len': starting_path_len[k],
                  'num_agents': num_agents, 'gamma': 0.5,
                  'coverage_constr_0': 0.3, 'coverage_constr_f': 0.95,
                  'coverage_aging': 80, 'org_params': org_params}
    return map_params, path_params, org_params, gen_params


def build_world(config="big", num_agents=6, blend=None):
    """Build reference Mappy/PathMaker for config 'big' (C2) or 'small' (C1, F14)."""
    import numpy as np
    import cv2
    ga, mappy_m, pm_m, got = load()
    map_params, path_params, org_params, gen_params = driver_params(num_agents)
    if blend is not None:
        map_params['coverage_blend'] = blend
    if config == "big":
        img_f, g_f, w_f = "big_map_scaled.png", "wilk_3_graph_big.npy", "wilk_3_wpts_big.npy"
    elif config == "small":
        img_f, g_f, w_f = "map_scaled.png", "wilk_3_graph.npy", "wilk_3_wpts.npy"
        if num_agents == 6:
            org_params['start_idx'] = [207, 207, 1, 10, 305, 207]
    else:
        raise ValueError(config)
    img = cv2.imread(os.path.join(REF_DATA, img_f), cv2.IMREAD_GRAYSCALE) / 255
    m = mappy_m.Mappy(img, map_params)
    p = pm_m.PathMaker(m, path_params)
    p.loadTraversableGraph(os.path.join(REF_DATA, g_f))
    p.loadWptsXY(os.path.join(REF_DATA, w_f))
    p.findCrossoverCand()  # F11: the shipped driver forgets this
    m.setWaypoints(p.waypoint_locs)
    m.computeFrustums(p._graph)
    return m, p, (map_params, path_params, org_params, gen_params)


class stable_argsort(object):
    """SURVEY F9/F15 canonicalisation: np.argsort -> kind='stable' while active."""

    def __enter__(self):
        import numpy as np
        self._np, self._orig = np, np.argsort

        def _stable(a, *args, **kw):
            kw['kind'] = 'stable'
            return self._orig(a, *args, **kw)
        np.argsort = _stable

    def __exit__(self, *exc):
        self._np.argsort = self._orig


CROP = (40, 200, 20, 200)      # r0, r1, c0, c1 of data/map_scaled.png (tests/golden/make_golden_legacy.py)


def build_cropped_world(num_agents=3):
    """Reference Mappy / PathMaker on a crop of the small map with the shipped waypoints that fall
    inside it (8 px margin); graph from the reference's own computeTraversableGraph."""
    import numpy as np
    import cv2
    ga, mappy_m, pm_m, got = load()
    mp, pp, op, gp = driver_params(num_agents)
    img = cv2.imread(os.path.join(REF_DATA, "map_scaled.png"), cv2.IMREAD_GRAYSCALE) / 255
    r0, r1, c0, c1 = CROP
    img = np.ascontiguousarray(img[r0:r1, c0:c1])
    xy = np.load(os.path.join(REF_DATA, "wilk_3_wpts.npy"))
    keep = (xy[:, 0] >= r0 + 8) & (xy[:, 0] < r1 - 8) & (xy[:, 1] >= c0 + 8) & (xy[:, 1] < c1 - 8)
    xy = xy[keep] - np.array([r0, c0])
    m = mappy_m.Mappy(img, mp)
    p = pm_m.PathMaker(m, pp)
    p._XY = xy.astype(int)
    p.computeTraversableGraph()
    m.setWaypoints(p.waypoint_locs)
    m.computeFrustums(p._graph)
    return m, p, (mp, pp, op, gp)
