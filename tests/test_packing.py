"""Host packing: sub-tile union == reference painted set; window raster == full-map raster."""
import os

import cv2
import numpy as np
import pytest

import fitness_oracle as fo
from genetic_coverage_planning_b200 import packing
from helpers import edge_ids_for_dna, painted_blocks


@pytest.fixture(scope="module")
def pw_big(world_big):
    w = world_big
    return packing.pack_world(w.img, w.xy, w.row_ptr, w.col, w.edge_poly, w.edge_theta,
                              w.edge_dist, w.map_params)


def _blocks_to_map(pw, acc):
    full = np.zeros((pw.nby * 64, pw.nbx * 64), dtype=bool)
    for b, rows in acc.items():
        by, bx = divmod(b, pw.nbx)
        bits = np.unpackbits(rows.view(np.uint8).reshape(64, 8), axis=1, bitorder='little')
        full[by * 64:by * 64 + 64, bx * 64:bx * 64 + 64] = bits.astype(bool)
    return full[:pw.H, :pw.W]


def test_window_raster_equals_full_map(world_big, pw_big):
    """Every edge (and every null edge): the union of its packed sub-tiles equals
    cv2.fillPoly on the full map, which is what mappy.py:355 executes."""
    w, pw = world_big, pw_big
    src = np.repeat(np.arange(w.N), np.diff(w.row_ptr))
    rng = np.random.RandomState(0)
    edges = list(rng.choice(w.E, 300, replace=False)) + list(w.E + rng.choice(w.N, 60, replace=False))
    zero = np.zeros_like(w.edge_poly[0])
    for e in edges:
        node = src[e] if e < w.E else e - w.E
        poly = w.edge_poly[e] if e < w.E else zero
        canvas = np.zeros(w.shape, dtype=np.uint8)
        cv2.fillPoly(canvas, [poly], 1, offset=(int(w.xy[node, 1]), int(w.xy[node, 0])))
        got = _blocks_to_map(pw, painted_blocks(pw, np.array([e])))
        assert np.array_equal(got, canvas.astype(bool)), e
        assert pw.edge_pixels[e] == canvas.sum()


def test_union_equals_reference_painted_set(world_big, pw_big, golden_dir):
    z = np.load(os.path.join(golden_dir, "eval_big.npz"))
    ze = np.load(os.path.join(golden_dir, "edge_big.npz"))
    mask = world_big.buffer_mask
    for zz, n in ((z, 6), (ze, len(ze['dna']))):
        for k in range(n):
            dna = zz['dna'][k].astype(int)
            eid = edge_ids_for_dna(pw_big, dna)
            painted = _blocks_to_map(pw_big, painted_blocks(pw_big, eid))
            assert int(painted.sum()) == zz['n_painted'][k]
            assert int((painted & mask).sum()) == zz['n_painted_mask'][k]
            # per-agent costs straight from the packed edge records, sequential fp64 adds
            for a in range(dna.shape[0]):
                d = t = 0.0
                for e in eid[a][eid[a] >= 0]:
                    d += pw_big.edge_cost[e, 0]
                    t += pw_big.edge_cost[e, 1]
                assert d == zz['dist'][k][a] and t == zz['turn'][k][a]


def test_gray_weights_reproduce_reference_sums(world_big, pw_big, golden_dir):
    """Integer restatement of the two float sums of mappy.py:358,361."""
    pw = pw_big
    z = np.load(os.path.join(golden_dir, "eval_big.npz"))
    g = np.rint(world_big.img * 255).astype(np.int64)
    mask = world_big.buffer_mask
    assert pw.mask_size == mask.sum() and pw.gsum_mask == g[mask].sum()
    for k in range(4):
        dna = z['dna'][k].astype(int)
        painted = _blocks_to_map(pw, painted_blocks(pw, edge_ids_for_dna(pw, dna)))
        w_mask = int((255 - g)[painted & mask].sum())
        w_all = int((255 - g)[painted].sum())
        assert abs((pw.gsum_mask + w_mask) / 255.0 - z['sum_draw_mask'][k]) <= 1e-9 * z['sum_draw_mask'][k]
        assert abs((pw.num_occluded + w_all / 255.0) - z['sum_draw'][k]) <= 1e-9 * z['sum_draw'][k]
        wall = (pw.gsum_mask + w_mask) / 255.0 / pw.mask_size
        assert abs(-wall - z['neg_cov'][k]) <= 1e-12


def test_layout_invariants(pw_big):
    pw = pw_big
    assert pw.sub_ptr[-1] == len(pw.sub_block) == len(pw.sub_payload)
    assert len(pw.tile_data) == pw.n_quarters * 16
    assert (pw.sub_block < pw.n_blocks).all()
    nq = np.array([bin(int(p) & 15).count("1") for p in pw.sub_payload])
    assert nq.min() >= 1 and nq.sum() == pw.n_quarters
    # mask0 is a subset of free0; out-of-map bits are zero
    assert not (pw.blk_mask0 & ~pw.blk_free0).any()
    assert pw.gray_ptr[-1] == len(pw.gray_ent)


def test_pack_from_live_reference_objects(world_small):
    """The drop-in path: packing.pack_from_reference on the reference's own Mappy / PathMaker
    objects (this container only) gives exactly the tables packed from the committed fixture."""
    import ref_shim
    if not ref_shim.available():
        pytest.skip("needs /root/reference")
    from genetic_coverage_planning_b200 import packing
    m, p, _ = ref_shim.build_world("small", 6)
    a = packing.pack_from_reference(m, p)
    w = world_small
    b = packing.pack_world(w.img, w.xy, w.row_ptr, w.col, w.edge_poly, w.edge_theta, w.edge_dist,
                           w.map_params)
    for name in ("row_ptr", "col", "xy", "xover_cand", "edge_cost", "edge_theta", "sub_ptr", "sub_block",
                 "sub_payload", "tile_data", "m_sub_ptr", "m_sub_block", "m_sub_payload", "m_tile_data",
                 "blk_mask0", "blk_free0", "gray_ptr", "gray_ent"):
        assert np.array_equal(getattr(a, name), getattr(b, name)), name
    assert (a.mask_size, a.gsum_mask, a.map_size, a.num_occluded) == \
           (b.mask_size, b.gsum_mask, b.map_size, b.num_occluded)


def test_legacy_pickle_opens_without_the_reference(golden_dir):
    """A population pickled by the unmodified reference (tests/golden/make_golden_legacy.py) opens
    through legacy.load_population with no reference module importable, and carries the state the
    facade needs."""
    import os
    from genetic_coverage_planning_b200 import legacy, packing
    old = legacy.load_population(os.path.join(golden_dir, "legacy_population.pkl.gz"))
    for obj, name in ((old, "geneticalgorithm.GeneticalGorithm"), (old._gen_parent[0], "geneticalgorithm.Organism"),
                      (old._gen_parent[0]._mappy, "mappy.Mappy"), (old._gen_parent[0]._pather, "pathmaker.PathMaker")):
        assert type(obj).__module__ == "genetic_coverage_planning_b200.legacy"      # attribute bags, not reference classes
        assert type(obj)._legacy_class == name
    z = np.load(os.path.join(golden_dir, "legacy_population.npz"))
    assert old._gen_num == int(z['gen_num']) and old._G_sz == len(z['dna'])
    assert np.array_equal(np.array([o._dna for o in old._gen_parent]), z['dna'])
    assert np.array_equal(np.array([o._obj_val for o in old._gen_parent], dtype=float), z['obj'])
    first = old._gen_parent[0]
    assert np.array_equal(first._pather._graph, z['graph']) and np.array_equal(first._pather._XY, z['xy'])
    pw = packing.pack_from_reference(first._mappy, first._pather)      # tables from the pickled state
    assert pw.N == len(z['xy']) and pw.E == int(z['graph'].sum())


def test_footprints_clipped_at_the_map_border_like_cv2(golden_dir):
    """The cropped legacy world has waypoints 8 px from the map border, so most footprints are
    clipped: the packed tiles must equal what cv2.fillPoly leaves on the map itself (the oracle
    paints the full map like mappy.py:353-355).  A polygon drawn unclipped and cropped afterwards
    differs along the border (cv2 restarts clipped lines) - this pins the difference."""
    import os
    import fitness_oracle as fo
    from world import World
    import helpers
    from genetic_coverage_planning_b200 import legacy, packing
    old = legacy.load_population(os.path.join(golden_dir, "legacy_population.pkl.gz"))
    first = old._gen_parent[0]
    w = World.from_reference(first._mappy, first._pather, first._org_params)
    pw = packing.pack_from_reference(first._mappy, first._pather)
    H, W = w.img.shape
    for o in old._gen_parent:
        dna = np.asarray(o._dna)
        acc = helpers.painted_blocks(pw, helpers.edge_ids_for_dna(pw, dna))
        mine = np.zeros((pw.nby * 64, pw.nbx * 64), dtype=bool)
        for b, rows in acc.items():
            by, bx = divmod(b, pw.nbx)
            mine[by * 64:by * 64 + 64, bx * 64:bx * 64 + 64] = \
                ((rows[:, None] >> np.arange(64, dtype='<u8')[None, :]) & 1).astype(bool)
        assert not mine[H:].any() and not mine[:, W:].any()
        assert np.array_equal(mine[:H, :W], fo._painted_map(w, dna))
        assert tuple(fo.calc_obj(w, dna)) == tuple(o._obj_val)       # oracle == the reference's stored objectives


def test_parallel_packing_is_identical():
    """pack_world(workers=N) rasterises edge ranges in forked processes and stitches the stores:
    every array must equal the single-process result."""
    from genetic_coverage_planning_b200 import packing, synth
    w = synth.make_world(640, seed=4)
    a = packing.pack_world(w.img, w.xy, w.row_ptr, w.col, w.edge_poly, w.edge_theta, w.edge_dist,
                           w.map_params, workers=1)
    b = packing.pack_world(w.img, w.xy, w.row_ptr, w.col, w.edge_poly, w.edge_theta, w.edge_dist,
                           w.map_params, workers=3)
    assert set(a.__dict__) == set(b.__dict__)
    for k, v in a.__dict__.items():
        u = b.__dict__[k]
        if isinstance(v, np.ndarray):
            assert v.dtype == u.dtype and np.array_equal(v, u), k
        else:
            assert v == u, k
