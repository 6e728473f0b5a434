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
