"""Plain-numpy container for everything the fitness path reads ("the world").

TEST INFRASTRUCTURE (oracle side).  The oracle works on this container so that it
can run where /root/reference does not exist (the GPU box): a World is built
either from live reference objects (`from_reference`, this container only), from
a committed golden fixture (`load_npz`), or by the synthetic generator of the
product package.

Reference tables restated (all citations into /root/reference/src):
  mappy.py:274-328   _traverse_frustum[N,N,17,2] / _traverse_angles / _traverse_dists
                     -> stored per directed edge in CSR order (edge_poly/theta/dist);
                     table[i, j] for a non-edge is all zeros (np.zeros default, :275-277)
  pathmaker.py:166-173  _graph (dense 0/1 float64) -> row_ptr/col, _XY -> xy
  mappy.py:36-39     _safety_img = 0.25*img + 0.25*dilate(img, 3x3, int((hall/2)/scale))
"""
import numpy as np


class World(object):
    def __init__(self, img, xy, row_ptr, col, edge_poly, edge_theta, edge_dist,
                 map_params, org_params=None, name="world"):
        self.name = name
        self.img = np.ascontiguousarray(img, dtype=np.float64)          # g/255
        self.xy = np.ascontiguousarray(xy, dtype=np.int64)              # (row, col)
        self.row_ptr = np.ascontiguousarray(row_ptr, dtype=np.int64)
        self.col = np.ascontiguousarray(col, dtype=np.int64)
        self.edge_poly = np.ascontiguousarray(edge_poly, dtype=np.int32)  # [E,R+2,2] (x,y)
        self.edge_theta = np.ascontiguousarray(edge_theta, dtype=np.float64)
        self.edge_dist = np.ascontiguousarray(edge_dist, dtype=np.float64)
        self.map_params = dict(map_params)
        self.org_params = dict(org_params) if org_params is not None else None
        self.N = int(self.xy.shape[0])
        self.E = int(self.col.shape[0])
        self.shape = self.img.shape
        self._safety = None
        self._edge_lut = None
        # mappy.py:25
        self.num_occluded = np.sum(self.img)

    # -- derived quantities -------------------------------------------------
    @property
    def safety_img(self):
        """mappy.py:36-39 (same cv2.dilate call on the same float64 image)."""
        if self._safety is None:
            import cv2
            hall = self.map_params['narrowest_hall']
            scale = self.map_params['scale']
            num_dilations = int((hall / 2) / scale)
            kernel = np.ones((3, 3), np.uint8)
            dil = cv2.dilate(self.img, kernel, iterations=num_dilations)
            self._safety = 0.25 * self.img + 0.25 * dil
        return self._safety

    @property
    def buffer_mask(self):
        """mappy.py:337"""
        s = self.safety_img
        return np.logical_and(s < 0.3, s > 0)

    def edge_id(self, frm, to):
        """Index of directed edge frm->to in CSR order, or -1 (non-edge).
        `to` follows numpy's negative-index rule (F3: -1 -> N-1)."""
        if to < 0:
            to += self.N
        if self._edge_lut is None:
            src = np.repeat(np.arange(self.N, dtype=np.int64), np.diff(self.row_ptr))
            self._edge_lut = dict(zip((src * self.N + self.col).tolist(), range(self.E)))
        return self._edge_lut.get(int(frm) * self.N + int(to), -1)

    def crossover_cand(self):
        """pathmaker.py:157-161: nodes with out-degree > 2."""
        return np.where(np.diff(self.row_ptr) > 2)[0]

    # -- constructors -------------------------------------------------------
    @staticmethod
    def from_reference(mappy, pather, org_params=None, name="ref"):
        g = np.asarray(pather._graph)
        src, dst = np.nonzero(g == 1)                     # row-major => CSR order
        N = g.shape[0]
        row_ptr = np.zeros(N + 1, dtype=np.int64)
        np.add.at(row_ptr, src + 1, 1)
        row_ptr = np.cumsum(row_ptr)
        mp = {'scale': mappy._scale, 'narrowest_hall': mappy._hall_width,
              'num_rays': mappy._num_rays, 'min_view': mappy._min_view,
              'max_view': mappy._max_view, 'view_angle': mappy._view_angle,
              'rho': mappy._rho, 'solo_sep_thresh': mappy._solo_sep_thresh,
              'coverage_blend': mappy._coverage_blend}
        return World(mappy._img, pather._XY, row_ptr, dst,
                     mappy._traverse_frustum[src, dst],
                     mappy._traverse_angles[src, dst],
                     mappy._traverse_dists[src, dst], mp, org_params, name)

    def save_npz(self, path):
        img_u8 = np.rint(self.img * 255).astype(np.uint8)
        assert np.array_equal(img_u8 / 255, self.img), "map is not an 8-bit image / 255"
        mp_keys = sorted(self.map_params)
        op = self.org_params or {}
        op_keys = sorted(k for k in op if k != 'start_idx')
        np.savez_compressed(
            path, img_u8=img_u8, xy=self.xy.astype(np.int32),
            row_ptr=self.row_ptr.astype(np.int32), col=self.col.astype(np.int32),
            edge_poly=self.edge_poly.astype(np.int8 if np.abs(self.edge_poly).max() < 128 else np.int32),
            edge_theta=self.edge_theta, edge_dist=self.edge_dist,
            mp_keys=np.array(mp_keys), mp_vals=np.array([float(self.map_params[k]) for k in mp_keys]),
            op_keys=np.array(op_keys), op_vals=np.array([float(op[k]) for k in op_keys]),
            start_idx=np.array(op.get('start_idx', []), dtype=np.int64), name=np.array(self.name))

    @staticmethod
    def load_npz(path):
        z = np.load(path)
        mp = {str(k): float(v) for k, v in zip(z['mp_keys'], z['mp_vals'])}
        for k in ('num_rays', 'solo_sep_thresh'):
            if k in mp and float(mp[k]).is_integer():
                mp[k] = int(mp[k])
        op = {str(k): float(v) for k, v in zip(z['op_keys'], z['op_vals'])}
        for k, v in list(op.items()):
            if float(v).is_integer() and k not in ('flight_time_scale', 'crossover_prob',
                                                   'mutation_prob', 'muterpolate_prob',
                                                   'muterpolation_sub_prob'):
                op[k] = int(v)
        op['start_idx'] = [int(s) for s in z['start_idx']]
        return World(z['img_u8'] / 255, z['xy'], z['row_ptr'], z['col'], z['edge_poly'],
                     z['edge_theta'], z['edge_dist'], mp, op if len(op) > 1 else None,
                     str(z['name']))
