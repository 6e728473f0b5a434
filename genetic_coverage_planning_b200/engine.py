"""FitnessEngine: the device side of `Organism.calcObj` / `GeneticalGorithm.rackNstack`
for whole populations.  PyTorch is used only to own device buffers and streams; all
compute is in libgcp.so (hand-written sm_100a kernels) behind the C ABI of include/gcp.h.

Reference call sites replaced (file:line into /root/reference/src):
  geneticalgorithm.py:177-201  Organism.calcObj           -> FitnessEngine.evaluate
  geneticalgorithm.py:86-123   GeneticalGorithm.rackNstack -> FitnessEngine.maximin
  geneticalgorithm.py:78-84    GeneticalGorithm.arena      -> FitnessEngine.arena_order
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import GcpEvalOut, GcpMapConsts, GcpParams, GcpShards, GcpVaryParams


class GcpError(RuntimeError):
    pass


def _ptr(a):
    """Host pointer of a C-contiguous numpy array (kept alive by the caller)."""
    return a.ctypes.data_as(C.c_void_p)


class EvalResult(object):
    """Device tensors produced by one evaluate() call."""
    __slots__ = ("obj", "neg_cov", "dist", "turn", "cov", "lc_mat", "flags")

    def feasible(self):
        return self.flags[:, 0]

    def traversable(self):
        return self.flags[:, 1]

    def n_components(self):
        return self.flags[:, 2]

    def error(self):
        return self.flags[:, 3]


class _RawCuda(object):
    """Lets torch view device memory that libgcp.so allocated (gcp_shard_alloc)."""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1",
                                         "data": (int(ptr), False), "version": 2}


class ShardBuffer(object):
    """Device memory owned by the engine's context that the other ranks of the box can map
    (CUDA IPC).  `handle`: the 64-byte IPC handle to send to the peers; `tensor`: uint8 view."""

    def __init__(self, engine, nbytes):
        ptr = C.c_void_p()
        h = (C.c_uint8 * 64)()
        engine._check(engine.lib.gcp_shard_alloc(engine.ctx, int(nbytes), C.byref(ptr), h))
        self.ptr, self.nbytes, self.handle = int(ptr.value), int(nbytes), bytes(h)
        self.engine = engine
        self.tensor = torch.as_tensor(_RawCuda(self.ptr, max(int(nbytes), 1)), device=engine.device)

    def view(self, dtype, shape):
        return self.tensor.view(dtype).view(*shape)


class FitnessEngine(object):
    def __init__(self, packed, org_params, num_agents=None, device=0):
        """packed: packing.PackedWorld; org_params: the reference's org_params dict
        (pather_driver.py:179-191).  Raises if CUDA or libgcp.so is unavailable."""
        if not torch.cuda.is_available():
            raise GcpError("FitnessEngine needs a CUDA device: there is no CPU fallback")
        self.lib = _lib.load()
        self.device = torch.device("cuda", device)
        self.pw = packed
        ctx = C.c_void_p()
        rc = self.lib.gcp_create(C.byref(ctx), int(device))
        if rc != 0:
            raise GcpError("gcp_create: %s" % self.lib.gcp_last_error(None).decode())
        self.ctx = ctx
        pw = packed
        self._keep = []

        def arr(a, dt):
            a = np.ascontiguousarray(a, dtype=dt)
            self._keep.append(a)
            return a
        self._check(self.lib.gcp_upload_graph(
            ctx, pw.N, pw.E, _ptr(arr(pw.row_ptr, np.uint32)), _ptr(arr(pw.col, np.uint32)),
            _ptr(arr(pw.xy, np.int32)), _ptr(arr(pw.xover_cand, np.uint8))))
        self._check(self.lib.gcp_upload_edges(
            ctx, _ptr(arr(pw.edge_cost, np.float64)), _ptr(arr(pw.edge_theta, np.float64)),
            _ptr(arr(pw.sub_ptr, np.uint32)), _ptr(arr(pw.sub_block, np.uint32)),
            _ptr(arr(pw.sub_payload, np.uint32)), len(pw.sub_block),
            _ptr(arr(pw.tile_data, np.uint64)), pw.n_quarters))
        self.has_masked = hasattr(pw, "m_sub_ptr")
        if self.has_masked:
            self._check(self.lib.gcp_upload_masked_footprints(
                ctx, _ptr(arr(pw.m_sub_ptr, np.uint32)), _ptr(arr(pw.m_sub_block, np.uint32)),
                _ptr(arr(pw.m_sub_payload, np.uint32)), len(pw.m_sub_block),
                _ptr(arr(pw.m_tile_data, np.uint64)), pw.m_n_quarters))
        mc = GcpMapConsts(pw.H, pw.W, pw.nbx, pw.nby, pw.mask_size, pw.gsum_mask, pw.map_size,
                          pw.num_occluded)
        self._check(self.lib.gcp_upload_map(
            ctx, C.byref(mc), _ptr(arr(pw.blk_mask0, np.uint64)),
            _ptr(arr(pw.blk_free0, np.uint64)), _ptr(arr(pw.gray_ptr, np.uint32)),
            _ptr(arr(pw.gray_ent, np.uint32)), len(pw.gray_ent)))
        self._keep = []
        self.A = int(num_agents if num_agents is not None else len(org_params['start_idx']))
        self.L = int(org_params['max_dna_len'])
        self.set_params(org_params)
        self._scratch = None
        self._rank_scratch = None

    # ------------------------------------------------------------------
    def set_params(self, org_params, solo_sep_thresh=None, blend=None):
        pw = self.pw
        blend = pw.blend if blend is None else float(blend)
        p = GcpParams(self.A, self.L,
                      int(pw.solo_sep_thresh if solo_sep_thresh is None else solo_sep_thresh),
                      int(org_params['min_solo_lcs']), int(org_params['min_comb_lcs']), 0,
                      float(org_params['flight_time_scale']), blend, 1 - blend)
        self._check(self.lib.gcp_set_params(self.ctx, C.byref(p)))
        self.org_params = dict(org_params)

    def _check(self, rc):
        if rc != 0:
            raise GcpError("libgcp error %d: %s" % (rc, self.lib.gcp_last_error(self.ctx).decode()))

    def close(self):
        if getattr(self, "ctx", None):
            self.lib.gcp_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _get_scratch(self, P):
        need = self.lib.gcp_eval_scratch_bytes(self.ctx, P)
        if self._scratch is None or self._scratch.numel() < need:
            self._scratch = torch.empty(need, dtype=torch.uint8, device=self.device)
        return self._scratch

    # ------------------------------------------------------------------
    def to_device_dna(self, dna):
        """Accepts numpy / torch int arrays [P,A,L]; returns contiguous int32 on device."""
        if isinstance(dna, np.ndarray):
            dna = torch.from_numpy(np.ascontiguousarray(dna, dtype=np.int32))
        dna = dna.to(device=self.device, dtype=torch.int32).contiguous()
        if dna.dim() != 3 or dna.shape[1] != self.A or dna.shape[2] != self.L:
            raise ValueError("dna must be [P,%d,%d], got %s" % (self.A, self.L, tuple(dna.shape)))
        return dna

    def evaluate(self, dna, detail=True, pixel_counts=True):
        """Organism.calcObj for every organism.  dna: [P,A,L] int32 device tensor (or
        anything to_device_dna accepts).  Returns EvalResult of device tensors.
        detail: also per-agent costs, -coverage before the feasibility zeroing, loop-closure
        matrix; pixel_counts (with detail): the six pixel counters, which selects the general
        coverage kernel instead of the objectives-only fast path."""
        dna = self.to_device_dna(dna)
        P = dna.shape[0]
        dev, A = self.device, self.A
        res = EvalResult()
        res.obj = torch.empty((P, 2), dtype=torch.float64, device=dev)
        res.flags = torch.zeros((P, 4), dtype=torch.uint8, device=dev)
        out = GcpEvalOut()
        out.obj = res.obj.data_ptr()
        out.flags = res.flags.data_ptr()
        if detail:
            res.neg_cov = torch.empty(P, dtype=torch.float64, device=dev)
            res.dist = torch.empty((P, A), dtype=torch.float64, device=dev)
            res.turn = torch.empty((P, A), dtype=torch.float64, device=dev)
            res.lc_mat = torch.empty((P, A, A), dtype=torch.int32, device=dev)
            out.neg_cov, out.dist, out.turn = res.neg_cov.data_ptr(), res.dist.data_ptr(), res.turn.data_ptr()
            out.lc_mat = res.lc_mat.data_ptr()
            res.cov = None
            if pixel_counts:
                res.cov = torch.empty((P, 8), dtype=torch.int32, device=dev)
                out.cov = res.cov.data_ptr()
        else:
            res.neg_cov = res.dist = res.turn = res.cov = res.lc_mat = None
        scratch = self._get_scratch(P)
        self._check(self.lib.gcp_eval(self.ctx, C.c_void_p(dna.data_ptr()), P, C.byref(out),
                                      C.c_void_p(scratch.data_ptr()), scratch.numel(),
                                      self._stream()))
        return res

    def raise_on_error(self, result=None, flags=None):
        """The kernels report a gene outside [-1, N) (bit 0) or a coverage pass that ran out of
        hash partitions (bit 1; its coverage is forced to 0) in flags[:, 3].  Raises GcpError if
        any organism carries one.  Synchronises with the device: call it off the hot loop."""
        f = result.flags if result is not None else flags
        if f is None:
            return
        e = f[:, 3]
        if bool((e != 0).any().item()):
            k = int(torch.nonzero(e)[0].item())
            raise GcpError("evaluation error flags set for %d organism(s), first at %d: %s"
                           % (int((e != 0).sum().item()), k,
                              "gene out of range" if int(e[k].item()) & 1 else "coverage pass overflow"))

    def evaluate_host(self, dna_host, obj_host=None, flags_host=None):
        """End-to-end call with HOST buffers (the reference-facing shape of the path):
        dna_host int32 [P,A,L] (pinned torch tensor or numpy); returns (obj, flags) on host."""
        if isinstance(dna_host, np.ndarray):
            dna_host = torch.from_numpy(np.ascontiguousarray(dna_host, dtype=np.int32))
        assert dna_host.dtype in (torch.int32, torch.int16) and dna_host.is_contiguous() \
            and not dna_host.is_cuda
        P = dna_host.shape[0]
        if obj_host is None:
            obj_host = torch.empty((P, 2), dtype=torch.float64).pin_memory()
        if flags_host is None:
            flags_host = torch.empty((P, 4), dtype=torch.uint8).pin_memory()
        fn = self.lib.gcp_eval_host16 if dna_host.dtype == torch.int16 else self.lib.gcp_eval_host
        self._check(fn(self.ctx, C.c_void_p(dna_host.data_ptr()), P,
                       C.c_void_p(obj_host.data_ptr()), C.c_void_p(flags_host.data_ptr())))
        return obj_host, flags_host

    def evaluate16(self, dna16):
        """Objectives from an int16 device chromosome tensor [P,A,L] (N <= 32767): fused path only."""
        assert dna16.is_cuda and dna16.dtype == torch.int16 and dna16.is_contiguous()
        P = dna16.shape[0]
        res = EvalResult()
        res.obj = torch.empty((P, 2), dtype=torch.float64, device=self.device)
        res.flags = torch.zeros((P, 4), dtype=torch.uint8, device=self.device)
        res.neg_cov = res.dist = res.turn = res.cov = res.lc_mat = None
        out = GcpEvalOut()
        out.obj = res.obj.data_ptr()
        out.flags = res.flags.data_ptr()
        self._check(self.lib.gcp_eval16(self.ctx, C.c_void_p(dna16.data_ptr()), P, C.byref(out),
                                        self._stream()))
        return res

    # ---- individual stages (tests / profiling) ------------------------------
    def path_costs(self, dna, with_step_info=False):
        dna = self.to_device_dna(dna)
        P, dev = dna.shape[0], self.device
        edge_ids = torch.empty((P, self.A, self.L), dtype=torch.int32, device=dev)
        step_info = (torch.empty((P, self.A, self.L), dtype=torch.int32, device=dev)
                     if with_step_info else None)
        dist = torch.empty((P, self.A), dtype=torch.float64, device=dev)
        turn = torch.empty((P, self.A), dtype=torch.float64, device=dev)
        flags = torch.zeros((P, 4), dtype=torch.uint8, device=dev)
        self._check(self.lib.gcp_path_costs(
            self.ctx, C.c_void_p(dna.data_ptr()), P, C.c_void_p(edge_ids.data_ptr()),
            C.c_void_p(step_info.data_ptr()) if with_step_info else None,
            C.c_void_p(dist.data_ptr()), C.c_void_p(turn.data_ptr()),
            C.c_void_p(flags.data_ptr()), self._stream()))
        if with_step_info:
            return edge_ids, step_info, dist, turn, flags
        return edge_ids, dist, turn, flags

    def coverage_masked(self, step_info, cov=None, flags=None):
        P = step_info.shape[0]
        if cov is None:
            cov = torch.empty((P, 8), dtype=torch.int32, device=self.device)
        if flags is None:
            flags = torch.zeros((P, 4), dtype=torch.uint8, device=self.device)
        self._check(self.lib.gcp_coverage_masked(
            self.ctx, C.c_void_p(step_info.data_ptr()), P, C.c_void_p(cov.data_ptr()),
            C.c_void_p(flags.data_ptr()), self._stream()))
        return cov, flags

    def coverage(self, edge_ids, cov=None, flags=None):
        P = edge_ids.shape[0]
        if cov is None:
            cov = torch.empty((P, 8), dtype=torch.int32, device=self.device)
        if flags is None:
            flags = torch.zeros((P, 4), dtype=torch.uint8, device=self.device)
        self._check(self.lib.gcp_coverage(
            self.ctx, C.c_void_p(edge_ids.data_ptr()), P, C.c_void_p(cov.data_ptr()),
            C.c_void_p(flags.data_ptr()), self._stream()))
        return cov, flags

    def loop_closures(self, dna, dist, turn, cov, flags=None):
        dna = self.to_device_dna(dna)
        P, dev, A = dna.shape[0], self.device, self.A
        lc = torch.empty((P, A, A), dtype=torch.int32, device=dev)
        obj = torch.empty((P, 2), dtype=torch.float64, device=dev)
        neg = torch.empty(P, dtype=torch.float64, device=dev)
        if flags is None:
            flags = torch.zeros((P, 4), dtype=torch.uint8, device=dev)
        self._check(self.lib.gcp_loop_closures(
            self.ctx, C.c_void_p(dna.data_ptr()), P, C.c_void_p(dist.data_ptr()),
            C.c_void_p(turn.data_ptr()), C.c_void_p(cov.data_ptr()), C.c_void_p(lc.data_ptr()),
            C.c_void_p(obj.data_ptr()), C.c_void_p(neg.data_ptr()), C.c_void_p(flags.data_ptr()),
            self._stream()))
        return lc, obj, neg, flags

    # ------------------------------------------------------------------
    def _get_rank_scratch(self, n):
        need = self.lib.gcp_rank_scratch_bytes(self.ctx, n)
        if self._rank_scratch is None or self._rank_scratch.numel() < need:
            self._rank_scratch = torch.empty(need, dtype=torch.uint8, device=self.device)
        return self._rank_scratch

    def maximin(self, obj, coverage_constr, row_range=None, return_scaled=False):
        """GeneticalGorithm.rackNstack: obj [n,2] float64 device tensor -> fit [n]."""
        obj = obj.to(device=self.device, dtype=torch.float64).contiguous()
        n = obj.shape[0]
        r0, r1 = (0, n) if row_range is None else row_range
        fit = torch.empty(n, dtype=torch.float64, device=self.device)
        sc = torch.empty((n, 2), dtype=torch.float64, device=self.device) if return_scaled else None
        s = self._get_rank_scratch(n)
        self._check(self.lib.gcp_maximin(
            self.ctx, C.c_void_p(obj.data_ptr()), n, float(coverage_constr), r0, r1,
            C.c_void_p(sc.data_ptr()) if sc is not None else None, C.c_void_p(fit.data_ptr()),
            C.c_void_p(s.data_ptr()), s.numel(), self._stream()))
        return (fit, sc) if return_scaled else fit

    def arena_order(self, fit):
        """got.sortBy canonical order: stable ascending by (fit, index)."""
        fit = fit.to(device=self.device, dtype=torch.float64).contiguous()
        n = fit.shape[0]
        order = torch.empty(n, dtype=torch.int32, device=self.device)
        nondom = torch.empty(n, dtype=torch.uint8, device=self.device)
        s = self._get_rank_scratch(n)
        self._check(self.lib.gcp_arena_order(
            self.ctx, C.c_void_p(fit.data_ptr()), n, C.c_void_p(order.data_ptr()),
            C.c_void_p(nondom.data_ptr()), C.c_void_p(s.data_ptr()), s.numel(), self._stream()))
        return order, nondom

    # ---- selection + variation ------------------------------------------------
    def set_vary_params(self, org_params, path_params=None):
        """org_params / path_params dicts of pather_driver.py:173-191."""
        pp = path_params or {}
        v = GcpVaryParams(int(org_params['min_dna_len']), int(org_params['crossover_time_thresh']),
                          int(org_params['num_muterpolations']),
                          int(org_params['muterpolation_srch_dist']),
                          int(pp.get('path_memory', 10)), 0,
                          float(org_params['crossover_prob']), float(org_params['mutation_prob']),
                          float(org_params['muterpolate_prob']),
                          float(org_params['muterpolation_sub_prob']),
                          float(pp.get('log_scale_weight', 0.7)))
        self._check(self.lib.gcp_set_vary_params(self.ctx, C.byref(v)))

    def low_var_select(self, fit, gamma, r0):
        """got.lowVarSample -> indices of the resampled parents (device int32 [M])."""
        fit = fit.to(device=self.device, dtype=torch.float64).contiguous()
        M = fit.shape[0]
        scratch = torch.empty(M + 2 * ((M + 1023) // 1024), dtype=torch.float64, device=self.device)
        out = torch.empty(M, dtype=torch.int32, device=self.device)
        self._check(self.lib.gcp_low_var_select(
            self.ctx, C.c_void_p(fit.data_ptr()), M, float(gamma), float(r0),
            C.c_void_p(scratch.data_ptr()), C.c_void_p(out.data_ptr()), self._stream()))
        return out

    def crossover(self, parents, strong, seed, gen):
        parents = self.to_device_dna(parents)
        strong = strong.to(device=self.device, dtype=torch.int32).contiguous()
        P = strong.shape[0]
        kids = torch.empty((P, self.A, self.L), dtype=torch.int32, device=self.device)
        self._check(self.lib.gcp_crossover(
            self.ctx, C.c_void_p(parents.data_ptr()), C.c_void_p(strong.data_ptr()), P,
            int(seed), int(gen), C.c_void_p(kids.data_ptr()), self._stream()))
        return kids

    def mutate(self, dna, seed, gen, salt=0):
        """In place on a device int32 [P,A,L] tensor: the Organism constructor's variation."""
        assert dna.is_cuda and dna.dtype == torch.int32 and dna.is_contiguous()
        self._check(self.lib.gcp_mutate(self.ctx, C.c_void_p(dna.data_ptr()), dna.shape[0],
                                        int(seed), int(gen), int(salt), self._stream()))
        return dna

    def select_vary(self, parents, parent_fit, gamma, seed, gen):
        """One generation's children: lowVarSample + crossover + constructor variation."""
        parents = self.to_device_dna(parents)
        parent_fit = parent_fit.to(device=self.device, dtype=torch.float64).contiguous()
        P = parents.shape[0]
        kids = torch.empty_like(parents)
        scratch = torch.empty(P * 16 + 64, dtype=torch.uint8, device=self.device)
        self._check(self.lib.gcp_select_vary(
            self.ctx, C.c_void_p(parents.data_ptr()), C.c_void_p(parent_fit.data_ptr()), P,
            float(gamma), int(seed), int(gen), C.c_void_p(kids.data_ptr()),
            C.c_void_p(scratch.data_ptr()), scratch.numel(), self._stream()))
        return kids

    def random_walks(self, n_org, start_idx, path_len, seed, batch=0):
        start = torch.as_tensor(np.asarray(start_idx, dtype=np.int32), device=self.device)
        assert start.numel() == self.A
        dna = torch.empty((n_org, self.A, self.L), dtype=torch.int32, device=self.device)
        self._check(self.lib.gcp_random_walks(
            self.ctx, C.c_void_p(dna.data_ptr()), n_org, C.c_void_p(start.data_ptr()),
            int(path_len), int(seed), int(batch), self._stream()))
        return dna

    def gather_survivors(self, order, par_dna, kid_dna, glad_obj, glad_fit):
        P = par_dna.shape[0]
        out_dna = torch.empty_like(par_dna)
        out_obj = torch.empty((P, 2), dtype=torch.float64, device=self.device)
        out_fit = torch.empty(P, dtype=torch.float64, device=self.device)
        self._check(self.lib.gcp_gather_survivors(
            self.ctx, C.c_void_p(order.data_ptr()), P, C.c_void_p(par_dna.data_ptr()),
            C.c_void_p(kid_dna.data_ptr()), C.c_void_p(glad_obj.data_ptr()),
            C.c_void_p(glad_fit.data_ptr()), C.c_void_p(out_dna.data_ptr()),
            C.c_void_p(out_obj.data_ptr()), C.c_void_p(out_fit.data_ptr()), self._stream()))
        return out_dna, out_obj, out_fit

    # ---- population shards (multi-GPU generations) ---------------------------
    def gene16_available(self):
        """int16 chromosomes end to end (N <= 32767 and the fused objectives path)."""
        return bool(self.lib.gcp_eval16_available(self.ctx))

    def shard_alloc(self, nbytes):
        return ShardBuffer(self, nbytes)

    def shard_open(self, handle):
        """Map a peer's ShardBuffer from its 64-byte IPC handle -> device pointer (int)."""
        h = (C.c_uint8 * 64).from_buffer_copy(bytes(handle))
        ptr = C.c_void_p()
        self._check(self.lib.gcp_shard_open(self.ctx, h, C.byref(ptr)))
        return int(ptr.value)

    @staticmethod
    def shard_table(ptrs, per_shard):
        t = GcpShards()
        t.world, t.per_shard = len(ptrs), int(per_shard)
        for w, p in enumerate(ptrs):
            t.base[w] = int(p)
        return t

    def evaluate_genes(self, dna, out_obj=None, out_flags=None):
        """Objectives of a device chromosome tensor [P,A,L], int16 (fused path) or int32."""
        if dna.dtype == torch.int32:
            r = self.evaluate(dna, detail=False)
            if out_obj is not None:
                out_obj.copy_(r.obj)
                r.obj = out_obj
            return r
        assert dna.is_cuda and dna.dtype == torch.int16 and dna.is_contiguous()
        P = dna.shape[0]
        res = EvalResult()
        res.obj = out_obj if out_obj is not None else torch.empty((P, 2), dtype=torch.float64, device=self.device)
        res.flags = out_flags if out_flags is not None else torch.zeros((P, 4), dtype=torch.uint8, device=self.device)
        res.neg_cov = res.dist = res.turn = res.cov = res.lc_mat = None
        out = GcpEvalOut()
        out.obj = res.obj.data_ptr()
        out.flags = res.flags.data_ptr()
        self._check(self.lib.gcp_eval16(self.ctx, C.c_void_p(dna.data_ptr()), P, C.byref(out),
                                        self._stream()))
        return res

    def crossover_shards(self, parents_tab, gene_bytes, strong, n_org_global, pair_begin, n_pairs,
                         seed, gen, children):
        """children: device gene tensor [2 * n_pairs, A, L] (written)."""
        self._check(self.lib.gcp_crossover_shards(
            self.ctx, C.byref(parents_tab), int(gene_bytes), C.c_void_p(strong.data_ptr()),
            int(n_org_global), int(pair_begin), int(n_pairs), int(seed), int(gen),
            C.c_void_p(children.data_ptr()), self._stream()))
        return children

    def mutate_genes(self, dna, org_offset, seed, gen, salt=0):
        assert dna.is_cuda and dna.is_contiguous() and dna.dtype in (torch.int16, torch.int32)
        self._check(self.lib.gcp_mutate_genes(
            self.ctx, C.c_void_p(dna.data_ptr()), dna.element_size(), dna.shape[0], int(org_offset),
            int(seed), int(gen), int(salt), self._stream()))
        return dna

    def gather_shards(self, order, n_org_global, k_begin, n_local, parents_tab, children_tab,
                      gene_bytes, glad_obj, glad_fit, out_dna, out_obj, out_fit):
        self._check(self.lib.gcp_gather_shards(
            self.ctx, C.c_void_p(order.data_ptr()), int(n_org_global), int(k_begin), int(n_local),
            C.byref(parents_tab), C.byref(children_tab), int(gene_bytes),
            C.c_void_p(glad_obj.data_ptr()), C.c_void_p(glad_fit.data_ptr()),
            C.c_void_p(out_dna.data_ptr()),
            C.c_void_p(out_obj.data_ptr()) if out_obj is not None else None,
            C.c_void_p(out_fit.data_ptr()) if out_fit is not None else None, self._stream()))

    def peer_barrier(self, flags_tab, rank, epoch, timed_out):
        self._check(self.lib.gcp_peer_barrier(self.ctx, C.byref(flags_tab), int(rank), int(epoch),
                                              C.c_void_p(timed_out.data_ptr()), self._stream()))

    # ------------------------------------------------------------------
    def set_option(self, name, value):
        self._check(self.lib.gcp_set_option(self.ctx, name.encode(), int(value)))

    def launch_count(self):
        return int(self.lib.gcp_launch_count(self.ctx))

    def set_timing(self, on):
        self.lib.gcp_set_timing(self.ctx, 1 if on else 0)

    def stage_ms(self, stage):
        return float(self.lib.gcp_stage_ms(self.ctx, stage))
