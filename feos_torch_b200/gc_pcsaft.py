"""``GcPcSaftMix`` / ``GcPcSaft`` -- heterosegmented group-contribution PC-SAFT for binary mixtures with
the public API of the reference (`/root/reference/feos_torch/gc_pcsaft.py:13-528`, native class
`/root/reference/src/gc_pcsaft.rs:15-100`), computed by the sm_100a kernels behind
``include/feos_b200.h``.

The constructor replaces the reference's dense count tensors ([N,2,S] segments, [N,2,S,S] bonds:
8.5 KB per state for the 23-segment table) by a de-duplicated table of molecule pairs
(``feos_gc_pair`` records listing only the segment types and bonded type pairs that are present) and a
state -> pair index.  Gradients flow to the eight per-segment parameter vectors, the k_ab values of the
binary segment records, ``phi`` and the temperature; segment-parameter and k_ab gradients are summed
over the states inside the backward kernel.
"""
import ctypes

import numpy as np
import torch

from . import _lib
from .pcsaft_pure import _back, _on_device, _cuda_device, _prep, _stream, _upstream, compact

MAX_SEG = 12
MAX_BOND = 16
PAIR_DTYPE = np.dtype(
    [("nseg", "<i4"), ("nbond", "<i4"), ("seg", "<i4", (MAX_SEG,)), ("cnt", "<f8", (2, MAX_SEG)), ("ba", "<i4", (MAX_BOND,)),
     ("bb", "<i4", (MAX_BOND,)), ("bcnt", "<f8", (2, MAX_BOND)), ("krec", "i1", (MAX_SEG, MAX_SEG))],
    align=True,
)


def build_pair_table(segment_identifier, segment_lists, bond_lists, binary_segment_records):
    """De-duplicate the molecule pairs of a batch.

    Returns (pairs: structured array [U] of ``feos_gc_pair``, pair_index: int32 [N]).  Follows the
    counting rules of gc_pcsaft.py:24-52 (segment counts per type, bonds counted per unordered type
    pair) restricted to the types present in the pair."""
    index = {s: i for i, s in enumerate(segment_identifier)}
    records = [(index[s1], index[s2]) for s1, s2, _ in binary_segment_records]
    table, keys, by_id = [], {}, {}
    pair_index = np.empty(len(segment_lists), dtype=np.int32)
    for n, (segs, bonds) in enumerate(zip(segment_lists, bond_lists)):
        ident = (id(segs), id(bonds))
        u = by_id.get(ident)
        if u is None:
            key = (tuple(tuple(s) for s in segs), tuple(tuple(tuple(b) for b in bl) for bl in bonds))
            u = keys.get(key)
            if u is None:
                u = len(table)
                keys[key] = u
                table.append(_make_pair(index, segs, bonds, records))
            by_id[ident] = u
        pair_index[n] = u
    pairs = np.zeros(len(table), dtype=PAIR_DTYPE)
    for u, rec in enumerate(table):
        pairs[u] = rec
    return pairs, pair_index


def _make_pair(index, segs, bonds, records):
    if len(segs) != 2 or len(bonds) != 2:
        raise ValueError("group-contribution PC-SAFT is implemented for binary mixtures")
    present = []
    for mol in segs:
        for s in mol:
            g = index[s]
            if g not in present:
                present.append(g)
    if len(present) > MAX_SEG:
        raise ValueError(f"more than {MAX_SEG} distinct segment types in one molecule pair")
    local = {g: a for a, g in enumerate(present)}
    rec = np.zeros((), dtype=PAIR_DTYPE)
    rec["nseg"] = len(present)
    rec["seg"][: len(present)] = present
    for c, mol in enumerate(segs):
        for s in mol:
            rec["cnt"][c, local[index[s]]] += 1.0
    btypes = []
    for c, (mol, bl) in enumerate(zip(segs, bonds)):
        for i, j in bl:
            a, b = local[index[mol[i]]], local[index[mol[j]]]
            t = (max(a, b), min(a, b))
            if t not in btypes:
                if len(btypes) == MAX_BOND:
                    raise ValueError(f"more than {MAX_BOND} distinct bonded segment-type pairs in one molecule pair")
                btypes.append(t)
            rec["bcnt"][c, btypes.index(t)] += 1.0
    rec["nbond"] = len(btypes)
    for k, (a, b) in enumerate(btypes):
        rec["ba"][k], rec["bb"][k] = a, b
    rec["krec"][:] = -1
    for r, (g1, g2) in enumerate(records):
        if g1 in local and g2 in local:
            rec["krec"][local[g1], local[g2]] = r
            rec["krec"][local[g2], local[g1]] = r
    return rec


SLIM_MAX_BONDS = 8  # csrc/gc_pcsaft.cuh: kGcSlimBond
PLAIN_ONLY = 4      # include/feos_b200.h: FEOS_PLAIN_ONLY


def _check_association(pairs, seg_params):
    """gc_pcsaft.py:76-80: at most one associating segment per component.  Returns whether EVERY pair is of the
    kernels' "slim" class (no association term, at most SLIM_MAX_BONDS bond types: csrc/gc_pcsaft.cuh gc_is_slim) --
    the forward entry point then skips the launch of the general instantiation (FEOS_PLAIN_ONLY)."""
    host = seg_params.detach().cpu().numpy()
    kap_eps = host[:, 4] * host[:, 5]
    all_slim = True
    for rec in pairs:
        ns = int(rec["nseg"])
        seg = rec["seg"][:ns]
        for c in range(2):
            cnt = rec["cnt"][c, :ns]
            if float((cnt * np.sign(kap_eps[seg])).sum()) > 1:
                raise Exception("Only up to one associating segment per component is allowed!")
            if float((cnt * host[seg, 4]).sum()) * float((cnt * host[seg, 5]).sum()) != 0.0:
                all_slim = False
        if int(rec["nbond"]) > SLIM_MAX_BONDS:
            all_slim = False
    return all_slim


class _GcPoint(torch.autograd.Function):
    @staticmethod
    def forward(ctx, seg_params, kab, phi, temperature, molefracs, pressure, bubble, tables, orders):
        pairs_dev, pair_index, n_pairs, flags = tables
        order, order_bwd = orders
        n = temperature.shape[0]
        dev = temperature.device
        rho = torch.empty((n, 4), dtype=torch.float64, device=dev)
        val = torch.empty((n,), dtype=torch.float64, device=dev)
        failed = torch.empty((n,), dtype=torch.uint8, device=dev)
        with _on_device(dev):
            _lib.timed_call(
                "feos_dev_gc_point_fwd", _lib.ptr(pairs_dev), n_pairs, _lib.ptr(seg_params), seg_params.shape[0],
                _lib.ptr(kab), kab.shape[0], _lib.ptr(pair_index), _lib.ptr(phi), _lib.ptr(temperature),
                _lib.ptr(molefracs), _lib.ptr(pressure), n, int(bubble), flags, _lib.ptr(rho), _lib.ptr(val), _lib.ptr(failed),
                None, _lib.ptr(order), _stream(),
            )
            out, pos = compact(failed, val)
        ctx.save_for_backward(seg_params, kab, phi, temperature, rho, failed, pos, pairs_dev, pair_index, order_bwd)
        ctx.bubble, ctx.n_pairs = int(bubble), n_pairs
        ctx.mark_non_differentiable(rho, failed)
        ctx.set_materialize_grads(False)  # no zero-filled gradients for the mask / density outputs
        return out, failed, rho

    @staticmethod
    def backward(ctx, g_val, _g_failed, _g_rho):
        seg_params, kab, phi, temperature, rho, failed, pos, pairs_dev, pair_index, order = ctx.saved_tensors
        n = temperature.shape[0]
        if g_val is None:  # only the mask / density outputs were used downstream
            return (None,) * 9
        g, bc = _upstream(g_val)
        need_seg, need_kab, need_phi, need_t = ctx.needs_input_grad[:4]
        g_seg = torch.zeros_like(seg_params) if need_seg else None
        g_kab = torch.zeros_like(kab) if need_kab and kab.shape[0] else None
        g_phi = torch.empty_like(phi) if need_phi and phi is not None else None
        g_t = torch.empty_like(temperature) if need_t else None
        with _on_device(temperature.device):
            _lib.timed_call(
                "feos_dev_gc_point_bwd", _lib.ptr(pairs_dev), ctx.n_pairs, _lib.ptr(seg_params), seg_params.shape[0],
                _lib.ptr(kab), kab.shape[0], _lib.ptr(pair_index), _lib.ptr(phi), _lib.ptr(temperature), _lib.ptr(rho),
                _lib.ptr(failed), _lib.ptr(pos), _lib.ptr(g), bc, n, ctx.bubble, _lib.ptr(g_seg), _lib.ptr(g_kab),
                _lib.ptr(g_phi), _lib.ptr(g_t), _lib.ptr(order), _stream(),
            )
        if need_kab and g_kab is None:
            g_kab = torch.zeros_like(kab)
        return g_seg, g_kab, g_phi, g_t, None, None, None, None, None


class _GcState(torch.autograd.Function):
    """(a, p, mu, v) = derivatives(T, rho [N,2]); backward = dual passes of csrc/state_grad.cuh."""

    @staticmethod
    def forward(ctx, seg_params, kab, phi, temperature, density, tables):
        pairs_dev, pair_index, n_pairs = tables
        n = temperature.shape[0]
        dev = temperature.device
        a = torch.empty((n,), dtype=torch.float64, device=dev)
        p = torch.empty((n,), dtype=torch.float64, device=dev)
        mu = torch.empty((n, 2), dtype=torch.float64, device=dev)
        v = torch.empty((n, 2), dtype=torch.float64, device=dev)
        with _on_device(dev):
            _lib.call("feos_dev_gc_state", _lib.ptr(pairs_dev), n_pairs, _lib.ptr(seg_params), seg_params.shape[0],
                      _lib.ptr(kab), kab.shape[0], _lib.ptr(pair_index), _lib.ptr(phi), _lib.ptr(temperature),
                      _lib.ptr(density), n, _lib.ptr(a), _lib.ptr(p), _lib.ptr(mu), _lib.ptr(v), _stream())
        ctx.save_for_backward(seg_params, kab, phi, temperature, density, pairs_dev, pair_index)
        ctx.n_pairs = n_pairs
        return a, p, mu, v

    @staticmethod
    def backward(ctx, ga, gp, gmu, gv):
        seg_params, kab, phi, temperature, density, pairs_dev, pair_index = ctx.saved_tensors
        n = temperature.shape[0]
        ga, gp, gmu, gv = (None if g is None else g.contiguous() for g in (ga, gp, gmu, gv))
        need_seg, need_kab, need_phi, need_t, need_r = ctx.needs_input_grad[:5]
        g_seg = torch.zeros_like(seg_params) if need_seg else None
        g_kab = torch.zeros_like(kab) if need_kab and kab.shape[0] else None
        g_phi = torch.empty_like(phi) if need_phi and phi is not None else None
        g_t = torch.empty_like(temperature) if need_t else None
        g_r = torch.empty_like(density) if need_r else None
        if n:
            with _on_device(temperature.device):
                _lib.call("feos_dev_gc_state_bwd", _lib.ptr(pairs_dev), ctx.n_pairs, _lib.ptr(seg_params), seg_params.shape[0],
                          _lib.ptr(kab), kab.shape[0], _lib.ptr(pair_index), _lib.ptr(phi), _lib.ptr(temperature),
                          _lib.ptr(density), n, _lib.ptr(ga), _lib.ptr(gp), _lib.ptr(gmu), _lib.ptr(gv), _lib.ptr(g_seg),
                          _lib.ptr(g_kab), _lib.ptr(g_phi), _lib.ptr(g_t), _lib.ptr(g_r), _stream())
        if need_kab and g_kab is None:
            g_kab = torch.zeros_like(kab)
        return g_seg, g_kab, g_phi, g_t, g_r, None


class GcPcSaftMix:
    """GcPcSaftMix(segment_identifier, parameter, segment_lists, bond_lists, binary_segment_records, phi=None)
    -- gc_pcsaft.py:14-22.  ``parameter`` = 8 tensors of shape [S] (m, sigma, epsilon_k, mu, kappa_ab,
    epsilon_k_ab, na, nb); ``segment_lists[n] = [segments of molecule 0, segments of molecule 1]``;
    ``bond_lists[n][c]`` = index pairs into ``segment_lists[n][c]``; ``binary_segment_records`` =
    [(segment, segment, k_ab)] with k_ab a float or a 0-dim tensor; ``phi`` [N, 2]."""

    def __init__(self, segment_identifier, parameter, segment_lists, bond_lists, binary_segment_records, phi=None):
        _lib.load()
        if len(parameter) != 8:
            raise ValueError("parameter must hold 8 tensors (m, sigma, epsilon_k, mu, kappa_ab, epsilon_k_ab, na, nb)")
        ref = parameter[0]
        self._out_device = ref.device
        self._dev = _cuda_device(ref)
        self._seg = torch.stack([_prep(p, self._dev) for p in parameter], dim=1).contiguous()  # [S, 8]
        ks = [k if isinstance(k, torch.Tensor) else torch.tensor(float(k), dtype=torch.float64) for _, _, k in binary_segment_records]
        self._kab = (torch.stack([_prep(k.reshape(()), self._dev) for k in ks]) if ks
                     else torch.zeros((0,), dtype=torch.float64, device=self._dev))
        pairs, pair_index = build_pair_table(segment_identifier, segment_lists, bond_lists, binary_segment_records)
        self._flags = PLAIN_ONLY if _check_association(pairs, self._seg) else 0
        self._n = len(segment_lists)
        self._pairs_host, self._pair_index_host = pairs, pair_index
        self._pairs = torch.from_numpy(pairs.view(np.uint8).reshape(len(pairs), PAIR_DTYPE.itemsize).copy()).to(self._dev)
        self._pair_index = torch.from_numpy(pair_index).to(self._dev)
        self._n_pairs = len(pairs)
        self._phi = None if phi is None else _prep(phi, self._dev)
        self.phi = phi
        self.gc_pcsaft = GcPcSaft._from_tables(self)

    def _back(self, t):
        return _back(t, self._out_device)

    def _point(self, temperature, molefracs, pressure, bubble):
        for name, x in (("temperature", temperature), ("molefracs", molefracs), ("pressure", pressure)):
            if x is not None and (x.dim() != 1 or x.shape[0] != self._n):
                raise ValueError(f"{name} must have shape [{self._n}] (one entry per state of the constructor lists; the model "
                                 f"is not reduced after failed states, unlike the reference's `reduce`), got {tuple(x.shape)}")
        t = _prep(temperature, self._dev)
        z = _prep(molefracs.detach(), self._dev)
        p0 = None if pressure is None else _prep(pressure.detach(), self._dev)
        tables = (self._pairs, self._pair_index, self._n_pairs, self._flags)
        orders = (None, None)
        n = t.shape[0]
        if _lib.state_order_enabled(n):
            # forward: states of one molecule pair and neighbouring (T, x) share a warp (csrc/order.cuh); backward
            # (no iteration): grouped by pair only, input order inside a pair keeps its accesses local.  The data
            # set is recognised by the caller's (T, x) tensor objects in the next iteration of a fit loop.
            def compute(pair_only, out):
                o = out if out is not None else torch.empty((n,), dtype=torch.int32, device=self._dev)
                with _on_device(self._dev):
                    _lib.timed_call("feos_dev_gc_point_order", _lib.ptr(self._pair_index), self._n_pairs, _lib.ptr(t),
                                    _lib.ptr(z), n, pair_only, _lib.ptr(o), _stream())
                return o
            key = (self._pair_index, temperature, molefracs)
            dev = self._dev.index
            orders = (_lib.cached_order(("gc", dev), key, lambda out: compute(0, out), n=n),
                      _lib.cached_order(("gc-pair", dev), key, lambda out: compute(1, out), n=n))
        val, failed, _ = _GcPoint.apply(self._seg, self._kab, self._phi, t, z, p0, bubble, tables, orders)
        return self._back(val), self._back(failed.view(torch.bool))

    def bubble_point(self, temperature, liquid_molefracs, pressure):
        """(p_bubble [n_ok] in Pa, failed mask [N]) -- gc_pcsaft.py:470-490.  `pressure`: the reference's starting
        value (src/gc_pcsaft.rs:125-132), used by the retry path where it cannot build a start of its own."""
        return self._point(temperature, liquid_molefracs, pressure, True)

    def dew_point(self, temperature, vapor_molefracs, pressure):
        """(p_dew [n_ok] in Pa, failed mask [N]) -- gc_pcsaft.py:492-512."""
        return self._point(temperature, vapor_molefracs, pressure, False)

    def derivatives(self, temperature, density):
        """(a [N], p/kT [N], mu_i [N,2], v_i [N,2]) at (T, rho [N,2]) -- gc_pcsaft.py:443-468; carries gradients
        w.r.t. the segment parameters, k_ab, phi, T and rho."""
        if temperature.shape[0] != self._n or density.shape != (self._n, 2):
            raise ValueError(f"temperature / density must have shapes [{self._n}] / [{self._n}, 2]")
        t = _prep(temperature, self._dev)
        r = _prep(density, self._dev)
        tables = (self._pairs, self._pair_index, self._n_pairs)
        a, p, mu, v = _GcState.apply(self._seg, self._kab, self._phi, t, r, tables)
        return self._back(a), self._back(p), self._back(mu), self._back(v)

    def helmholtz_energy_density(self, temperature, density, kab=None):
        """a = A/(V kT) [N,1] -- gc_pcsaft.py:116-253.  `kab` is part of the reference's signature (gc_pcsaft.py:116)
        but never read there (the dispersion term uses `self.kab`, gc_pcsaft.py:188): it has no effect here either."""
        return self.derivatives(temperature, density)[0][:, None]


class GcPcSaft:
    """numpy-in / numpy-out class with the signature of the reference's native ``GcPcSaft``
    (src/gc_pcsaft.rs:25-99): ``GcPcSaft(segment_records, segments, bonds, binary_segment_records, phi)``
    with ``segment_records = [(identifier, f64[8])]``; ``bubble_point`` / ``dew_point(T, z1, p0)`` ->
    ``(rho f64[n_ok,4], failed bool[n])``."""

    def __init__(self, segment_records, segments, bonds, binary_segment_records, phi):
        _lib.load()
        ident = [r[0] for r in segment_records]
        self._seg = np.ascontiguousarray(np.stack([np.asarray(r[1], dtype=np.float64) for r in segment_records]))
        self._kab = np.ascontiguousarray(np.array([float(k) for _, _, k in binary_segment_records], dtype=np.float64))
        self._pairs, self._pair_index = build_pair_table(ident, segments, bonds, binary_segment_records)
        self._phi = None if phi is None else np.ascontiguousarray(phi, dtype=np.float64)

    @classmethod
    def _from_tables(cls, mix):
        self = cls.__new__(cls)
        self._seg = mix._seg.detach().cpu().numpy()
        self._kab = mix._kab.detach().cpu().numpy()
        self._pairs, self._pair_index = mix._pairs_host, mix._pair_index_host
        self._phi = None if mix._phi is None else mix._phi.detach().cpu().numpy()
        return self

    def _point(self, temperature, molefracs, pressure, bubble):
        t = np.ascontiguousarray(temperature, dtype=np.float64)
        z = np.ascontiguousarray(molefracs, dtype=np.float64)
        pr = np.ascontiguousarray(pressure, dtype=np.float64)
        n = t.shape[0]
        if n != self._pair_index.shape[0]:
            raise ValueError("temperature must have one entry per state of the constructor lists")
        rho = np.empty((n, 4), dtype=np.float64)
        status = np.empty((n,), dtype=np.uint8)
        n_ok = ctypes.c_int64(0)
        _lib.call("feos_host_gc_point", _lib.ptr(self._pairs), len(self._pairs), _lib.ptr(self._seg), self._seg.shape[0],
                  _lib.ptr(self._kab), self._kab.shape[0], _lib.ptr(self._pair_index), _lib.ptr(self._phi), _lib.ptr(t),
                  _lib.ptr(z), _lib.ptr(pr), n, int(bubble), _lib.ptr(rho), _lib.ptr(status),
                  ctypes.cast(ctypes.byref(n_ok), ctypes.c_void_p))
        return rho[: n_ok.value].copy(), status.astype(bool)

    def bubble_point(self, temperature, liquid_molefracs, pressure):
        return self._point(temperature, liquid_molefracs, pressure, True)

    def dew_point(self, temperature, vapor_molefracs, pressure):
        return self._point(temperature, vapor_molefracs, pressure, False)
