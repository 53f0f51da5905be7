"""``PcSaftMix`` -- same public API as the reference (`/root/reference/feos_torch/pcsaft_mix.py:12-479`)
for binary mixtures with k_ij, computed by the sm_100a kernels behind ``include/feos_b200.h``.

``bubble_point`` / ``dew_point`` return ``(pressure [n_ok] in Pa, failed mask [N])`` -- value first,
unlike PcSaftPure (pcsaft_mix.py:444, :468).  Gradients flow to ``parameters`` [N,2,8], ``kij`` [N,2]
and ``temperature``; mole fractions and the starting pressure carry none (pcsaft_mix.py:427).
"""
import torch

from . import _lib
from .pcsaft_pure import _back, _on_device, _check_rows, _cuda_device, _prep, _stream, _upstream, compact


SHARED_PARAMS, SHARED_KIJ, PLAIN_ONLY = 1, 2, 4  # include/feos_b200.h: FEOS_SHARED_PARAMS, FEOS_SHARED_KIJ, FEOS_PLAIN_ONLY


def _shared_row(t, inner):
    """If ``t`` [N, *inner] is ONE row expanded along the batch axis (``row.expand(N, *inner)`` -- how a fit of shared
    parameters is written in the reference's API, BASELINE config C4), return that row as a tensor of shape ``inner``
    that is part of the autograd graph of ``t`` (its view base), else None.  The kernels then read the one row and the
    backward kernel reduces its gradient over the states itself, instead of N materialised rows and an [N, *inner]
    gradient summed afterwards by torch."""
    base = t._base
    if (base is None or t.dim() != len(inner) + 1 or t.shape[0] < 2 or t.stride(0) != 0
            or tuple(t.shape[1:]) != tuple(inner) or t.dtype != torch.float64 or base.dtype != torch.float64):
        return None
    if tuple(base.shape) != tuple(inner) or not base.is_contiguous() or base.data_ptr() != t.data_ptr():
        return None
    want = tuple(base.stride())
    if tuple(t.stride()[1:]) != want:
        return None
    return base


_CLASS_CACHE = []  # (weakref to the parameter tensor, version, shape, all pairs plain)


class _MixPoint(torch.autograd.Function):
    @staticmethod
    def forward(ctx, params, kij, temperature, molefracs, pressure, bubble, order, flags):
        n = temperature.shape[0]
        dev = params.device
        rho = torch.empty((n, 4), dtype=torch.float64, device=dev)
        val = torch.empty((n,), dtype=torch.float64, device=dev)
        failed = torch.empty((n,), dtype=torch.uint8, device=dev)
        with _on_device(dev):
            _lib.timed_call(
                "feos_dev_mix_point_fwd", _lib.ptr(params), _lib.ptr(kij), _lib.ptr(temperature), _lib.ptr(molefracs),
                _lib.ptr(pressure), n, int(bubble), flags, _lib.ptr(rho), _lib.ptr(val), _lib.ptr(failed), None,
                _lib.ptr(order), _stream(),
            )
            out, pos = compact(failed, val)
        # the backward kernel runs no iteration: the order only pays where code paths are mixed, and costs
        # scattered gradient writes otherwise -- it keeps the input order
        ctx.save_for_backward(params, kij, temperature, rho, failed, pos, None)
        ctx.bubble, ctx.flags = int(bubble), flags
        ctx.mark_non_differentiable(rho, failed)
        ctx.set_materialize_grads(False)  # no zero-filled gradients for the mask / density outputs
        return out, failed, rho

    @staticmethod
    def backward(ctx, g_val, _g_failed, _g_rho):
        params, kij, temperature, rho, failed, pos, order = ctx.saved_tensors
        n = temperature.shape[0]
        if g_val is None:  # only the mask / density outputs were used downstream
            return (None,) * 8
        g, bc = _upstream(g_val)
        need_p, need_k, need_t = ctx.needs_input_grad[:3]
        # shared rows: the kernel accumulates the sum over the states (zero first); per-state rows are written
        gp = (torch.zeros_like(params) if ctx.flags & SHARED_PARAMS else torch.empty_like(params)) if need_p else None
        gk = (torch.zeros_like(kij) if ctx.flags & SHARED_KIJ else torch.empty_like(kij)) if need_k else None
        gt = torch.empty_like(temperature) if need_t else None
        with _on_device(params.device):
            _lib.timed_call(
                "feos_dev_mix_point_bwd", _lib.ptr(params), _lib.ptr(kij), _lib.ptr(temperature), _lib.ptr(rho),
                _lib.ptr(failed), _lib.ptr(pos), _lib.ptr(g), bc, n, ctx.bubble, ctx.flags, _lib.ptr(gp), _lib.ptr(gk),
                _lib.ptr(gt), _lib.ptr(order), _stream(),
            )
        return gp, gk, gt, None, None, None, None, None


class _MixState(torch.autograd.Function):
    """(a, p, mu, v) = derivatives(T, rho [N,2]); backward = dual passes of csrc/state_grad.cuh."""

    @staticmethod
    def forward(ctx, params, kij, temperature, density):
        n = temperature.shape[0]
        dev = params.device
        a = torch.empty((n,), dtype=torch.float64, device=dev)
        p = torch.empty((n,), dtype=torch.float64, device=dev)
        mu = torch.empty((n, 2), dtype=torch.float64, device=dev)
        v = torch.empty((n, 2), dtype=torch.float64, device=dev)
        with _on_device(dev):
            _lib.call("feos_dev_mix_state", _lib.ptr(params), _lib.ptr(kij), _lib.ptr(temperature), _lib.ptr(density), n,
                      _lib.ptr(a), _lib.ptr(p), _lib.ptr(mu), _lib.ptr(v), _stream())
        ctx.save_for_backward(params, kij, temperature, density)
        return a, p, mu, v

    @staticmethod
    def backward(ctx, ga, gp, gmu, gv):
        params, kij, temperature, density = ctx.saved_tensors
        n = temperature.shape[0]
        ga, gp, gmu, gv = (None if g is None else g.contiguous() for g in (ga, gp, gmu, gv))
        need = ctx.needs_input_grad
        out = [torch.empty_like(x) if nd else None for x, nd in zip((params, kij, temperature, density), need)]
        if n:
            with _on_device(params.device):
                _lib.call("feos_dev_mix_state_bwd", _lib.ptr(params), _lib.ptr(kij), _lib.ptr(temperature), _lib.ptr(density), n,
                          _lib.ptr(ga), _lib.ptr(gp), _lib.ptr(gmu), _lib.ptr(gv), _lib.ptr(out[0]), _lib.ptr(out[1]),
                          _lib.ptr(out[2]), _lib.ptr(out[3]), _stream())
        return tuple(out)


class PcSaftMix:
    """parameters [N, 2, 8] (rows as in PcSaftPure), kij [N, 2] = (k_ij, epsilon_k_aibj; 0 = combining rule).

    Binary mixtures only, like the reference: its dispersion loop is written for n components (pcsaft_mix.py:73-90),
    but no call with n != 2 gets through its public API -- with `kij` it raises "kij can only be used for binary
    mixtures!" (pcsaft_mix.py:74-75), without it `self.kij[cross_associating, 1]` fails on None (pcsaft_mix.py:141),
    and the native layer is binary (src/pcsaft.rs:173, :206).  tests/test_reference_limits.py pins this.

    Shared parameters: `row.expand(N, 2, 8)` / `kij_row.expand(N, 2)` are recognised (stride 0 along the batch) and
    handled as ONE row -- nothing is materialised per state and the gradient sum over the states is reduced inside
    the backward kernel."""

    def __init__(self, parameters, kij=None):
        if parameters.dim() != 3 or parameters.shape[2] != 8:
            raise ValueError("parameters must have shape [N, 2, 8]")
        if parameters.shape[1] != 2:
            # the reference raises for every n != 2 as well (pcsaft_mix.py:74-75 with kij, :141 without)
            raise NotImplementedError("binary mixtures only ([N, 2, 8]); the reference's API cannot evaluate n != 2 either")
        _lib.load()
        self._n = parameters.shape[0]
        self._nc = parameters.shape[1]
        self._out_device = parameters.device
        self._dev = _cuda_device(parameters)
        if kij is None:
            # the reference accepts kij=None in the signature but needs it (pcsaft_mix.py:141); zero
            # k_ij / combining-rule eps_AiBj is what "no kij" means physically
            kij = torch.zeros((parameters.shape[0], 2), dtype=torch.float64, device=parameters.device)
        if kij.shape != (parameters.shape[0], 2):
            raise ValueError("kij must have shape [N, 2]")
        self._source = (parameters, kij)
        self._flags = 0
        self._params_row = _shared_row(parameters, (2, 8)) if self._nc == 2 else None
        self._kij_row = _shared_row(kij, (2,)) if self._nc == 2 else None
        self._params_full = self._kij_full = None
        if self._params_row is not None:
            self._flags |= SHARED_PARAMS
            self._params_row = _prep(self._params_row, self._dev)
        else:
            self._params_full = _prep(parameters, self._dev)
        if self._kij_row is not None:
            self._flags |= SHARED_KIJ
            self._kij_row = _prep(self._kij_row, self._dev)
        else:
            self._kij_full = _prep(kij, self._dev)
        if self._check_association(parameters if self._params_row is None else self._params_row[None]):
            self._flags |= PLAIN_ONLY

    # per-state [N, n, 8] / [N, 2] device tensors for the consumers that want them (state functions)
    @property
    def _params(self):
        if self._params_full is None:
            self._params_full = _prep(self._source[0], self._dev)
        return self._params_full

    @property
    def _kij(self):
        if self._kij_full is None:
            self._kij_full = _prep(self._source[1], self._dev)
        return self._kij_full

    def _check_association(self, parameters):
        """The reference's constructor check (pcsaft_mix.py:120-121) and, from the same look at the parameters, whether
        every pair is plain (no dipole, no association sites: the kernels then skip the launch of the instantiation for
        such pairs, FEOS_PLAIN_ONLY).  The answer needs one device->host read, so it is kept per parameter tensor
        object and version: a fit that builds a model per step from the same (unchanged) parameter pair pays it once."""
        import weakref

        key = parameters if parameters._base is None else parameters._base
        for ref, version, shape, plain in _CLASS_CACHE:
            if ref() is key and version == key._version and shape == tuple(parameters.shape):
                return plain
        p = parameters.detach()
        sites = p[:, :, 6] + p[:, :, 7]
        too_many = (torch.count_nonzero(sites, dim=1) > 2).any()
        not_plain = (p[:, :, 3] != 0).any() | (sites != 0).any()
        too_many, not_plain = torch.stack([too_many, not_plain]).tolist()
        if too_many:
            raise Exception("Only up to two associating components are allowed!")
        _CLASS_CACHE[:] = [e for e in _CLASS_CACHE if e[0]() is not None][-7:]
        _CLASS_CACHE.append((weakref.ref(key), key._version, tuple(parameters.shape), not not_plain))
        return not not_plain

    @property
    def parameters(self):
        return self._source[0].detach().cpu().numpy()

    @property
    def kij_np(self):
        return self._source[1].detach().cpu().numpy()

    def _back(self, t):
        return _back(t, self._out_device)

    def _point(self, temperature, molefracs, pressure, bubble):
        _check_rows(self._n, temperature=temperature, molefracs=molefracs, pressure=pressure)
        t = _prep(temperature, self._dev)
        z = _prep(molefracs.detach(), self._dev)
        p0 = None if pressure is None else _prep(pressure.detach(), self._dev)
        params = self._params_row if self._params_row is not None else self._params_full
        kij = self._kij_row if self._kij_row is not None else self._kij_full
        order = None
        n = t.shape[0]
        if _lib.state_order_enabled(n):
            # states with the same code path and neighbouring (T, x) share a warp (csrc/order.cuh); the data set
            # is recognised by the caller's (T, x) tensor objects in the next iteration of a fit loop
            def compute(out):
                o = out if out is not None else torch.empty((n,), dtype=torch.int32, device=self._dev)
                with _on_device(self._dev):
                    _lib.timed_call("feos_dev_mix_point_order", _lib.ptr(params), _lib.ptr(t), _lib.ptr(z), n,
                                    self._flags & SHARED_PARAMS, _lib.ptr(o), _stream())
                return o
            order = _lib.cached_order(("mix", self._dev.index), (temperature, molefracs), compute, n=n)
        val, failed, _ = _MixPoint.apply(params, kij, t, z, p0, bubble, order, self._flags)
        return self._back(val), self._back(failed.view(torch.bool))

    def bubble_point(self, temperature, liquid_molefracs, pressure):
        """(p_bubble [n_ok] in Pa, failed mask [N]) -- pcsaft_mix.py:422-444.  `pressure` is the starting value the
        reference hands to its CPU solver (src/pcsaft.rs:170-177); here the fast path builds its own start and the
        retry path starts its successive substitution from it where no start of its own exists (near-critical)."""
        return self._point(temperature, liquid_molefracs, pressure, True)

    def dew_point(self, temperature, vapor_molefracs, pressure):
        """(p_dew [n_ok] in Pa, failed mask [N]) -- pcsaft_mix.py:446-468.  `pressure`: see bubble_point."""
        return self._point(temperature, vapor_molefracs, pressure, False)

    def derivatives(self, temperature, density):
        """(a [N], p/kT [N], mu_i^res/kT [N,2], v_i [N,2]) at (T, rho [N,2]) -- pcsaft_mix.py:395-420; carries
        gradients w.r.t. parameters, kij, T and rho like the reference's hyper-dual evaluation under autograd."""
        _check_rows(self._n, temperature=temperature, density=density)
        if density.dim() != 2 or density.shape[1] != self._nc:
            raise ValueError(f"density must have shape [N, {self._nc}]")
        t = _prep(temperature, self._dev)
        r = _prep(density, self._dev)
        a, p, mu, v = _MixState.apply(self._params, self._kij, t, r)
        return self._back(a), self._back(p), self._back(mu), self._back(v)

    def helmholtz_energy_density(self, temperature, density):
        """a = A/(V kT) [N,1] -- pcsaft_mix.py:31-154."""
        return self.derivatives(temperature, density)[0][:, None]
