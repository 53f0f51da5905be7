"""``PcSaftMix`` -- same public API as the reference (`/root/reference/feos_torch/pcsaft_mix.py:12-479`)
for binary mixtures with k_ij, computed by the sm_100a kernels behind ``include/feos_b200.h``.

``bubble_point`` / ``dew_point`` return ``(pressure [n_ok] in Pa, failed mask [N])`` -- value first,
unlike PcSaftPure (pcsaft_mix.py:444, :468).  Gradients flow to ``parameters`` [N,2,8], ``kij`` [N,2]
and ``temperature``; mole fractions and the starting pressure carry none (pcsaft_mix.py:427).
"""
import torch

from . import _lib
from .pcsaft_pure import _back, _cuda_device, _prep, _stream, compact


class _MixPoint(torch.autograd.Function):
    @staticmethod
    def forward(ctx, params, kij, temperature, molefracs, bubble, order):
        n = temperature.shape[0]
        dev = params.device
        rho = torch.empty((n, 4), dtype=torch.float64, device=dev)
        val = torch.empty((n,), dtype=torch.float64, device=dev)
        failed = torch.empty((n,), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            _lib.timed_call(
                "feos_dev_mix_point_fwd", _lib.ptr(params), _lib.ptr(kij), _lib.ptr(temperature), _lib.ptr(molefracs), n,
                int(bubble), _lib.ptr(rho), _lib.ptr(val), _lib.ptr(failed), None, _lib.ptr(order), _stream(),
            )
            out, pos = compact(failed, val)
        # the backward kernel runs no iteration: the order only pays where code paths are mixed, and costs
        # scattered gradient writes otherwise -- it keeps the input order
        ctx.save_for_backward(params, kij, temperature, rho, failed, pos, None)
        ctx.bubble = int(bubble)
        ctx.mark_non_differentiable(rho, failed)
        return out, failed, rho

    @staticmethod
    def backward(ctx, g_val, _g_failed, _g_rho):
        params, kij, temperature, rho, failed, pos, order = ctx.saved_tensors
        n = temperature.shape[0]
        g = g_val.contiguous()
        need_p, need_k, need_t = ctx.needs_input_grad[:3]
        gp = torch.empty_like(params) if need_p else None
        gk = torch.empty_like(kij) if need_k else None
        gt = torch.empty_like(temperature) if need_t else None
        with torch.cuda.device(params.device):
            _lib.timed_call(
                "feos_dev_mix_point_bwd", _lib.ptr(params), _lib.ptr(kij), _lib.ptr(temperature), _lib.ptr(rho),
                _lib.ptr(failed), _lib.ptr(pos), _lib.ptr(g), n, ctx.bubble, _lib.ptr(gp), _lib.ptr(gk), _lib.ptr(gt),
                _lib.ptr(order), _stream(),
            )
        return gp, gk, gt, None, None, None


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
        with torch.cuda.device(dev):
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
            with torch.cuda.device(params.device):
                _lib.call("feos_dev_mix_state_bwd", _lib.ptr(params), _lib.ptr(kij), _lib.ptr(temperature), _lib.ptr(density), n,
                          _lib.ptr(ga), _lib.ptr(gp), _lib.ptr(gmu), _lib.ptr(gv), _lib.ptr(out[0]), _lib.ptr(out[1]),
                          _lib.ptr(out[2]), _lib.ptr(out[3]), _stream())
        return tuple(out)


class PcSaftMix:
    """parameters [N, 2, 8] (rows as in PcSaftPure), kij [N, 2] = (k_ij, epsilon_k_aibj; 0 = combining rule)."""

    def __init__(self, parameters, kij=None):
        if parameters.dim() != 3 or parameters.shape[1] != 2 or parameters.shape[2] != 8:
            raise ValueError("parameters must have shape [N, 2, 8] (binary mixtures)")
        _lib.load()
        self._out_device = parameters.device
        self._dev = _cuda_device(parameters)
        self._params = _prep(parameters, self._dev)
        if kij is None:
            # the reference accepts kij=None in the signature but needs it (pcsaft_mix.py:141); zero
            # k_ij / combining-rule eps_AiBj is what "no kij" means physically
            kij = torch.zeros((parameters.shape[0], 2), dtype=torch.float64, device=parameters.device)
        if kij.shape != (parameters.shape[0], 2):
            raise ValueError("kij must have shape [N, 2]")
        self._kij = _prep(kij, self._dev)
        self._check_association(parameters)
        self._source = (parameters, kij)

    @staticmethod
    def _check_association(parameters):
        na, nb = parameters[:, :, 6].detach(), parameters[:, :, 7].detach()
        # pcsaft_mix.py:120-121 (can never trigger for two components; kept for API parity)
        if bool((torch.count_nonzero(na + nb, dim=1) > 2).any()):
            raise Exception("Only up to two associating components are allowed!")

    @property
    def parameters(self):
        return self._source[0].detach().cpu().numpy()

    @property
    def kij_np(self):
        return self._source[1].detach().cpu().numpy()

    def _back(self, t):
        return _back(t, self._out_device)

    def _point(self, temperature, molefracs, bubble):
        t = _prep(temperature, self._dev)
        z = _prep(molefracs.detach(), self._dev)
        order = None
        n = t.shape[0]
        if _lib.state_order_enabled(n):
            # states with the same code path and neighbouring (T, x) share a warp (csrc/order.cuh); the data set
            # is recognised by the caller's (T, x) tensor objects in the next iteration of a fit loop
            def compute():
                o = torch.empty((n,), dtype=torch.int32, device=self._dev)
                with torch.cuda.device(self._dev):
                    _lib.timed_call("feos_dev_mix_point_order", _lib.ptr(self._params), _lib.ptr(t), _lib.ptr(z), n,
                                    _lib.ptr(o), _stream())
                return o
            order = _lib.cached_order(("mix", self._dev.index), (temperature, molefracs), compute)
        val, failed, _ = _MixPoint.apply(self._params, self._kij, t, z, bubble, order)
        return self._back(val), self._back(failed.bool())

    def bubble_point(self, temperature, liquid_molefracs, pressure):
        """(p_bubble [n_ok] in Pa, failed mask [N]) -- pcsaft_mix.py:422-444.  `pressure` is the
        reference's starting value for its CPU solver; the kernels build their own starting point."""
        return self._point(temperature, liquid_molefracs, True)

    def dew_point(self, temperature, vapor_molefracs, pressure):
        """(p_dew [n_ok] in Pa, failed mask [N]) -- pcsaft_mix.py:446-468."""
        return self._point(temperature, vapor_molefracs, False)

    def derivatives(self, temperature, density):
        """(a [N], p/kT [N], mu_i^res/kT [N,2], v_i [N,2]) at (T, rho [N,2]) -- pcsaft_mix.py:395-420; carries
        gradients w.r.t. parameters, kij, T and rho like the reference's hyper-dual evaluation under autograd."""
        t = _prep(temperature, self._dev)
        r = _prep(density, self._dev)
        a, p, mu, v = _MixState.apply(self._params, self._kij, t, r)
        return self._back(a), self._back(p), self._back(mu), self._back(v)

    def helmholtz_energy_density(self, temperature, density):
        """a = A/(V kT) [N,1] -- pcsaft_mix.py:31-154."""
        return self.derivatives(temperature, density)[0][:, None]
