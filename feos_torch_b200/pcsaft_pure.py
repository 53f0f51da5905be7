"""``PcSaftPure`` -- same public API as the reference (`/root/reference/feos_torch/pcsaft_pure.py:89-243`),
computed by the sm_100a kernels behind ``include/feos_b200.h``.

The reference runs a CPU solve (Rust/feos) and then re-evaluates the Helmholtz energy in PyTorch to
obtain gradients.  Here one forward kernel solves and evaluates the property, a three-launch
compaction removes the rows of failed states (the reference's return convention), and one backward
kernel applies the implicit-function-theorem gradient; all wrapped in ``torch.autograd.Function``.

Differences that are not observable through returned values: the model object is not mutated
(`reduce`, pcsaft_pure.py:235-243), and tensors may live on the CPU -- they are moved to the
current CUDA device for the computation and the results are returned on the parameters' device.
"""
import torch

from . import _lib


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _cuda_device(t):
    return t.device if t.is_cuda else torch.device("cuda", torch.cuda.current_device())


class _HostToDevice(torch.autograd.Function):
    """Host tensor -> device for the computation; the gradient comes back into PINNED host memory
    (torch's caching host allocator), so the device->host copy runs at PCIe speed instead of paging
    in a fresh pageable buffer every step."""

    @staticmethod
    def forward(ctx, t, dev):
        return t.to(device=dev, dtype=torch.float64, non_blocking=True).contiguous()

    @staticmethod
    def backward(ctx, g):
        return _to_host(g), None


def _to_host(t):
    out = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
    out.copy_(t, non_blocking=True)
    torch.cuda.current_stream(t.device).synchronize()
    return out


class _DeviceToHost(torch.autograd.Function):
    """Result -> pinned host memory; the incoming host gradient goes back to the device."""

    @staticmethod
    def forward(ctx, t):
        ctx.dev = t.device
        return _to_host(t)

    @staticmethod
    def backward(ctx, g):
        return g.to(ctx.dev, non_blocking=True)


def _prep(t, dev):
    if not t.is_cuda and t.requires_grad:
        return _HostToDevice.apply(t, dev)
    return t.to(device=dev, dtype=torch.float64).contiguous()


def _back(t, out_device):
    """Return a result on the device the caller's tensors live on (the reference's `.to(self.m.device)`)."""
    if out_device.type == "cpu":
        return _DeviceToHost.apply(t) if t.dtype == torch.float64 else t.cpu()
    return t.to(out_device)


def compact(failed, values, out_cols=None):
    """Rows of ``values`` [n] or [n, c] whose state converged, in input order, plus the row map
    ``pos`` [n] (int32, -1 = failed).  One host synchronisation (the output length)."""
    n = failed.shape[0]
    dev = failed.device
    in_cols = 1 if values.dim() == 1 else values.shape[1]
    out_cols = in_cols if out_cols is None else out_cols
    pos = torch.empty((n,), dtype=torch.int32, device=dev)
    n_ok = torch.empty((1,), dtype=torch.int64, device=dev)
    out = torch.empty((n, out_cols) if values.dim() == 2 else (n,), dtype=torch.float64, device=dev)
    _lib.call("feos_dev_compact", _lib.ptr(failed), n, _lib.ptr(pos), _lib.ptr(n_ok), _lib.ptr(values), in_cols,
              _lib.ptr(out), out_cols, _stream())
    k = int(n_ok.item())
    return out[:k], pos


class _PureVle(torch.autograd.Function):
    """which = 0: vapor pressure [Pa]; which = 1: saturated liquid density [kmol/m^3].
    Returns (compacted values [n_ok], failed [n] uint8, rho [n, 2])."""

    @staticmethod
    def forward(ctx, params, temperature, which):
        n = temperature.shape[0]
        dev = params.device
        rho = torch.empty((n, 2), dtype=torch.float64, device=dev)
        val = torch.empty((n,), dtype=torch.float64, device=dev)
        failed = torch.empty((n,), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            _lib.timed_call(
                "feos_dev_pure_vle_fwd", _lib.ptr(params), _lib.ptr(temperature), n, _lib.ptr(rho),
                _lib.ptr(val) if which == 0 else None, _lib.ptr(val) if which == 1 else None,
                _lib.ptr(failed), None, _stream(),
            )
            out, pos = compact(failed, val)
        ctx.save_for_backward(params, temperature, rho, failed, pos)
        ctx.which = which
        ctx.mark_non_differentiable(rho, failed)
        return out, failed, rho

    @staticmethod
    def backward(ctx, g_val, _g_failed, _g_rho):
        params, temperature, rho, failed, pos = ctx.saved_tensors
        n = temperature.shape[0]
        g = g_val.contiguous()
        grad_params = torch.empty_like(params)
        grad_t = torch.empty_like(temperature) if ctx.needs_input_grad[1] else None
        with torch.cuda.device(params.device):
            _lib.timed_call(
                "feos_dev_pure_vle_bwd", _lib.ptr(params), _lib.ptr(temperature), _lib.ptr(rho), _lib.ptr(failed),
                _lib.ptr(pos), _lib.ptr(g) if ctx.which == 0 else None, _lib.ptr(g) if ctx.which == 1 else None, n,
                _lib.ptr(grad_params), _lib.ptr(grad_t), _stream(),
            )
        return grad_params, grad_t, None


class _PureRhoTp(torch.autograd.Function):
    @staticmethod
    def forward(ctx, params, temperature, pressure):
        n = temperature.shape[0]
        dev = params.device
        rho = torch.empty((n,), dtype=torch.float64, device=dev)
        val = torch.empty((n,), dtype=torch.float64, device=dev)
        failed = torch.empty((n,), dtype=torch.uint8, device=dev)
        with torch.cuda.device(dev):
            _lib.timed_call(
                "feos_dev_pure_rho_tp_fwd", _lib.ptr(params), _lib.ptr(temperature), _lib.ptr(pressure), n,
                _lib.ptr(rho), _lib.ptr(val), _lib.ptr(failed), None, _stream(),
            )
            out, pos = compact(failed, val)
        ctx.save_for_backward(params, temperature, pressure, rho, failed, pos)
        ctx.mark_non_differentiable(rho, failed)
        return out, failed, rho

    @staticmethod
    def backward(ctx, g_val, _g_failed, _g_rho):
        params, temperature, pressure, rho, failed, pos = ctx.saved_tensors
        n = temperature.shape[0]
        g = g_val.contiguous()
        need_p, need_t, need_pr = ctx.needs_input_grad
        grad_params = torch.empty_like(params) if need_p else None
        grad_t = torch.empty_like(temperature) if need_t else None
        grad_pr = torch.empty_like(pressure) if need_pr else None
        with torch.cuda.device(params.device):
            _lib.timed_call(
                "feos_dev_pure_rho_tp_bwd", _lib.ptr(params), _lib.ptr(temperature), _lib.ptr(pressure),
                _lib.ptr(rho), _lib.ptr(failed), _lib.ptr(pos), _lib.ptr(g), n, _lib.ptr(grad_params),
                _lib.ptr(grad_t), _lib.ptr(grad_pr), _stream(),
            )
        return grad_params, grad_t, grad_pr


class PcSaftPure:
    """parameters [N, 8] = m, sigma, epsilon_k, mu, kappa_ab, epsilon_k_ab, na, nb (README.md:12)."""

    def __init__(self, parameters):
        if parameters.dim() != 2 or parameters.shape[1] != 8:
            raise ValueError("parameters must have shape [N, 8]")
        _lib.load()  # fail loudly before any work if the CUDA extension is missing
        self._out_device = parameters.device
        self._dev = _cuda_device(parameters)
        self._params = _prep(parameters, self._dev)
        self._source = parameters

    @property
    def parameters(self):
        """numpy copy of the parameters (pcsaft_pure.py:104), made on demand: the kernels read the
        device tensor, so no device->host copy happens on the hot path."""
        return self._source.detach().cpu().numpy()

    def _back(self, t):
        return _back(t, self._out_device)

    # ---- properties on the phase-equilibrium hot path ------------------------------------------
    def vapor_pressure(self, temperature):
        """(failed mask [N], p_sat [n_ok] in Pa) -- pcsaft_pure.py:201-215."""
        t = _prep(temperature, self._dev)
        val, failed, _ = _PureVle.apply(self._params, t, 0)
        return self._back(failed.bool()), self._back(val)

    def equilibrium_liquid_density(self, temperature):
        """(failed mask [N], rho_L^sat [n_ok] in kmol/m^3) -- pcsaft_pure.py:217-233."""
        t = _prep(temperature, self._dev)
        val, failed, _ = _PureVle.apply(self._params, t, 1)
        return self._back(failed.bool()), self._back(val)

    def liquid_density(self, temperature, pressure):
        """(failed mask [N], rho_L(T, p) [n_ok] in kmol/m^3) -- pcsaft_pure.py:184-199."""
        t = _prep(temperature, self._dev)
        p = _prep(pressure, self._dev)
        val, failed, _ = _PureRhoTp.apply(self._params, t, p)
        return self._back(failed.bool()), self._back(val)

    # ---- state functions (values only for now) -----------------------
    def derivatives(self, temperature, density):
        """(a, p/kT, d(p/kT)/drho) at (T, rho [1/A^3]) -- pcsaft_pure.py:180-182."""
        t = _prep(temperature, self._dev)
        r = _prep(density, self._dev)
        n = t.shape[0]
        a, p, dp = (torch.empty((n,), dtype=torch.float64, device=self._dev) for _ in range(3))
        with torch.cuda.device(self._dev):
            _lib.call("feos_dev_pure_state", _lib.ptr(self._params.detach()), _lib.ptr(t.detach()), _lib.ptr(r.detach()),
                      n, _lib.ptr(a), _lib.ptr(p), _lib.ptr(dp), _stream())
        return self._back(a), self._back(p), self._back(dp)

    def helmholtz_energy(self, temperature, density):
        """a = A/(V kT) in 1/A^3 -- pcsaft_pure.py:106-178."""
        return self.derivatives(temperature, density)[0]
