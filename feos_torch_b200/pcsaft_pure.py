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
import weakref

import torch

from . import _lib


def _stream():
    """Raw handle of the current stream of the current device (torch._C call: the Python wrapper objects of
    torch.cuda.current_stream cost several microseconds per launch)."""
    return torch._C._cuda_getCurrentRawStream(torch.cuda.current_device())


class _on_device:
    """`with _on_device(dev):` = torch.cuda.device(dev) that costs nothing when `dev` is already current (the usual
    case: one process per GPU)."""
    __slots__ = ("ctx",)

    def __init__(self, dev):
        self.ctx = None if dev.index is None or dev.index == torch.cuda.current_device() else torch.cuda.device(dev)

    def __enter__(self):
        if self.ctx is not None:
            self.ctx.__enter__()

    def __exit__(self, *exc):
        if self.ctx is not None:
            return self.ctx.__exit__(*exc)
        return False


def _cuda_device(t):
    return t.device if t.is_cuda else torch.device("cuda", torch.cuda.current_device())


# ---- host tensors in / out: chunked, stream-pipelined transfers --------------------------------
# A [N, 8] parameter tensor on the host is 64 B per state each way (parameters in, gradients out): at
# PCIe speed that is as long as the solves themselves.  Large pinned inputs are therefore copied in row
# chunks on a copy stream with one event per chunk; the forward kernels are launched per chunk behind
# its event, and the backward kernels record one event per chunk so that the gradient's device->host
# copy of chunk k runs while chunk k + 1 is being computed.
_CHUNK_MIN = 1 << 18  # rows; below 2 chunks of this size a single copy is used
_COPY_STREAMS = {}


class _Chunks:
    """Row ranges of a device tensor and the events after which each range is valid."""
    __slots__ = ("bounds", "events", "version")

    def __init__(self, bounds, events, version):
        self.bounds, self.events, self.version = bounds, events, version


def _chunk_bounds(n):
    k = max(1, min(16, n // _CHUNK_MIN))
    step = (-(-n // k) + 127) // 128 * 128
    return [(a, min(a + step, n)) for a in range(0, n, step)]


def _copy_stream(dev, kind):
    key = (dev.index, kind)
    if key not in _COPY_STREAMS:
        _COPY_STREAMS[key] = torch.cuda.Stream(device=dev)
    return _COPY_STREAMS[key]


def _chunks_of(t):
    """The pending-chunk record of a tensor, if it is still the tensor the record was made for."""
    ch = getattr(t, "_feos_chunks", None)
    if ch is not None and ch.version != t._version:  # modified in place since (e.g. gradient accumulation)
        return None
    return ch


def _wait_chunks(t):
    """Make the current stream wait for every pending chunk of ``t`` (consumers that read it whole)."""
    ch = _chunks_of(t)
    if ch is not None:
        cur = torch.cuda.current_stream(t.device)
        for ev in ch.events:
            cur.wait_event(ev)
    return t


class _HostToDevice(torch.autograd.Function):
    """Host tensor -> device for the computation; the gradient comes back into PINNED host memory
    (torch's caching host allocator), so the device->host copy runs at PCIe speed instead of paging
    in a fresh pageable buffer every step."""
    last_chunks = None

    @staticmethod
    def forward(ctx, t, dev, chunked):
        n = t.shape[0] if t.dim() else 0
        if chunked and n >= 2 * _CHUNK_MIN and t.dtype == torch.float64 and t.is_contiguous() and t.is_pinned():
            out = torch.empty(t.shape, dtype=torch.float64, device=dev)
            cur = torch.cuda.current_stream(dev)
            cs = _copy_stream(dev, "h2d")
            cs.wait_stream(cur)  # the allocation may recycle memory still in use on the current stream
            bounds, events = _chunk_bounds(n), []
            with torch.cuda.stream(cs):
                for a, b in bounds:
                    out[a:b].copy_(t[a:b], non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(cs)
                    events.append(ev)
            out.record_stream(cs)
            _HostToDevice.last_chunks = _Chunks(bounds, events, out._version)
            return out
        _HostToDevice.last_chunks = None
        return t.to(device=dev, dtype=torch.float64, non_blocking=True).contiguous()

    @staticmethod
    def backward(ctx, g):
        return _to_host(g, _chunks_of(g)), None, None


# Result buffers in pinned host memory.  torch's caching host allocator hands the most recently freed block back
# first, i.e. the block whose lines the caller's CPU threads have just read (e.g. the loss reduction); a device->host
# copy into such a block ran at 10-15 GB/s instead of 50 GB/s on the development boxes.  Holding the last few result
# buffers a little longer makes the allocator rotate through cold blocks instead (FEOS_B200_PINNED_RING, 0 = off).
import collections as _collections
import os as _os

_PINNED_RING = _collections.deque(maxlen=max(0, int(_os.environ.get("FEOS_B200_PINNED_RING", "6"))) or None)
_PINNED_RING_ON = _PINNED_RING.maxlen is not None


def _to_host(t, chunks=None):
    out = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
    if _PINNED_RING_ON and out.numel() * out.element_size() <= (32 << 20):
        _PINNED_RING.append(out)
    if chunks is not None:
        ds = _copy_stream(t.device, "d2h")
        with torch.cuda.stream(ds):
            for (a, b), ev in zip(chunks.bounds, chunks.events):
                ds.wait_event(ev)
                out[a:b].copy_(t[a:b], non_blocking=True)
        t.record_stream(ds)
        ds.synchronize()
        return out
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
        if g.numel() > 1 and all(st == 0 for st in g.stride()):
            # gradient of a sum: one value expanded -- move the value, expand on the device
            return g.as_strided((1,), (1,), g.storage_offset()).to(ctx.dev, non_blocking=True).expand(g.shape)
        return g.to(ctx.dev, non_blocking=True)


def _prep(t, dev, chunked=False):
    if not t.is_cuda and t.requires_grad:
        out = _HostToDevice.apply(t, dev, chunked)
        if _HostToDevice.last_chunks is not None:
            out._feos_chunks = _HostToDevice.last_chunks
            _HostToDevice.last_chunks = None
        return out
    # (asynchronous for pinned sources; every property call synchronises before it returns)
    return t.to(device=dev, dtype=torch.float64, non_blocking=True).contiguous()


def _off(t, rows):
    """Pointer to row ``rows`` of a contiguous tensor (chunked launches through the C ABI)."""
    if t is None:
        return None
    if rows == 0:
        return t.data_ptr()
    stride = t.stride(0) if t.dim() > 1 else 1
    return t.data_ptr() + rows * stride * t.element_size()


def _launch_chunked(params, ranges_events, launch):
    """``launch(a, b)`` per pending chunk of ``params`` behind its event, or once over all rows."""
    cur = torch.cuda.current_stream(params.device)
    for (a, b), ev in ranges_events:
        if ev is not None:
            cur.wait_event(ev)
        launch(a, b)


def _record_chunk_events(bounds, launch, dev):
    """Run ``launch(a, b)`` per range on the current stream, recording one event after each."""
    cur = torch.cuda.current_stream(dev)
    events = []
    for a, b in bounds:
        launch(a, b)
        ev = torch.cuda.Event()
        ev.record(cur)
        events.append(ev)
    return events


def _back(t, out_device):
    """Return a result on the device the caller's tensors live on (the reference's `.to(self.m.device)`)."""
    if out_device.type == "cpu":
        # masks too go through pinned memory: a pageable device->host copy runs at a fraction of PCIe speed
        return _DeviceToHost.apply(t) if t.dtype == torch.float64 else _to_host(t)
    return t.to(out_device)


def _check_rows(n, **tensors):
    """Every per-state tensor must have one row per state of the model.  The reference's native layer panics on any
    length mismatch (ndarray Zip, src/pcsaft.rs:86-88, :157-162); the C ABI only sees raw pointers, so the check lives here.
    Unlike the reference (`reduce`, pcsaft_pure.py:235-243, pcsaft_mix.py:470-479) the model object is NOT reduced after a call with failed
    states: follow-up calls take full-length [N] tensors again."""
    for name, t in tensors.items():
        if t is None:
            continue
        if t.dim() < 1 or t.shape[0] != n:
            raise ValueError(
                f"{name} must have one row per state of the model ({n}), got shape {tuple(t.shape)}. "
                "This model is not reduced after a call with failed states (the reference mutates it, `reduce`): "
                "pass full-length tensors, or build a new model from the surviving parameter rows.")


# The number of converged states is the one value the host needs before a property call can return (it is the
# length of the returned tensor).  The compaction's scan kernel writes it straight into pinned host memory and the
# host polls that word: no device->host memcpy, no stream synchronisation, and the scatter kernel that follows runs
# while Python is already preparing the next call.
_COUNT_SLOTS = 64
_count_state = {}


def _count_slot(dev):
    st = _count_state.get(dev.index)
    if st is None:
        buf = torch.empty((_COUNT_SLOTS,), dtype=torch.int64).pin_memory()
        st = _count_state[dev.index] = [buf, buf.numpy(), 0]
    st[2] = (st[2] + 1) % _COUNT_SLOTS
    return st[0], st[1], st[2]


def _wait_count(arr, k, dev):
    for _ in range(2_000_000):  # ~0.1 s of polling, far beyond any kernel of this library
        v = arr[k]
        if v >= 0:
            return int(v)
    torch.cuda.current_stream(dev).synchronize()
    if arr[k] < 0:
        raise _lib.FeosB200Error("compaction count did not arrive")
    return int(arr[k])


def compact(failed, values, out_cols=None):
    """Rows of ``values`` [n] or [n, c] whose state converged, in input order, plus the row map
    ``pos`` [n] (int32, -1 = failed).  One host wait (the output length)."""
    n = failed.shape[0]
    dev = failed.device
    in_cols = 1 if values.dim() == 1 else values.shape[1]
    out_cols = in_cols if out_cols is None else out_cols
    pos = torch.empty((n,), dtype=torch.int32, device=dev)
    out = torch.empty((n, out_cols) if values.dim() == 2 else (n,), dtype=torch.float64, device=dev)
    if n == 0:
        return out, pos
    buf, arr, k = _count_slot(dev)
    arr[k] = -1
    _lib.call("feos_dev_compact", _lib.ptr(failed), n, _lib.ptr(pos), _lib.ptr_at(buf, 8 * k), _lib.ptr(values), in_cols,
              _lib.ptr(out), out_cols, _stream())
    return out[:_wait_count(arr, k, dev)], pos


def _upstream(g):
    """Upstream gradient for the backward kernels: (contiguous tensor, broadcast flag).  The gradient of
    `values.sum()` arrives as ONE value expanded with stride 0; the kernels read it in place (g_broadcast)."""
    if g.numel() > 1 and all(st == 0 for st in g.stride()):
        return g.as_strided((1,), (1,), g.storage_offset()), 1
    return g.contiguous(), 0


def _state_order(params, temperature, key):
    """Permutation that puts states of the same code path and similar (m, T / T_c) into one warp
    (csrc/order.cuh), for batches large enough to be throughput-bound; ``None`` = input order.  With a chunked
    host upload every chunk gets its own permutation (`_chunk_order`).  ``key``: the caller's own tensor objects,
    by which the data set is recognised in the next iteration of a fit loop."""
    n = temperature.shape[0]
    if not _lib.state_order_enabled(n) or _chunks_of(params) is not None:
        return None

    def compute(out):
        o = out if out is not None else torch.empty((n,), dtype=torch.int32, device=params.device)
        _lib.timed_call("feos_dev_pure_order", _lib.ptr(params), _lib.ptr(temperature), n, _lib.ptr(o), _stream())
        return o
    # key = (parameter tensor, temperature tensor) of the caller: the temperatures identify the data set (hard), the
    # parameters are "soft" -- an optimizer updates them in place every step, see _lib.cached_order
    with _on_device(params.device):
        return _lib.cached_order(("pure", params.device.index), key[1:], compute, n=n, soft=key[:1])


def _chunk_order(params, temperature, key, a, b):
    """State order of rows [a, b) of a chunked host upload (indices relative to a): computed on the chunk once its
    rows have arrived (the caller has made the stream wait for the chunk's event) and kept like the whole-batch
    order -- a fit loop uploads the same pinned tensors every iteration."""
    n = b - a
    if key is None or not _lib.state_order_enabled(n):
        return None

    def compute(out):
        o = out if out is not None else torch.empty((n,), dtype=torch.int32, device=params.device)
        _lib.timed_call("feos_dev_pure_order", _off(params, a), _off(temperature, a), n, _lib.ptr(o), _stream())
        return o
    return _lib.cached_order(("pure-chunk", params.device.index, a, b), key[1:], compute, n=n, soft=key[:1])


class _PureVle(torch.autograd.Function):
    """which = 0: vapor pressure [Pa]; which = 1: saturated liquid density [kmol/m^3].
    Returns (compacted values [n_ok], failed [n] uint8, rho [n, 2])."""

    @staticmethod
    def forward(ctx, params, temperature, which, order, key):
        n = temperature.shape[0]
        dev = params.device
        rho = torch.empty((n, 2), dtype=torch.float64, device=dev)
        val = torch.empty((n,), dtype=torch.float64, device=dev)
        failed = torch.empty((n,), dtype=torch.uint8, device=dev)
        ch = _chunks_of(params)
        with _on_device(dev):
            def launch(a, b):
                # chunked host upload: every chunk is launched on its own rows, in its own state order
                o = order if ch is None else _chunk_order(params, temperature, key, a, b)
                _lib.timed_call(
                    "feos_dev_pure_vle_fwd", _off(params, a), _off(temperature, a), b - a, _off(rho, a),
                    _off(val, a) if which == 0 else None, _off(val, a) if which == 1 else None,
                    _off(failed, a), None, _lib.ptr(o), _stream(),
                )
            _launch_chunked(params, zip(ch.bounds, ch.events) if ch else [((0, n), None)], launch)
            out, pos = compact(failed, val)
        ctx.save_for_backward(params, temperature, rho, failed, pos)
        ctx.which = which
        ctx.to_host = ch is not None  # parameters came from the host: their gradient goes back there
        ctx.mark_non_differentiable(rho, failed)
        ctx.set_materialize_grads(False)  # no zero-filled gradients for the mask / density outputs
        return out, failed, rho

    @staticmethod
    def backward(ctx, g_val, _g_failed, _g_rho):
        params, temperature, rho, failed, pos = ctx.saved_tensors
        n = temperature.shape[0]
        if g_val is None:  # only the mask / density outputs were used downstream
            return (None,) * 5
        g, bc = _upstream(g_val)
        grad_params = torch.empty_like(params)
        grad_t = torch.empty_like(temperature) if ctx.needs_input_grad[1] else None
        with _on_device(params.device):
            def launch(a, b):
                _lib.timed_call(
                    "feos_dev_pure_vle_bwd", _off(params, a), _off(temperature, a), _off(rho, a), _off(failed, a),
                    _off(pos, a), _lib.ptr(g) if ctx.which == 0 else None, _lib.ptr(g) if ctx.which == 1 else None, bc, b - a,
                    _off(grad_params, a), _off(grad_t, a), _stream(),
                )
            if ctx.to_host and n >= 2 * _CHUNK_MIN:
                bounds = _chunk_bounds(n)
                events = _record_chunk_events(bounds, launch, params.device)
                grad_params._feos_chunks = _Chunks(bounds, events, grad_params._version)
            else:
                launch(0, n)
        return grad_params, grad_t, None, None, None


class _PureRhoTp(torch.autograd.Function):
    @staticmethod
    def forward(ctx, params, temperature, pressure, order, key):
        n = temperature.shape[0]
        dev = params.device
        rho = torch.empty((n,), dtype=torch.float64, device=dev)
        val = torch.empty((n,), dtype=torch.float64, device=dev)
        failed = torch.empty((n,), dtype=torch.uint8, device=dev)
        ch = _chunks_of(params)
        with _on_device(dev):
            def launch(a, b):
                o = order if ch is None else _chunk_order(params, temperature, key, a, b)
                _lib.timed_call(
                    "feos_dev_pure_rho_tp_fwd", _off(params, a), _off(temperature, a), _off(pressure, a), b - a,
                    _off(rho, a), _off(val, a), _off(failed, a), None, _lib.ptr(o), _stream(),
                )
            _launch_chunked(params, zip(ch.bounds, ch.events) if ch else [((0, n), None)], launch)
            out, pos = compact(failed, val)
        ctx.save_for_backward(params, temperature, pressure, rho, failed, pos)
        ctx.to_host = ch is not None
        ctx.mark_non_differentiable(rho, failed)
        ctx.set_materialize_grads(False)  # no zero-filled gradients for the mask / density outputs
        return out, failed, rho

    @staticmethod
    def backward(ctx, g_val, _g_failed, _g_rho):
        params, temperature, pressure, rho, failed, pos = ctx.saved_tensors
        n = temperature.shape[0]
        if g_val is None:  # only the mask / density outputs were used downstream
            return (None,) * 5
        g, bc = _upstream(g_val)
        need_p, need_t, need_pr = ctx.needs_input_grad[:3]
        grad_params = torch.empty_like(params) if need_p else None
        grad_t = torch.empty_like(temperature) if need_t else None
        grad_pr = torch.empty_like(pressure) if need_pr else None
        with _on_device(params.device):
            def launch(a, b):
                _lib.timed_call(
                    "feos_dev_pure_rho_tp_bwd", _off(params, a), _off(temperature, a), _off(pressure, a),
                    _off(rho, a), _off(failed, a), _off(pos, a), _lib.ptr(g), bc, b - a, _off(grad_params, a),
                    _off(grad_t, a), _off(grad_pr, a), _stream(),
                )
            if ctx.to_host and need_p and n >= 2 * _CHUNK_MIN:
                bounds = _chunk_bounds(n)
                events = _record_chunk_events(bounds, launch, params.device)
                grad_params._feos_chunks = _Chunks(bounds, events, grad_params._version)
            else:
                launch(0, n)
        return grad_params, grad_t, grad_pr, None, None


def _grad_or_none(g):
    return None if g is None else g.contiguous()


class _PureState(torch.autograd.Function):
    """(a, p, dp) = derivatives(T, rho); backward = one kernel over nested dual numbers (csrc/state_grad.cuh)."""

    @staticmethod
    def forward(ctx, params, temperature, density):
        n = temperature.shape[0]
        a, p, dp = (torch.empty((n,), dtype=torch.float64, device=params.device) for _ in range(3))
        with _on_device(params.device):
            _lib.call("feos_dev_pure_state", _lib.ptr(params), _lib.ptr(temperature), _lib.ptr(density), n,
                      _lib.ptr(a), _lib.ptr(p), _lib.ptr(dp), _stream())
        ctx.save_for_backward(params, temperature, density)
        return a, p, dp

    @staticmethod
    def backward(ctx, ga, gp, gdp):
        params, temperature, density = ctx.saved_tensors
        n = temperature.shape[0]
        ga, gp, gdp = _grad_or_none(ga), _grad_or_none(gp), _grad_or_none(gdp)
        need_p, need_t, need_r = ctx.needs_input_grad
        grad_p = torch.empty_like(params) if need_p else None
        grad_t = torch.empty_like(temperature) if need_t else None
        grad_r = torch.empty_like(density) if need_r else None
        if n:
            with _on_device(params.device):
                _lib.call("feos_dev_pure_state_bwd", _lib.ptr(params), _lib.ptr(temperature), _lib.ptr(density), n,
                          _lib.ptr(ga), _lib.ptr(gp), _lib.ptr(gdp), _lib.ptr(grad_p), _lib.ptr(grad_t), _lib.ptr(grad_r),
                          _stream())
        return grad_p, grad_t, grad_r


class PcSaftPure:
    """parameters [N, 8] = m, sigma, epsilon_k, mu, kappa_ab, epsilon_k_ab, na, nb (README.md:12)."""

    def __init__(self, parameters):
        if parameters.dim() != 2 or parameters.shape[1] != 8:
            raise ValueError("parameters must have shape [N, 8]")
        _lib.load()  # fail loudly before any work if the CUDA extension is missing
        self._n = parameters.shape[0]
        self._out_device = parameters.device
        self._dev = _cuda_device(parameters)
        self._source = parameters
        # host parameters are copied at the first property call, AFTER that call's (small) state tensors: the
        # copy engine works in submission order and the first chunk's kernel needs its temperatures
        self._params_dev = None if (not parameters.is_cuda and parameters.requires_grad) else _prep(parameters, self._dev)
        self._uploads = {}

    @property
    def _params(self):
        if self._params_dev is None:
            self._params_dev = _prep(self._source, self._dev, chunked=True)  # the consumers below wait per chunk
        return self._params_dev

    @property
    def parameters(self):
        """numpy copy of the parameters (pcsaft_pure.py:104), made on demand: the kernels read the
        device tensor, so no device->host copy happens on the hot path."""
        return self._source.detach().cpu().numpy()

    def _back(self, t):
        return _back(t, self._out_device)

    def _state(self, t):
        """State tensor (T, p) on the device.  A host tensor that carries no gradient is uploaded once per model
        object and tensor version: `vapor_pressure(T)` followed by `liquid_density(T, p)` moves T once."""
        if t.is_cuda or t.requires_grad:
            return _prep(t, self._dev)
        hit = self._uploads.get(id(t))
        if hit is not None and hit[0]() is t and hit[1] == t._version:
            return hit[2]
        d = _prep(t, self._dev)
        for k in [k for k, v in self._uploads.items() if v[0]() is None]:
            del self._uploads[k]  # tensors that are gone (their id may be reused)
        self._uploads[id(t)] = (weakref.ref(t), t._version, d)
        return d

    # ---- properties on the phase-equilibrium hot path ------------------------------------------
    def vapor_pressure(self, temperature):
        """(failed mask [N], p_sat [n_ok] in Pa) -- pcsaft_pure.py:201-215."""
        _check_rows(self._n, temperature=temperature)
        t = self._state(temperature)
        key = (self._source, temperature)
        val, failed, _ = _PureVle.apply(self._params, t, 0, _state_order(self._params, t, key), key)
        return self._back(failed.view(torch.bool)), self._back(val)

    def equilibrium_liquid_density(self, temperature):
        """(failed mask [N], rho_L^sat [n_ok] in kmol/m^3) -- pcsaft_pure.py:217-233."""
        _check_rows(self._n, temperature=temperature)
        t = self._state(temperature)
        key = (self._source, temperature)
        val, failed, _ = _PureVle.apply(self._params, t, 1, _state_order(self._params, t, key), key)
        return self._back(failed.view(torch.bool)), self._back(val)

    def liquid_density(self, temperature, pressure):
        """(failed mask [N], rho_L(T, p) [n_ok] in kmol/m^3) -- pcsaft_pure.py:184-199."""
        _check_rows(self._n, temperature=temperature, pressure=pressure)
        t = self._state(temperature)
        p = self._state(pressure)
        key = (self._source, temperature)
        val, failed, _ = _PureRhoTp.apply(self._params, t, p, _state_order(self._params, t, key), key)
        return self._back(failed.view(torch.bool)), self._back(val)

    # ---- state functions at given (T, rho), differentiable ---------------
    def derivatives(self, temperature, density):
        """(a, p/kT, d(p/kT)/drho) at (T, rho [1/A^3]) -- pcsaft_pure.py:180-182; carries gradients with
        respect to the parameters, T and rho like the reference's Dual3 evaluation under torch autograd."""
        _check_rows(self._n, temperature=temperature, density=density)
        t = _prep(temperature, self._dev)
        r = _prep(density, self._dev)
        a, p, dp = _PureState.apply(_wait_chunks(self._params), t, r)
        return self._back(a), self._back(p), self._back(dp)

    def helmholtz_energy(self, temperature, density):
        """a = A/(V kT) in 1/A^3 -- pcsaft_pure.py:106-178."""
        return self.derivatives(temperature, density)[0]
