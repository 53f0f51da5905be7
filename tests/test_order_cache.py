"""Host logic of the state-order cache (feos_torch_b200/_lib.py: cached_order): hard keys, soft (ageing) keys, the
row-count guard.  No GPU and no native call involved."""
import torch

from feos_torch_b200 import _lib


def _fresh():
    _lib._ORDER_CACHE.clear()


def test_hard_key_identity_and_version():
    _fresh()
    T = torch.zeros(8)
    calls = []

    def compute(out=None):
        calls.append(1)
        return torch.arange(8, dtype=torch.int32)

    a = _lib.cached_order("k", (T,), compute, n=8)
    b = _lib.cached_order("k", (T,), compute, n=8)
    assert a is b and len(calls) == 1
    T.add_(1.0)  # the data set changed in place: new order
    _lib.cached_order("k", (T,), compute, n=8)
    assert len(calls) == 2
    T2 = T.clone()  # another tensor object: new order even though the values are equal
    _lib.cached_order("k", (T2,), compute, n=8)
    assert len(calls) == 3
    _lib.cached_order("other", (T2,), compute, n=8)
    assert len(calls) == 4


def test_row_count_guard():
    _fresh()
    T = torch.zeros(8)
    o8 = _lib.cached_order("k", (T,), lambda out: torch.arange(8, dtype=torch.int32), n=8)
    o4 = _lib.cached_order("k", (T,), lambda out: torch.arange(4, dtype=torch.int32), n=4)
    assert o8.shape[0] == 8 and o4.shape[0] == 4


def test_soft_tensors_age_the_entry(monkeypatch):
    """Per-state parameters change every optimizer step: the permutation survives ORDER_REFRESH - 1 changes."""
    _fresh()
    monkeypatch.setattr(_lib, "ORDER_REFRESH", 4)
    T, P = torch.zeros(8), torch.zeros(8, 8)
    calls = []

    def compute(out=None):
        calls.append(1)
        return torch.arange(8, dtype=torch.int32)

    for step in range(9):
        P.add_(0.1)  # optimizer step
        for _ in range(3):  # three property calls per step see the same parameter version
            _lib.cached_order("pure", (T,), compute, n=8, soft=(P,))
    assert len(calls) == 3  # steps 0, 4, 8
    # a new parameter tensor object per step (parameters computed from other leaves) ages the entry the same way
    _fresh()
    calls.clear()
    for step in range(8):
        _lib.cached_order("pure", (T,), compute, n=8, soft=(P.clone(),))
    assert len(calls) == 2
