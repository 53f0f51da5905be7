"""CPU tests of the sort keys behind the bubble / dew state order (csrc/order.cuh, compiled for the host):
the key groups states by code path first, then by temperature bin, then by mole fraction."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "hostsim"))
import sim  # noqa: E402

ROW_PLAIN = [1.5, 3.5, 200.0, 0, 0, 0, 0, 0]
ROW_POLAR = [1.5, 3.5, 200.0, 2.0, 0, 0, 0, 0]
ROW_ASSOC = [1.5, 3.5, 200.0, 0, 0.03, 2500.0, 1, 1]


def test_mix_key_layout():
    rng = np.random.default_rng(0)
    n = 4000
    rows = [ROW_PLAIN + ROW_PLAIN, ROW_POLAR + ROW_PLAIN, ROW_PLAIN + ROW_ASSOC, ROW_ASSOC + ROW_POLAR]
    which = rng.integers(0, 4, n)
    P = np.array([rows[w] for w in which])
    T = rng.uniform(150.0, 400.0, n)
    z = rng.uniform(0.0, 1.0, n)
    keys = sim.mix_order_key(P, T, z)
    bits = sim.order_bits(0)
    assert int(keys.max()) < (1 << bits)
    path = (keys >> np.uint64(bits - 3)).astype(int)
    # bit 0: dipole term, bit 1 / 2: association sites on component 1 / 2
    assert set(path[which == 0]) == {0} and set(path[which == 1]) == {1}
    assert set(path[which == 2]) == {4} and set(path[which == 3]) == {1 | 2}
    # same parameters: the key is monotone in T at fixed x, and in x inside one temperature bin
    Tm = np.linspace(150.0, 400.0, 500)
    k = sim.mix_order_key(np.tile(rows[0], (500, 1)), Tm, np.full(500, 0.3))
    assert np.all(np.diff(k.astype(np.int64)) >= 0)
    zm = np.linspace(0.001, 0.999, 500)
    k = sim.mix_order_key(np.tile([1.0, 3.5, 200.0, 0, 0, 0, 0, 0] * 2, (500, 1)), np.full(500, 250.0), zm)
    assert np.all(np.diff(k.astype(np.int64)) >= 0) and len(set(k >> np.uint64(12))) == 1
    # 64 temperature bins per octave: 1 % steps in T land at most one bin apart
    tb = (sim.mix_order_key(np.tile(rows[0], (2, 1)), np.array([250.0, 252.5]), np.array([0.5, 0.5])) >> np.uint64(12)).astype(int)
    assert 0 <= tb[1] - tb[0] <= 1


def test_degenerate_inputs_do_not_escape_the_key_range():
    P = np.tile(ROW_PLAIN + ROW_PLAIN, (6, 1))
    T = np.array([np.nan, -5.0, 0.0, 1e308, 300.0, 300.0])
    z = np.array([0.5, 0.5, 0.5, 0.5, np.nan, 2.0])
    assert int(sim.mix_order_key(P, T, z).max()) < (1 << sim.order_bits(0))
    k = sim.gc_order_key(np.zeros(6, dtype=np.int32), T, z)
    assert int(k.max()) < (1 << sim.order_bits(1))


def test_gc_key_groups_by_pair():
    rng = np.random.default_rng(1)
    n = 2000
    pair = rng.integers(0, 37, n).astype(np.int32)
    keys = sim.gc_order_key(pair, rng.uniform(150.0, 300.0, n), rng.uniform(0.05, 0.95, n))
    order = np.argsort(keys, kind="stable")
    assert np.all(np.diff(pair[order]) >= 0)
    assert np.array_equal(keys >> np.uint64(sim.order_bits(1)), pair.astype(np.uint64))


def test_pure_key_layout():
    rng = np.random.default_rng(2)
    n = 3000
    P = np.zeros((n, 8))
    P[:, 0], P[:, 1], P[:, 2] = rng.uniform(1, 4, n), rng.uniform(3, 4, n), rng.uniform(150, 350, n)
    polar = rng.random(n) < 0.3
    assoc = rng.random(n) < 0.3
    P[polar, 3] = 2.0
    P[assoc, 4], P[assoc, 5], P[assoc, 6], P[assoc, 7] = 0.03, 2000.0, 1, 1
    tc = P[:, 2] * (0.6380 + 0.7070 * P[:, 0] - 0.06418 * P[:, 0] ** 2)
    T = rng.uniform(0.5, 1.5, n) * tc
    keys, bits = sim.pure_order_key(P, T)
    assert int(keys.max()) < (1 << bits)
    path = (keys >> np.uint64(bits - 2)).astype(int)
    assert np.array_equal(path, polar.astype(int) + 2 * assoc.astype(int))
    # fixed parameters: monotone in T
    Tm = np.linspace(100.0, 700.0, 400)
    k, _ = sim.pure_order_key(np.tile([2.0, 3.5, 250.0, 0, 0, 0, 0, 0], (400, 1)), Tm)
    assert np.all(np.diff(k.astype(np.int64)) >= 0) and len(set(k)) > 300
    # degenerate inputs stay inside the key range
    k, _ = sim.pure_order_key(np.tile([2.0, 3.5, 250.0, 0, 0, 0, 0, 0], (4, 1)), np.array([np.nan, -1.0, 0.0, 1e300]))
    assert int(k.max()) < (1 << bits)
    k, _ = sim.pure_order_key(np.array([[0.0, 3.5, 0.0, 0, 0, 0, 0, 0], [np.nan] * 8, [1e9, 1, 1e-300, 0, 0, 0, 0, 0]]), np.full(3, 300.0))
    assert int(k.max()) < (1 << bits)
