"""csrc/select.cu — the sort-free :cemppi elite selection — against the reference's definition (POL:455-461):
    order = sortperm(cost); elite = order[1:m]; stop = maximum(abs.(diff(cost[elite]))) < 10e-3
Integer work: the elite SET and the stop decision must be bit-exact, including ties (stable order by index),
−0.0 < +0.0, NaN last, ±Inf, and gaps straddling the 10e-3 threshold."""
import numpy as np
import pytest
from conftest import configure, engine_kwargs, julia_sortperm, make_env

from mpopis_b200.engine import Engine

pytestmark = pytest.mark.gpu


def reference(costs, m, k0=0, kloc=None):
    c = np.asarray(costs, dtype=np.float64)
    kloc = c.size - k0 if kloc is None else kloc
    order = julia_sortperm_nan_last(c)
    elite = order[:m]
    with np.errstate(invalid="ignore"):
        d = np.abs(np.diff(c[elite]))
    mx = np.nan if np.isnan(d).any() else (d.max() if d.size else -np.inf)
    stop = bool(mx < 10e-3)  # NaN < x is False, like Julia
    ids = np.sort(elite[(elite >= k0) & (elite < k0 + kloc)])
    return ids, stop


def julia_sortperm_nan_last(x):
    """conftest.julia_sortperm + Base.isless's NaN rule (every NaN after +Inf, stable among themselves)."""
    x = np.asarray(x, dtype=np.float64)
    nan = np.isnan(x)
    fin = np.where(~nan)[0]
    return np.concatenate([fin[julia_sortperm(x[fin])], np.where(nan)[0]])


@pytest.fixture(scope="module")
def eng(gpu_bound):
    env = make_env("car")
    return configure(Engine(gpu_bound, **engine_kwargs("cemppi", env, 256, 10, 3)), env, "cemppi")


def cases():
    rng = np.random.default_rng(20261017)
    K = 65536
    yield "normal", rng.normal(1000, 50, K), 13107
    yield "wide-range", rng.normal(0, 1, K) * 10.0 ** rng.integers(-3, 7, K), 13107
    yield "penalties", rng.normal(-2000, 30, K) + 1e6 * (rng.random(K) < 0.3) + 5000 * (rng.random(K) < 0.2), 13107
    yield "integer-ties", rng.integers(0, 40, K).astype(float), 13107
    yield "all-equal", np.full(K, 3.25), 13107
    yield "two-values", np.where(rng.random(K) < 0.5, 1.0, 1.005), 13107
    yield "signed-zeros", np.where(rng.random(K) < 0.5, 0.0, -0.0), 40000
    z = rng.normal(0, 1, K)
    z[rng.integers(0, K, 50)] = np.nan
    yield "some-nan", z.copy(), 13107
    yield "mostly-nan", np.where(rng.random(K) < 0.9, np.nan, rng.normal(0, 1, K)), 13107
    z = rng.normal(0, 1, K)
    z[:20] = -np.inf
    z[20:40] = np.inf
    yield "infs", z, 13107
    yield "all-minus-inf-elites", np.where(np.arange(K) % 3 == 0, -np.inf, rng.normal(0, 1, K)), 13107
    # converged elites: spacing just below / just above the threshold
    base = np.arange(K) * 0.0099
    yield "gaps-below", rng.permutation(base), 13107
    base = np.arange(K) * 0.0099
    base[7000:] += 0.0002  # one gap of 0.0101 inside the elite range
    yield "one-gap-above", rng.permutation(base), 13107
    base = np.arange(K) * 0.0099
    base[13107:] += 5.0  # the large gap sits just OUTSIDE the elites
    yield "gap-outside-elites", rng.permutation(base), 13107
    yield "tight-cluster", 5.0 + rng.random(K) * 1e-4, 13107
    yield "cluster-plus-outlier", np.concatenate([[4.0], 5.0 + rng.random(K - 1) * 1e-3]), 13107
    yield "small-K", rng.normal(0, 1, 3000), 600
    yield "ragged-K", rng.normal(0, 1, 70001), 14000
    yield "m-equals-K", rng.normal(0, 1, 5000), 5000
    yield "m-2", rng.normal(0, 1, 5000), 2
    yield "big", rng.normal(500, 20, 1 << 20), 209715
    yield "big-converged", 7.0 + np.arange(1 << 20) * 1e-9, 209715


@pytest.mark.parametrize("name,costs,m", list(cases()), ids=[c[0] for c in cases()])
def test_elite_set_and_stop_decision(eng, name, costs, m):
    ids_ref, stop_ref = reference(costs, m)
    ids, stop, tau = eng.elite_select(costs, m)
    assert np.array_equal(ids, ids_ref), f"{name}: elite sets differ ({ids.size} vs {ids_ref.size})"
    assert stop == stop_ref, f"{name}: stop {stop} vs reference {stop_ref} (tau={tau})"
    assert not eng.elite_select(costs, m, early_stop=False)[1]


@pytest.mark.parametrize("G", [2, 3, 8])
def test_ownership_windows_partition_the_elite_set(eng, G):
    rng = np.random.default_rng(G)
    K = 8 * 3 * 1024
    costs = rng.normal(0, 1, K).round(2)  # plenty of ties across the shard boundaries
    m = 4915
    full, stop_full = reference(costs, m)
    got = []
    for r in range(G):
        ids, stop, _ = eng.elite_select(costs, m, k0=r * (K // G), kloc=K // G)
        assert stop == stop_full
        assert np.array_equal(ids, reference(costs, m, r * (K // G), K // G)[0])
        got.append(ids)
    assert np.array_equal(np.concatenate(got), full)
