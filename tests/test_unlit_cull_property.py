"""The "unlit light" reduction rests on one arithmetic fact: when both geometric factors of a
light's Blinn-Phong terms are exactly 0, the terms `((((sh*color)*k)*Od)*att)*0` (Renderer.hpp:290-296, evaluated left
to right in float32) have the same value — including the sign of the zero — for EVERY shadow coefficient sh in [0, 1],
so the coefficient need not be computed.  Checked here on random operands, negative colours and huge / tiny
magnitudes included; the accumulated sum is compared bit for bit."""
import numpy as np


def _terms(sh, color, k, od, att, factor):
    f = np.float32
    t = (sh * color).astype(f)
    t = (t * k).astype(f)
    t = (t * od).astype(f)
    t = (t * att).astype(f)
    return (t * factor).astype(f)


def test_zero_factor_makes_the_shadow_coefficient_irrelevant():
    rng = np.random.default_rng(7)
    n = 200_000
    f = np.float32
    mag = (10.0 ** rng.uniform(-30, 30, n)).astype(f)
    color = (rng.uniform(-1, 1, n).astype(f) * np.where(rng.random(n) < 0.3, mag, f(1))).astype(f)
    k = rng.uniform(-2, 2, n).astype(f)
    od = rng.uniform(-1, 1, n).astype(f)
    att = (10.0 ** rng.uniform(-6, 6, n)).astype(f)
    acc0 = rng.uniform(-1, 1, n).astype(f) * (rng.random(n) < 0.7)          # running sum before this light, often exactly 0
    zero = np.zeros(n, f)
    with np.errstate(over="ignore", under="ignore", invalid="ignore"):
        results = []
        for sh in (np.zeros(n, f), np.ones(n, f), rng.random(n).astype(f), (rng.integers(0, 51, n) / f(50)).astype(f)):
            term = _terms(sh, color, k, od, att, zero)
            results.append((acc0 + term).astype(f))
    base = results[0].view(np.uint32)
    finite = np.isfinite(results[0])
    for r in results[1:]:
        same = (r.view(np.uint32) == base) | (~finite & ~np.isfinite(r))     # inf * 0 = NaN on both sides
        assert same.all()
    # and the sign argument itself: a product of fixed-sign operands with a non-negative sh keeps its sign when sh -> +0
    t1 = _terms(np.full(n, f(0.37)), color, k, od, att, zero)
    t0 = _terms(np.zeros(n, f), color, k, od, att, zero)
    ok = np.isfinite(t1) & np.isfinite(t0)
    assert (np.signbit(t1[ok]) == np.signbit(t0[ok])).all() and (t1[ok] == 0).all()
