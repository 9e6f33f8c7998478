"""Size-independent properties of the oracle (CPU): they hold for any correct restatement of the reference path
(cosmogp/Gaussian_process.py:13-75, :270-361; cosmogp/pull.py:43-102) and are the same properties the GPU tests use at
sizes where a per-object oracle run would take too long."""
import numpy as np

from oracle import gp_oracle as O

HYP, NUG = [0.7, 2.3], 0.05


def _object(rng, n):
    x = np.sort(rng.uniform(-10, 40, n))
    return x, rng.standard_normal(n), rng.uniform(0.1, 0.4, n)


def test_likelihood_is_invariant_under_reordering_of_the_epochs():
    rng = np.random.default_rng(1)
    for n in (1, 7, 40, 61):
        x, y, ye = _object(rng, n)
        p = rng.permutation(n)
        a = O.log_likelihood(y, x, HYP, NUG, ye)
        b = O.log_likelihood(y[p], x[p], HYP, NUG, ye[p])
        assert abs(a - b) <= 1e-10 * abs(a)


def test_predictive_mean_is_linear_in_the_residuals_and_variance_ignores_them():
    rng = np.random.default_rng(2)
    x, y1, ye = _object(rng, 33)
    y2 = rng.standard_normal(33)
    grid = np.linspace(-12, 42, 29)
    m1, v1 = O.predict(y1, x, HYP, NUG, grid, ye, full_cov=False)
    m2, v2 = O.predict(y2, x, HYP, NUG, grid, ye, full_cov=False)
    m3, v3 = O.predict(2.0 * y1 - 3.0 * y2, x, HYP, NUG, grid, ye, full_cov=False)
    np.testing.assert_allclose(m3, 2.0 * m1 - 3.0 * m2, rtol=0, atol=1e-11)
    np.testing.assert_allclose(v1, v2, rtol=1e-13); np.testing.assert_allclose(v1, v3, rtol=1e-13)
    # a constant mean is added back unchanged: predict(y + c, y0 = c, new_y0 = c) = predict(y) + c
    mc, vc = O.predict(y1 + 1.5, x, HYP, NUG, grid, ye, y0=1.5, new_y0=1.5, full_cov=False)
    np.testing.assert_allclose(mc, m1 + 1.5, rtol=0, atol=1e-11); np.testing.assert_allclose(vc, v1, rtol=1e-13)


def test_variance_diagonal_equals_the_diagonal_of_the_full_covariance_and_shrinks_with_data():
    rng = np.random.default_rng(3)
    x, y, ye = _object(rng, 25)
    grid = np.linspace(-12, 42, 17)
    m, v = O.predict(y, x, HYP, NUG, grid, ye, full_cov=False)
    mf, cov = O.predict(y, x, HYP, NUG, grid, ye, full_cov=True)
    np.testing.assert_allclose(m, mf, rtol=0, atol=1e-12)
    np.testing.assert_allclose(v, np.diag(cov), rtol=1e-11, atol=1e-13)
    assert np.all(v <= HYP[0] ** 2 + NUG ** 2 + 1e-12)            # never above the prior variance
    _, v_more = O.predict(np.r_[y, 0.0], np.r_[x, grid[8]], HYP, NUG, grid, np.r_[ye, 0.1], full_cov=False)
    assert np.all(v_more <= v + 1e-12) and v_more[8] < v[8]         # one more epoch never raises it


def test_closed_form_pulls_equal_the_refits_of_the_reference_loop():
    """pull.py:66-94 refits N times; the closed form on K^-1 (what the device runs) must agree."""
    rng = np.random.default_rng(4)
    for n in (2, 9, 40):
        x, y, ye = _object(rng, n)
        brute = O.loo_bruteforce(y, x, HYP, NUG, ye)
        closed = O.loo_closed_form(y, x, HYP, NUG, ye)
        for a, b in zip(brute[:3], closed[:3]):
            np.testing.assert_allclose(np.asarray(b), np.asarray(a), rtol=1e-8, atol=1e-10)


def test_batched_helpers_sum_to_the_per_object_likelihood_and_match_at_full_width():
    rng = np.random.default_rng(5)
    b, n = 64, 60
    x = np.sort(rng.uniform(-10, 40, (b, n)), axis=1); y = rng.standard_normal((b, n)); ye = rng.uniform(0.1, 0.4, (b, n))
    ll = O.ll_batched_1d(x, y, np.zeros((b, n)), ye, HYP, NUG)
    ref = np.array([O.log_likelihood(y[i], x[i], HYP, NUG, ye[i]) for i in range(b)])
    np.testing.assert_allclose(ll, ref, rtol=1e-11)
    assert abs(ll.sum() - O.log_likelihood_sum(list(y), list(x), HYP, NUG, list(ye))) <= 1e-10 * abs(ll.sum())
