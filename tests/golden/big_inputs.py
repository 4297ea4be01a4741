"""Inputs of the full-size golden cases, reproducible on any host (no BLAS in y)."""
import hashlib

import numpy as np


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def c3_inputs():
    """BASELINE.json configs[2]: make_regression(1 000 000, 100, random_state=0).  X comes from the
    RandomState stream alone (checked by digest); y = X @ coef goes through BLAS and may differ in
    the last bits between CPUs, which the L1 root split does not depend on."""
    from sklearn.datasets import make_regression
    return make_regression(n_samples=1_000_000, n_features=100, random_state=0)


def c4_shape_inputs(n, F, seed=0, informative=10):
    """The make_regression recipe of configs[3] (standard-normal X, `informative` columns with
    coefficients 100 U(0,1), no noise) with y summed column by column in a fixed order, so every
    bit of y is reproducible (elementwise IEEE multiply / add only)."""
    rs = np.random.RandomState(seed)
    X = rs.standard_normal((n, F))
    cols = rs.permutation(F)[:informative]
    coef = 100.0 * rs.uniform(size=informative)
    y = np.zeros(n)
    for c, w in zip(cols, coef):
        y += X[:, c] * w
    return X, y
