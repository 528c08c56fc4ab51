"""The oracle against its own committed regression fixtures (tests/golden/oracle_regression.npz, made by
tests/golden/make_golden.py).  These are NOT reference outputs -- the reference cannot be built here -- they guard the
restatement against silent drift: HEVE p=3 / p=7 + filter, HEVI, the sound-wave column, a panel of the global set, the
tracer advection with its limiters."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))


def test_oracle_reproduces_its_fixtures():
    import make_golden
    gold = np.load(os.path.join(HERE, "golden", "oracle_regression.npz"))
    now = make_golden.compute()
    assert sorted(gold.files) == sorted(now)
    for k in gold.files:
        ref = gold[k]
        scale = max(np.abs(ref).max(), 1e-300)
        # same code, same compiler flags (-ffp-contract=off): equal to the last bits up to OpenMP-independent arithmetic
        assert np.abs(now[k] - ref).max() <= 1e-13 * scale, k
