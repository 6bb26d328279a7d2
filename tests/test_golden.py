"""Golden fixtures (tests/golden/*.npz, oracle-generated regression pins — see make_golden.py)."""
import glob
import os

import numpy as np
import pytest

from tests.util import oracle_kwargs, rel_linf

FIX = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))


def _load(f):
    d = np.load(f, allow_pickle=True)
    opts = tuple(str(o) for o in d["opts"])
    return d, opts


@pytest.mark.parametrize("f", FIX, ids=[os.path.basename(f) for f in FIX])
def test_oracle_reproduces_golden(oracle_c, f):
    d, opts = _load(f)
    y = oracle_c.das_spec(str(d["fun"]), d["Pi"], d["Pr"], d["Pv"], d["Nv"], d["x"], d["t0"], float(d["fs"]), float(d["c"]),
                          interp=str(d["interp"]), **oracle_kwargs(opts))
    assert np.array_equal(y, d["y32"])
    assert rel_linf(y, d["y64"]) < 2e-3


@pytest.mark.gpu
@pytest.mark.parametrize("f", FIX, ids=[os.path.basename(f) for f in FIX])
def test_cuda_matches_golden(f):
    import qups_b200
    d, opts = _load(f)
    y = qups_b200.das_spec(str(d["fun"]), d["Pi"], d["Pr"], d["Pv"], d["Nv"], d["x"], d["t0"], float(d["fs"]), float(d["c"]),
                           *opts, "interp", str(d["interp"]))
    ref = d["y32"][..., 0]
    if str(d["interp"]) == "nearest":
        assert np.array_equal(y, ref)
    else:
        assert rel_linf(y, ref) < 1e-5
