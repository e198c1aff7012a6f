"""The committed golden fixtures (tests/golden/*.npz, written by tests/golden/make_golden.py) pin the ORACLE: rebuilding them
must reproduce the stored arrays.  (They are oracle outputs, not deal.II outputs: the reference ships no vectors for this
program and cannot be built here)"""
import importlib.util
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def _load_generator():
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "golden", "make_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("name", ["pressure", "velocity", "rhs", "solvers", "general"])
def test_oracle_reproduces_golden(name):
    gen = _load_generator()
    fresh = {"pressure": gen.pressure_cases, "velocity": gen.velocity_cases, "rhs": gen.rhs_cases, "solvers": gen.solver_cases,
             "general": gen.general_cases}[name]()
    stored = np.load(os.path.join(HERE, "golden", name + ".npz"))
    assert sorted(stored.files) == sorted(fresh.keys())
    for key in stored.files:
        a, b = stored[key], np.asarray(fresh[key])
        if a.dtype.kind in "iUS":
            assert np.array_equal(a, b), key
        else:
            assert a.shape == b.shape, key
            scale = max(float(np.max(np.abs(a))), 1e-300)
            assert float(np.max(np.abs(a - b))) <= 1e-11 * scale, key
