"""Committed golden vectors (tests/golden/*.npz, made by tests/golden/make_golden.py from the oracle --
the reference itself has none): CPU test = the oracle still reproduces them bit-for-bit;
GPU test = the CUDA path reproduces them bit-for-bit through the C ABI."""
import os
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import make_golden as mg  # noqa: E402


def _load(name):
    return np.load(os.path.join(HERE, "golden", name + ".npz"))


@pytest.mark.parametrize("name", sorted(mg.CASES))
def test_oracle_reproduces_golden(orc, name):
    g = _load(name)
    for p, pos, vel, aux in mg.run(name):
        np.testing.assert_array_equal(pos, g["pos_%d" % p])
        np.testing.assert_array_equal(vel, g["vel_%d" % p])
        np.testing.assert_array_equal(aux, g["aux_%d" % p])


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(mg.CASES))
def test_cuda_reproduces_golden(sim, orc, name):
    g = _load(name)
    W, R, fld, mode, flags, kind, periods, spp = mg.CASES[name]
    b = sim.WorldBatch(W, R, fld=fld, init_mode=mode, params=sim.default_params(flags=flags))

    class Tmp:  # same deterministic edit as the generator, applied to the device state
        pass
    t = Tmp()
    t.R = R
    t.pos, vel, aux = b.read_state()
    mg.prepare(name, t)
    b.write_state(t.pos, vel, aux)
    for p in range(periods):
        if kind is not None:
            b.gen_commands(p, kind)
        b.step(spp)
        if "pos_%d" % p in g.files:
            pos, vel, aux = b.read_state()
            np.testing.assert_array_equal(pos, g["pos_%d" % p])
            np.testing.assert_array_equal(vel, g["vel_%d" % p])
            np.testing.assert_array_equal(aux, g["aux_%d" % p])
    assert int(g["aux_%d" % (periods - 1)][:, 7].sum()) > 0
