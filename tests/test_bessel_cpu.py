"""The FP64 Matern/Bessel device function (csrc/bessel.cuh) compiled for the host with g++ and checked
against scipy's AMOS-based kvp route the reference uses (magi_v2.py:787-815) and against mpmath."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle import magi_oracle as mo

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def host_lib(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("bessel") / "bessel_host.so")
    subprocess.run(["g++", "-O2", "-shared", "-fPIC", "-o", so, os.path.join(ROOT, "tests", "harness", "bessel_host.cpp")],
                   check=True)
    return ctypes.CDLL(so)


def _matern(lib, nu, phi1, phi2, l):
    l = np.ascontiguousarray(l, dtype=np.float64)
    out = [np.zeros_like(l) for _ in range(3)]
    P = ctypes.POINTER(ctypes.c_double)
    rc = lib.matern_lag_host(ctypes.c_double(nu), ctypes.c_double(phi1), ctypes.c_double(phi2),
                             l.ctypes.data_as(P), len(l), *[o.ctypes.data_as(P) for o in out])
    assert rc == 0
    return out


@pytest.mark.parametrize("nu", [2.01, 2.5, 3.0, 1.2])
@pytest.mark.parametrize("phi2", [0.05, 0.23, 1.0])
def test_against_reference_formulas(host_lib, nu, phi2):
    I = np.concatenate([[0.0], np.logspace(-3, 1.3, 50)])
    Kap, pK, Kpp = mo.matern_blocks(I, 0.02, phi2, nu)
    kap, dk, d2k = _matern(host_lib, nu, 0.02, phi2, I[1:])
    M_pK, M_Kpp = mo.matern_blocks_roundoff_scale(I, 0.02, phi2, nu)     # the reference formulas cancel
    eps = np.finfo(np.float64).eps
    ok = Kap[0, 1:] > 1e-250
    assert np.max(np.abs(kap - Kap[0, 1:])[ok] / np.abs(Kap[0, 1:])[ok]) < 5e-13
    assert np.all(np.abs(dk - pK[1:, 0]) <= 32 * eps * M_pK[1:, 0] + 1e-13 * np.abs(pK).max())
    assert np.all(np.abs(-d2k - Kpp[0, 1:]) <= 32 * eps * M_Kpp[0, 1:] + 1e-13 * np.abs(Kpp).max())


def test_against_mpmath(host_lib):
    mp = pytest.importorskip("mpmath")
    mp.mp.dps = 40
    nu, p1, p2 = mp.mpf("2.01"), mp.mpf("0.02"), mp.mpf("0.23")
    a = mp.sqrt(2 * nu) / p2
    c = p1 * 2 ** (1 - nu) / mp.gamma(nu)
    ls = [0.001, 0.025, 0.1, 0.5, 1.0, 4.0, 20.0]
    kap, dk, d2k = _matern(host_lib, 2.01, 0.02, 0.23, ls)
    for i, l in enumerate(ls):
        u = a * mp.mpf(l)
        t0 = c * u ** nu * mp.besselk(nu, u)
        t1 = -c * a * u ** nu * mp.besselk(nu - 1, u)
        t2 = c * a * a * u ** (nu - 1) * (u * mp.besselk(nu - 2, u) - mp.besselk(nu - 1, u))
        for got, want in ((kap[i], t0), (dk[i], t1), (d2k[i], t2)):
            assert abs(got - float(want)) <= 2e-13 * abs(float(want))


def test_large_lags_underflow_to_zero_without_nan(host_lib):
    kap, dk, d2k = _matern(host_lib, 2.01, 0.02, 0.05, [30.0, 100.0, 1e4])
    assert np.all(np.isfinite(kap)) and np.all(np.isfinite(dk)) and np.all(np.isfinite(d2k))
    assert kap[-1] == 0.0 and dk[-1] == 0.0


def test_rejects_invalid_smoothness(host_lib):
    z = np.ones(1)
    P = ctypes.POINTER(ctypes.c_double)
    assert host_lib.matern_lag_host(ctypes.c_double(0.5), ctypes.c_double(1.0), ctypes.c_double(1.0),
                                    z.ctypes.data_as(P), 1, z.ctypes.data_as(P), z.ctypes.data_as(P),
                                    z.ctypes.data_as(P)) == -1
