"""The arithmetic the CUDA kernels use (alfalfa_b200/csrc/vp8_math.cuh, bpred_lut.inc), compiled
for the CPU, against the oracle's restatement of the reference on random and extreme inputs:
IDCT (transform.cc:100-137), IWHT (:47-88), loop-filter edges (loopfilter_filters.hh:50-183), the ten
4x4 intra modes (prediction.cc:469-643) and six-tap prediction incl. skipped identity passes
(prediction.cc:645-653, 919-971).  Bit-exact."""
import ctypes as C
import os
import subprocess
import tempfile

import numpy as np
import pytest

import oracle_lib as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def libs():
    d = tempfile.mkdtemp()
    so = os.path.join(d, "math_host.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", os.path.join(ROOT, "tests", "math_host_shim.cc"),
                           "-o", so])
    return C.CDLL(so), O.lib()


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def test_idct_add(libs):
    M, L = libs
    rng = np.random.default_rng(1)
    for it in range(3000):
        scale = [8, 200, 2000, 32767][it % 4]
        c = rng.integers(-scale, scale + 1, 16).astype(np.int16)
        if it % 7 == 0:
            c[1:] = 0
        px = rng.integers(0, 256, 16).astype(np.uint8)
        a, b = px.copy(), px.copy()
        M.m_idct_add(_p(c), _p(a))
        L.vp8o_test_idct_add(_p(c), _p(b))
        assert np.array_equal(a, b), (it, c)


def test_iwht(libs):
    M, L = libs
    rng = np.random.default_rng(2)
    for it in range(3000):
        scale = [8, 500, 32767][it % 3]
        c = rng.integers(-scale, scale + 1, 16).astype(np.int16)
        a, b = np.zeros(16, np.int16), np.zeros(16, np.int16)
        M.m_iwht(_p(c), _p(a))
        L.vp8o_test_iwht(_p(c), _p(b))
        assert np.array_equal(a, b), it


def test_loop_filter_edges(libs):
    M, L = libs
    rng = np.random.default_rng(3)
    for it in range(20000):
        base = rng.integers(0, 256)
        spread = [2, 6, 20, 255][it % 4]
        px = np.clip(base + rng.integers(-spread, spread + 1, 8), 0, 255).astype(np.uint8)
        level, sharp, key, mbe = int(rng.integers(1, 64)), int(rng.integers(0, 8)), int(rng.integers(0, 2)), int(rng.integers(0, 2))
        a, b = px.copy(), px.copy()
        M.m_lf_edge(_p(a), level, sharp, key, mbe)
        L.vp8o_test_lf_edge(_p(b), level, sharp, key, mbe)
        assert np.array_equal(a, b), (it, px, level, sharp, key, mbe)


def test_bpred_modes(libs):
    M, L = libs
    rng = np.random.default_rng(4)
    for it in range(2000):
        s = rng.integers(0, 256, 13).astype(np.uint8)
        for mode in range(10):
            a, b = np.zeros(16, np.uint8), np.zeros(16, np.uint8)
            M.m_bpred(mode, _p(s), _p(a))
            L.vp8o_test_bpred(mode, _p(s), _p(b))
            assert np.array_equal(a, b), (it, mode)


@pytest.mark.parametrize("n", [4, 8, 16])
def test_sixtap(libs, n):
    M, L = libs
    rng = np.random.default_rng(5)
    for it in range(300):
        win = rng.integers(0, 256, (n + 5) * (n + 5)).astype(np.uint8)
        if it % 5 == 0:
            win[:] = rng.choice([0, 255], win.size)
        for mx in range(8):
            for my in range(8):
                a, b = np.zeros(n * n, np.uint8), np.zeros(n * n, np.uint8)
                M.m_sixtap(_p(win), n, mx, my, _p(a))
                L.vp8o_test_sixtap(_p(win), n, mx, my, _p(b))
                assert np.array_equal(a, b), (it, mx, my)
                # the packed arithmetic k_inter executes (dp4a rows, biased 16-bit pairs for columns)
                c = np.zeros(n * n, np.uint8)
                M.m_sixtap_packed(_p(win), n, mx, my, _p(c))
                assert np.array_equal(c, b), ("packed", it, mx, my)


def test_add_residual_packed(libs):
    """pixel + residual with saturation on packed words (k_inter's add_residual4) == clamp255(p + r)"""
    M, _ = libs
    rng = np.random.default_rng(8)
    for it in range(4000):
        px = rng.integers(0, 256, 4).astype(np.uint8)
        scale = [3, 300, 32767][it % 3]
        r = rng.integers(-scale - 1, scale + 1, 4).astype(np.int16)
        out = np.zeros(4, np.uint8)
        M.m_add_residual4(_p(px), _p(r), _p(out))
        want = np.clip(px.astype(np.int32) + r.astype(np.int32), 0, 255).astype(np.uint8)
        assert np.array_equal(out, want), (it, px, r)


def test_forward_dct_and_wht(libs):
    """encoder transforms (decoder/dct.cc:45-164)"""
    M, L = libs
    rng = np.random.default_rng(6)
    for it in range(3000):
        src = rng.integers(0, 256, 16).astype(np.uint8)
        pred = rng.integers(0, 256, 16).astype(np.uint8) if it % 3 else src.copy()
        a, b = np.zeros(16, np.int16), np.zeros(16, np.int16)
        M.m_fdct(_p(src), _p(pred), _p(a))
        L.vp8o_test_fdct(_p(src), _p(pred), _p(b))
        assert np.array_equal(a, b), it
        dc = rng.integers(-2040, 2041, 16).astype(np.int16)
        M.m_fwht(_p(dc), _p(a))
        L.vp8o_test_fwht(_p(dc), _p(b))
        assert np.array_equal(a, b), it
