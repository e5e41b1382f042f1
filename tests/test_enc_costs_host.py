"""The encoder's rate tables (alfalfa_b200/csrc/enc_costs.cc: what k_enc_rd prices modes and motion vectors
with) against the UNMODIFIED reference's Costs class (encoder/costs.cc:64-221), dumped by
oracle/_ref/ref_costs: sub-block mode costs, 16x16 mode costs, the census-dependent costs of ZEROMV /
NEARESTMV / NEARMV / NEWMV for every count vector, motion-vector component costs, SAD-search costs.
Where the reference tree is not built (GPU box without /root/reference) the committed golden dump
tests/golden/enc_costs.txt is used.  No GPU needed."""
import ctypes as C
import os
import subprocess
import tempfile

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_COSTS = os.path.join(ROOT, "oracle", "_ref", "ref_costs")
GOLDEN = os.path.join(ROOT, "tests", "golden", "enc_costs.txt")


class EncTables(C.Structure):
    _fields_ = [("bmode_cost", C.c_uint16 * 1000), ("ymode_cost", C.c_uint16 * 10), ("mvref_zero", C.c_uint16 * 24),
                ("mvref_one", C.c_uint16 * 24), ("mv_mag_cost", C.c_uint16 * 2048), ("mv_sign_cost", C.c_uint16 * 4),
                ("mv_sad_cost", C.c_uint16 * 256), ("pad", C.c_uint16 * 2)]


@pytest.fixture(scope="module")
def tables():
    d = tempfile.mkdtemp()
    so = os.path.join(d, "enc_costs.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", os.path.join(ROOT, "tests", "enc_costs_shim.cc"),
                           os.path.join(ROOT, "alfalfa_b200", "csrc", "enc_costs.cc"), "-o", so])
    L = C.CDLL(so)
    assert L.ec_size() == C.sizeof(EncTables)
    t = EncTables()
    L.ec_build(C.byref(t))
    return L, t


def reference_dump():
    if os.path.exists(REF_COSTS):
        text = subprocess.run([REF_COSTS], capture_output=True, text=True, check=True).stdout
        if os.path.exists(GOLDEN):
            assert text == open(GOLDEN).read(), "tests/golden/enc_costs.txt is stale"
        return text
    return open(GOLDEN).read()


def test_rate_tables_equal_the_reference(tables):
    _, t = tables
    ref = {line.split()[0]: np.array(line.split()[1:], dtype=np.int64) for line in reference_dump().splitlines()}
    assert np.array_equal(np.array(t.bmode_cost[:]), ref["bmode"])
    assert np.array_equal(np.array(t.ymode_cost[:]), ref["ymode"])
    z, o = np.array(t.mvref_zero[:]).reshape(4, 6), np.array(t.mvref_one[:]).reshape(4, 6)
    want = ref["mvref"].reshape(6, 6, 6, 4)
    for c0 in range(6):
        for c1 in range(6):
            for c2 in range(6):
                got = [z[0, c0], o[0, c0] + z[1, c1], o[0, c0] + o[1, c1] + z[2, c2], o[0, c0] + o[1, c1] + o[2, c2] + z[3, 0]]
                assert list(want[c0, c1, c2]) == got, (c0, c1, c2)
    mag, sign = np.array(t.mv_mag_cost[:]).reshape(2, 1024), np.array(t.mv_sign_cost[:]).reshape(2, 2)
    comp = ref["mvcomp"].reshape(2, 2, 1024)
    for c in range(2):
        for s in range(2):
            got = mag[c] + np.where(np.arange(1024) > 0, sign[c, s], 0)
            assert np.array_equal(got, comp[c, s]), (c, s)
    assert np.array_equal(np.array(t.mv_sad_cost[:]), ref["mvsad"])


def test_rd_multipliers(tables):
    """Encoder::update_rd_multipliers (encoder.cc:179-194)"""
    L, _ = tables
    for y_ac in (4, 10, 18, 19, 30, 58, 101, 157, 200):
        rm, dm = C.c_uint(0), C.c_uint(0)
        L.ec_rd(y_ac, C.byref(rm), C.byref(dm))
        q = min(y_ac, 160.0)
        want = int(q * q * 2.80)
        want_rm, want_dm = (want // 100, 1) if want > 1000 else (want, 100)
        assert (rm.value, dm.value) == (want_rm, want_dm), y_ac


def test_trellis_tables_equal_the_reference(tables):
    """two-pass key frames (encoder.cc:220-408): token costs of the default probabilities (Costs::fill_token_costs) and
    Costs::coeff_base_cost of every coefficient value"""
    L, _ = tables

    class TrellisTables(C.Structure):
        _fields_ = [("token_cost", C.c_uint16 * (4 * 8 * 3 * 12)), ("value_cost", C.c_uint16 * 4096)]
    assert L.ec_trellis_size() == C.sizeof(TrellisTables)
    t = TrellisTables()
    L.ec_trellis(C.byref(t))
    ref = {line.split()[0]: np.array(line.split()[1:], dtype=np.int64) for line in reference_dump().splitlines()}
    assert np.array_equal(np.array(t.token_cost[:]), ref["tokcost"])
    assert np.array_equal(np.array(t.value_cost[:]), ref["valcost"])
