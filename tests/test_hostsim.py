"""The product's __host__ __device__ per-read code (crass_b200/csrc/*.cuh) compiled by g++ (tests/hostsim) against the
oracle: the very source the CUDA kernels execute, fuzzed on a box without a GPU.  Test infrastructure only -- the product
has no CPU execution path and never loads this library."""
import ctypes as C
import json
import os
import random
import subprocess

import pytest

import checkers
import fuzzgen

HS_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "hostsim")
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def HS():
    subprocess.check_call(["make", "-C", HS_DIR], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return C.CDLL(os.path.join(HS_DIR, "libhostsim.so"))


@pytest.fixture(scope="module")
def P():
    return checkers.port()


def hs_search_core(HS, seq, params=None):
    cap = 4 * len(seq) + 64
    ss = (C.c_uint32 * cap)()
    n, rl = C.c_uint32(0), C.c_uint32(0)
    r = HS.hs_search_core(seq, len(seq), checkers.params_array(params), ss, cap, C.byref(n), C.byref(rl))
    return r, list(ss[: n.value]) if r == 1 else [], rl.value if r == 1 else 0


def hs_packed(HS, seq, shift, tail, mode):
    nw = 7 if len(seq) <= 112 else 10 if len(seq) <= 160 else 16 if len(seq) <= 256 else 19
    ss = (C.c_uint32 * 64)()
    n, rl = C.c_uint32(0), C.c_uint32(0)
    r = HS.hs_packed(seq, len(seq), shift, nw, tail, len(tail), mode, ss, C.byref(n), C.byref(rl))
    return r, list(ss[: n.value]) if r == 1 else [], rl.value if r == 1 else 0


def test_search_core_device_code_matches_the_oracle(HS, P):
    rng = random.Random(3)
    found = 0
    for _ in range(1500):
        seq = fuzzgen.fuzz_read(rng, 400)
        want = P.search_core(seq)
        got = hs_search_core(HS, seq)
        assert got[0] == (1 if want[0] else 0)
        if want[0]:
            assert (got[1], got[2]) == (list(want[1]), want[2])
            found += 1
    assert found > 100


def test_two_bit_filter_only_over_reports_and_the_packed_exact_path_matches(HS, P):
    """K1a's seed flags on the 2-bit stream (any alignment inside a tile, any bytes behind the read) are a superset of
    searchCore's hits; K1b's packed state machine returns searchCore's answer."""
    rng = random.Random(4)
    found = flagged = 0
    for _ in range(4000):
        seq = fuzzgen.fuzz_read(rng, 304)
        if len(seq) < 16:
            continue
        shift = rng.randint(0, 31)
        tail = fuzzgen.rand_seq(rng, rng.randint(0, 40), b"ACGTN")
        want = P.search_core(seq)
        flag = hs_packed(HS, seq, shift, tail, 0)[0]
        assert flag in (0, 1)
        if want[0]:
            assert flag == 1
        got = hs_packed(HS, seq, shift, tail, 1)
        assert got[0] == (1 if want[0] else 0)
        assert hs_packed(HS, seq, shift, tail, 2) == got     # the staged form (qcFoundRepeats cut at the edit distance)
        if want[0]:
            assert (got[1], got[2]) == (list(want[1]), want[2])
            found += 1
        flagged += flag
    assert found > 100 and flagged >= found


def hs_update_start_stops(HS, seq, ss, front, dr, low):
    arr = (C.c_uint32 * len(ss))(*ss)
    out = (C.c_uint32 * (len(ss) + 4))()
    n = C.c_uint32(0)
    st = HS.hs_update_start_stops(seq, len(seq), arr, len(ss), front, dr, len(dr), low, out, C.byref(n))
    return st, list(out[: n.value])


def hs_smith_waterman(HS, a, b, start, length, sim):
    s, e = C.c_int(0), C.c_int(0)
    ap, al, bp, bl = C.c_uint32(0), C.c_uint32(0), C.c_uint32(0), C.c_uint32(0)
    HS.hs_smith_waterman.argtypes = [C.c_char_p, C.c_uint32, C.c_char_p, C.c_uint32, C.c_int, C.c_int, C.c_double] + [C.c_void_p] * 6
    ok = HS.hs_smith_waterman(a, len(a), b, len(b), start, length, sim, C.byref(s), C.byref(e), C.byref(ap), C.byref(al), C.byref(bp), C.byref(bl))
    second = b[bp.value: bp.value + bl.value]
    return ok, s.value, e.value, al.value, bl.value, b.find(second), b.rfind(second)


def test_partial_repeat_recovery_device_code_matches_the_oracle(HS, P):
    """sw_core.cuh (one score row + forward-propagated end of the predecessor walk) == the full-matrix restatement."""
    rng = random.Random(5)
    grew = 0
    for _ in range(4000):
        seq, ss, front, dr = fuzzgen.uss_case(rng)
        low = rng.choice([26, 26, 20, 35])
        want = P.update_start_stops(seq, ss, front, dr, low)
        got = hs_update_start_stops(HS, seq, ss, front, dr, low)
        if want[0] == -3:
            assert got == (3, [])
            continue
        assert got == (0, want[1])
        grew += len(got[1]) > len(ss)
        start = rng.randint(0, len(seq) - 1)
        length = rng.randint(1, len(seq) - start)
        sim = 0.85 if rng.random() < 0.8 else 0.0
        assert hs_smith_waterman(HS, seq, dr, start, length, sim) == P.smith_waterman(seq, dr, start, length, sim)
    assert grew > 1000
    assert hs_update_start_stops(HS, b"ACGT" * 30, [10], 0, b"ACGTACGTAC", 26)[0] == 1          # odd list
    assert hs_update_start_stops(HS, b"ACGT" * 30, [10, 20], 0, b"A" * 128, 26)[0] == 2         # DR longer than 127


def test_partial_repeat_recovery_golden_vectors(HS):
    g = json.load(open(os.path.join(G, "update_start_stops_vectors.json")))
    for v in g["update_start_stops"]:
        assert hs_update_start_stops(HS, v["seq"].encode(), v["ss"], v["front"], v["dr"].encode(), v["low_spacer"]) == (0, v["ss_out"])
    for v in g["smith_waterman"]:
        assert list(hs_smith_waterman(HS, v["a"].encode(), v["b"].encode(), v["start"], v["len"], v["similarity"])) == v["out"]
