"""Drop-in proof (needs a B200 and the reference-compiled harness in oracle/_ref/).

oracle/_ref/libcrass_dropin.so is the REFERENCE's own code (ReadHolder, StringCheck, kseq, PatternMatcher, the harness
in oracle/refshim) with exactly one translation unit swapped: src/crass/libcrispr.cpp is replaced by the product's
crass_b200/csrc/dropin/libcrispr_b200.cpp, which exports the same C++ functions and runs the two hot loops on the GPU
through libcrass_b200.so.  The reference's containers (ReadMap, StringCheck, lookupTables) filled through it must be
identical to the ones the unmodified reference fills -- that is what lets WorkHorse / NodeManager / the XML writer run
unchanged.  crass-test-b200 is the reference's own Catch test binary linked the same way.
"""
import ctypes as C
import gzip
import os
import subprocess

import pytest

import checkers

pytestmark = pytest.mark.gpu

DROPIN_SO = os.path.join(checkers.ORACLE_DIR, "_ref", "libcrass_dropin.so")
CATCH_B200 = os.path.join(checkers.ORACLE_DIR, "_ref", "crass-test-b200")
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


class DropIn(checkers._Base):
    prefix = "ref_"

    def __init__(self):
        lib = C.CDLL(DROPIN_SO)
        lib.ref_init()
        super().__init__(lib)


@pytest.fixture(scope="module")
def D():
    if not os.path.exists(DROPIN_SO):
        pytest.skip("oracle/_ref/libcrass_dropin.so not built (needs /root/reference at build time)")
    return DropIn()


@pytest.mark.parametrize("name", ["Ill100.fx.gz", "CN_gDC.fa.gz", "Ill.nr.miss.fa.gz", "front_offset_bug.fa.gz", "poor_dr_ext.fa.gz"])
def test_reference_harness_with_swapped_libcrispr(D, name):
    path = os.path.join(checkers.REF_DATA, name)
    want = gzip.open(os.path.join(G, "bundled", name + ".dump.gz")).read().decode("latin-1")
    got, _ = D.run_files([path])
    assert got == want


def test_other_options_through_the_shim(D):
    P = checkers.port()
    path = os.path.join(checkers.REF_DATA, "CN_gDC.fa.gz")
    for prm in (dict(window=6), dict(min_repeats=3), dict(low_spacer=20, high_spacer=60)):
        want, _ = P.run_files([path], prm)
        got, _ = D.run_files([path], prm)
        assert got == want, prm


def test_single_read_entry_points(D):
    """searchCore / scanRight / extendPreRepeat / qcFoundRepeats of the shim, called one read at a time."""
    import random
    import fuzzgen
    P = checkers.port()
    rng = random.Random(31)
    hits = 0
    for _ in range(150):
        s = fuzzgen.fuzz_read(rng, max_len=300)
        a = P.search_core(s)
        b = D.search_core(s)
        assert a[0] == b[0]
        if a[0] == 1:
            hits += 1
            assert a == b
            assert D.qc_found_repeats(s, a[1]) == P.qc_found_repeats(s, a[1])
    assert hits > 10


def test_reference_catch_tests_on_the_gpu():
    if not os.path.exists(CATCH_B200):
        pytest.skip("crass-test-b200 not built")
    out = subprocess.run([CATCH_B200], capture_output=True, text=True, env=dict(os.environ, MALLOC_PERTURB_="255"))
    assert out.returncode == 0 and "139 assertions in 7 test cases" in out.stdout, out.stdout[-2000:]


_CHILD = r"""
import os, sys, hashlib, ctypes as C
sys.path.insert(0, %(tests)r)
import checkers
which, paths = sys.argv[1], sys.argv[2:]
if which == "dropin":
    lib = C.CDLL(%(dropin)r)
    lib.ref_init()
    class D(checkers._Base):
        prefix = "ref_"
    H = D(lib)
else:
    H = checkers.ref()
sys.stdout.flush()
dump, _ = H.run_files(paths)
sys.stdout.flush()
print("\nDUMP_MD5 " + hashlib.md5(dump.encode("latin-1")).hexdigest())
"""


def _run_child(which, paths, env=None):
    """the harness in a process of its own: the shim's engine is a process-wide object created on first use, and the
    progress lines go to the C++ std::cout of that process"""
    import re
    import sys
    code = _CHILD % dict(tests=os.path.dirname(os.path.abspath(__file__)), dropin=DROPIN_SO)
    out = subprocess.run([sys.executable, "-c", code, which] + list(paths), capture_output=True, text=True,
                         env=dict(os.environ, CRASS_REF_SHOW_PROGRESS="1", **(env or {})))
    assert out.returncode == 0, out.stderr[-3000:]
    md5 = re.search(r"DUMP_MD5 (\w+)", out.stdout).group(1)
    ticks = re.findall(r"\[crass_(\w+)\]: Processed (\d+) \.\.\.", out.stdout)
    return md5, ticks


@pytest.mark.parametrize("devices,stream_bytes", [("0", None), ("0,0", None), ("0,0,0,0", None), ("0", 3 << 20), ("0,0", 1 << 20)])
def test_shim_on_several_devices_and_its_progress_lines(D, devices, stream_bytes, tmp_path):
    """CRASS_B200_DEVICES shards the reads of searchFile / findSingletons over the named GPUs (here the same one several
    times); the reference's containers come out the same, and so do the progress lines the reference prints before every
    100 000th read of a file and at the end of each file (libcrispr.cpp:99-109,161-162,495-496,515-516).  With
    stream_bytes the shim's searchFile takes each file through the devices in ranges of that size (parse, K1 and the replay into
    the reference's containers overlapped), and findSingletons scans those ranges: same containers, same lines."""
    import random
    import fuzzgen
    if not checkers.have_ref():
        pytest.skip("oracle/_ref/libcrass_ref.so not built")
    rng = random.Random(77)
    pool = [fuzzgen.rand_seq(rng, rng.randint(24, 40)) for _ in range(5)]
    paths = []
    for f, n in enumerate((230_000, 100_000, 7)):                     # ticks inside a file, a file of exactly one tick, a tiny one
        p = str(tmp_path / ("f%d.fa" % f))
        with open(p, "wb") as fh:
            for i in range(n):
                r = rng.random()
                s = fuzzgen.planted_read(rng, 120, dr=rng.choice(pool)) if r < 0.01 else (fuzzgen.rand_seq(rng, 30) + rng.choice(pool) + fuzzgen.rand_seq(rng, 50) if r < 0.02 else fuzzgen.rand_seq(rng, 100))
                fh.write(b">f%d_%06d\n%s\n" % (f, i, s))
        paths.append(p)
    want_md5, want_ticks = _run_child("ref", paths)
    env = {"CRASS_B200_DEVICES": devices}
    if stream_bytes:
        env["CRASS_B200_STREAM_BYTES"] = str(stream_bytes)
    got_md5, got_ticks = _run_child("dropin", paths, env)
    assert got_md5 == want_md5
    assert got_ticks == want_ticks and len(want_ticks) >= 10
