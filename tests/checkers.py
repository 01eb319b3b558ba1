"""ctypes bindings for the two parity checkers (TEST INFRASTRUCTURE, never the product path).

  * ``Port``  -- oracle/libcrass_oracle.so : plain-C restatement (oracle/crass_oracle.c); builds anywhere.
  * ``Ref``   -- oracle/_ref/libcrass_ref.so : the unmodified reference hot path compiled from
                 /root/reference by oracle/Makefile (present when it was built in the dev container;
                 travels to the GPU box with the snapshot).

Both expose the same Python surface so that tests can be written once and run against either.
"""
import ctypes as C
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
REF_SO = os.path.join(ORACLE_DIR, "_ref", "libcrass_ref.so")
PORT_SO = os.path.join(ORACLE_DIR, "libcrass_oracle.so")
REF_DATA = os.path.join(ORACLE_DIR, "_ref", "data")

DEFAULT_PARAMS = dict(low_dr=23, high_dr=47, low_spacer=26, high_spacer=50, window=8, min_repeats=2, kmer_clust=6)
PARAM_ORDER = ("low_dr", "high_dr", "low_spacer", "high_spacer", "window", "min_repeats", "kmer_clust")


def params_array(params=None):
    p = dict(DEFAULT_PARAMS)
    if params:
        p.update(params)
    return (C.c_uint32 * 7)(*[p[k] for k in PARAM_ORDER])


def _u8(b):
    return C.cast(C.c_char_p(b), C.POINTER(C.c_uint8))


class _Base:
    prefix = ""

    def __init__(self, lib):
        self.lib = lib
        f = self._f
        u32p, i32p, cp = C.POINTER(C.c_uint32), C.POINTER(C.c_int), C.c_char_p
        f("search_core", C.c_int, [cp, C.c_uint32, u32p, u32p, C.c_uint32, u32p, u32p])
        f("scan_right", None if self.prefix == "orc_" else C.c_int,
          [cp, C.c_uint32, u32p, u32p, C.c_uint32, cp, C.c_uint32, C.c_uint32, C.c_uint32])
        f("extend_pre_repeat", C.c_uint32 if self.prefix == "orc_" else C.c_int, [cp, C.c_uint32, u32p, C.c_uint32, C.c_int, C.c_int])
        f("qc_found_repeats", C.c_int, [cp, C.c_uint32, u32p, C.c_uint32, C.c_int, C.c_int])
        f("edit_distance", C.c_int, [cp, C.c_uint32, cp, C.c_uint32])
        f("similarity", C.c_float, [cp, C.c_uint32, cp, C.c_uint32])
        f("low_complexity", C.c_int, [cp, C.c_uint32])
        f("revcomp", None, [cp, C.c_uint32, cp])
        f("dr_lowlexi", C.c_int, [cp, C.c_uint32, u32p, C.c_uint32] + ([cp] if self.prefix == "ref_" else []) + [cp, u32p, i32p])
        f("ac_create", C.c_void_p, [C.POINTER(cp), u32p, C.c_uint32])
        f("ac_first_match", C.c_int, [C.c_void_p, cp, C.c_uint32, i32p, i32p])
        f("ac_destroy", None, [C.c_void_p])
        f("kseq_dump", C.c_void_p, [cp])
        f("non_redundant", C.c_void_p, [C.POINTER(cp), u32p, C.c_uint32, C.c_int])
        f("run_files", C.c_void_p, [C.POINTER(cp), C.c_uint32, u32p, C.c_int, C.POINTER(C.c_double)])
        f("free", None, [C.c_void_p])
        f("update_start_stops", C.c_int, [cp, C.c_uint32, u32p, u32p, C.c_uint32, C.c_int, cp, C.c_uint32, C.c_uint32])

    def _f(self, name, restype, argtypes):
        fn = getattr(self.lib, self.prefix + name)
        fn.restype = restype
        fn.argtypes = argtypes
        setattr(self, "_" + name, fn)

    # -- helpers -------------------------------------------------------------------------------
    def _take(self, ptr):
        if not ptr:
            return None
        s = C.string_at(ptr)
        self._free(C.c_void_p(ptr))
        return s

    # -- per-read functions --------------------------------------------------------------------
    def search_core(self, seq, params=None):
        cap = 2 * (len(seq) // 4 + 8)
        ss = (C.c_uint32 * cap)()
        n = C.c_uint32(0)
        rl = C.c_uint32(0)
        r = self._search_core(seq, len(seq), params_array(params), ss, cap, C.byref(n), C.byref(rl))
        return r, list(ss[: n.value]), rl.value

    def scan_right(self, seq, ss, pattern, min_spacer, scan_range=24):
        cap = 2 * (len(seq) // 4 + 8)
        arr = (C.c_uint32 * cap)(*ss)
        n = C.c_uint32(len(ss))
        self._scan_right(seq, len(seq), arr, C.byref(n), cap, pattern, len(pattern), min_spacer, scan_range)
        return list(arr[: n.value])

    def extend_pre_repeat(self, seq, ss, window, min_spacer):
        arr = (C.c_uint32 * len(ss))(*ss)
        r = self._extend_pre_repeat(seq, len(seq), arr, len(ss), window, min_spacer)
        return int(r), list(arr)

    def qc_found_repeats(self, seq, ss, min_spacer=26, max_spacer=50):
        arr = (C.c_uint32 * len(ss))(*ss)
        return self._qc_found_repeats(seq, len(seq), arr, len(ss), min_spacer, max_spacer)

    def edit_distance(self, a, b):
        return self._edit_distance(a, len(a), b, len(b))

    def similarity(self, a, b):
        return self._similarity(a, len(a), b, len(b))

    def low_complexity(self, a):
        return self._low_complexity(a, len(a))

    def revcomp(self, a):
        out = C.create_string_buffer(len(a) + 1)
        self._revcomp(a, len(a), out)
        return out.raw[: len(a)]

    def dr_lowlexi(self, seq, ss):
        arr = (C.c_uint32 * len(ss))(*ss)
        dr = C.create_string_buffer(len(seq) + 2)
        drl = C.c_uint32(0)
        low = C.c_int(0)
        if self.prefix == "ref_":
            so = C.create_string_buffer(len(seq) + 1)
            r = self._dr_lowlexi(seq, len(seq), arr, len(ss), so, dr, C.byref(drl), C.byref(low))
            seq_out = so.raw[: len(seq)]
        else:
            buf = C.create_string_buffer(seq, len(seq) + 1)
            r = self._dr_lowlexi(buf, len(seq), arr, len(ss), dr, C.byref(drl), C.byref(low))
            seq_out = buf.raw[: len(seq)]
        assert r == 0
        return dr.raw[: drl.value], low.value, list(arr), seq_out

    # -- partial-DR recovery (updateStartStops / smithWaterman) ---------------------------------
    def update_start_stops(self, seq, ss, front_offset, dr, low_spacer=26):
        """-> (status, new start/stop list)"""
        cap = len(ss) + 8
        arr = (C.c_uint32 * cap)(*ss)
        n = C.c_uint32(len(ss))
        r = self._update_start_stops(seq, len(seq), arr, C.byref(n), cap, front_offset, dr, len(dr), low_spacer)
        return r, list(arr[: n.value])

    # -- consensus DR of a group (ksw_align + Aligner) ---------------------------------------
    def ksw_align(self, query, target, xtra=0x80000 | 0x40000 | 5):
        """query / target: nt4 code bytes -> (score, te, qe, score2, te2, tb, qb)"""
        out = (C.c_int * 7)()
        fn = getattr(self.lib, self.prefix + "ksw_align")
        fn.restype = C.c_int
        fn.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int, C.c_int, C.POINTER(C.c_int)]
        fn(bytes(query), len(query), bytes(target), len(target), xtra, out)
        return tuple(out)

    def consensus_group(self, case):
        """case: dict(reads=[(seq, ss, dr index)], drs=[bytes, ...] (0 = master), array_len) ->
        dict(status, place, flags, zone, consensus, conservation (uint32 bit patterns), coverage)"""
        import numpy as np
        reads, drs, n = case["reads"], case["drs"], case["array_len"]
        bases = np.frombuffer(b"".join(r[0] for r in reads), dtype=np.uint8).copy() if reads else np.zeros(1, np.uint8)
        offs = np.zeros(len(reads) + 1, dtype=np.uint64)
        offs[1:] = np.cumsum([len(r[0]) for r in reads])
        ss_offs = np.zeros(len(reads) + 1, dtype=np.uint32)
        ss_offs[1:] = np.cumsum([len(r[1]) for r in reads])
        pool = np.array([x for r in reads for x in r[1]] or [0], dtype=np.uint32)
        read_dr = np.array([r[2] for r in reads] or [0], dtype=np.uint32)
        dr_bytes = np.frombuffer(b"".join(drs), dtype=np.uint8).copy()
        dr_offs = np.zeros(len(drs) + 1, dtype=np.uint32)
        dr_offs[1:] = np.cumsum([len(d) for d in drs])
        place = np.zeros(len(drs), dtype=np.int32)
        flags = np.zeros(len(drs), dtype=np.uint8)
        zone = np.zeros(2, dtype=np.int32)
        cons = np.zeros(n, dtype=np.uint8)
        conserv = np.zeros(n, dtype=np.float32)
        cov = np.zeros(4 * n, dtype=np.int32)
        fn = getattr(self.lib, self.prefix + "consensus_group")
        fn.restype = C.c_int
        fn.argtypes = [C.c_void_p] * 2 + [C.c_uint32] + [C.c_void_p] * 5 + [C.c_uint32, C.c_uint32] + [C.c_void_p] * 6
        st = fn(bases.ctypes.data, offs.ctypes.data, len(reads), read_dr.ctypes.data, ss_offs.ctypes.data, pool.ctypes.data,
                dr_bytes.ctypes.data, dr_offs.ctypes.data, len(drs), n, place.ctypes.data, flags.ctypes.data, zone.ctypes.data,
                cons.ctypes.data, conserv.ctypes.data, cov.ctypes.data)
        return dict(status=st, place=place.tolist(), reversed=[int(f) & 1 for f in flags], zone=zone.tolist(),
                    consensus=cons.tobytes(), conservation=conserv.view(np.uint32).tolist(), coverage=cov.tolist())

    # -- automaton -----------------------------------------------------------------------------
    def ac_create(self, patterns):
        n = len(patterns)
        pats = (C.c_char_p * n)(*patterns)
        lens = (C.c_uint32 * n)(*[len(p) for p in patterns])
        h = self._ac_create(pats, lens, n)
        return C.c_void_p(h)

    def ac_first_match(self, h, text):
        e = C.c_int(0)
        l = C.c_int(0)
        r = self._ac_first_match(h, text, len(text), C.byref(e), C.byref(l))
        return (e.value, l.value) if r else None

    def ac_destroy(self, h):
        self._ac_destroy(h)

    # -- file level ----------------------------------------------------------------------------
    def kseq_dump(self, path):
        return self._take(self._kseq_dump(path.encode()))

    def non_redundant(self, drs, min_count=6):
        n = len(drs)
        arr = (C.c_char_p * n)(*drs)
        lens = (C.c_uint32 * n)(*[len(d) for d in drs])
        return self._take(self._non_redundant(arr, lens, n, min_count)).decode()

    def run_files(self, paths, params=None, phases=2):
        n = len(paths)
        arr = (C.c_char_p * n)(*[p.encode() for p in paths])
        tm = (C.c_double * 3)()
        s = self._take(self._run_files(arr, n, params_array(params), phases, tm))
        return s.decode("latin-1"), list(tm)


class Ref(_Base):
    prefix = "ref_"

    def __init__(self):
        lib = C.CDLL(REF_SO)
        lib.ref_init()
        super().__init__(lib)
        i32p, u32p = C.POINTER(C.c_int), C.POINTER(C.c_uint32)
        lib.ref_smith_waterman.restype = C.c_int
        lib.ref_smith_waterman.argtypes = [C.c_char_p, C.c_uint32, C.c_char_p, C.c_uint32, C.c_int, C.c_int, C.c_double,
                                           i32p, i32p, u32p, u32p, i32p, i32p]

    def smith_waterman(self, a, b, start, length, similarity=0.85):
        """-> (ok, a_start_align, a_end_align, len(first), len(second), b.find(second), b.rfind(second))"""
        s, e, f, r = C.c_int(0), C.c_int(0), C.c_int(0), C.c_int(0)
        l1, l2 = C.c_uint32(0), C.c_uint32(0)
        ok = self.lib.ref_smith_waterman(a, len(a), b, len(b), start, length, similarity, C.byref(s), C.byref(e),
                                         C.byref(l1), C.byref(l2), C.byref(f), C.byref(r))
        return ok, s.value, e.value, l1.value, l2.value, f.value, r.value


class Port(_Base):
    prefix = "orc_"

    def __init__(self):
        if not os.path.exists(PORT_SO) or os.path.getmtime(PORT_SO) < os.path.getmtime(os.path.join(ORACLE_DIR, "crass_oracle.c")):
            subprocess.check_call(["make", "-C", ORACLE_DIR, "port"], stdout=subprocess.DEVNULL)
        lib = C.CDLL(PORT_SO)
        super().__init__(lib)
        u64p = C.POINTER(C.c_uint64)
        lib.orc_phase1_batch.restype = C.c_uint64
        lib.orc_phase1_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32), C.c_void_p]
        lib.orc_phase2_batch.restype = C.c_uint64
        lib.orc_phase2_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]
        i32p, u32p = C.POINTER(C.c_int), C.POINTER(C.c_uint32)
        lib.orc_smith_waterman.restype = C.c_int
        lib.orc_smith_waterman.argtypes = [C.c_char_p, C.c_uint32, C.c_char_p, C.c_uint32, C.c_int, C.c_int, C.c_double,
                                           i32p, i32p, u32p, u32p, u32p, u32p]

    def smith_waterman(self, a, b, start, length, similarity=0.85):
        """Same tuple as Ref.smith_waterman (find/rfind of the second string evaluated here)."""
        s, e = C.c_int(0), C.c_int(0)
        ap, al, bp, bl = C.c_uint32(0), C.c_uint32(0), C.c_uint32(0), C.c_uint32(0)
        ok = self.lib.orc_smith_waterman(a, len(a), b, len(b), start, length, similarity, C.byref(s), C.byref(e),
                                         C.byref(ap), C.byref(al), C.byref(bp), C.byref(bl))
        second = b[bp.value: bp.value + bl.value]
        return ok, s.value, e.value, al.value, bl.value, b.find(second), b.rfind(second)

    def _f(self, name, restype, argtypes):
        # orc_search_core takes (const orc_params*) where ref_ takes the uint32[7] array: same layout
        super()._f(name, restype, argtypes)


def have_ref():
    return os.path.exists(REF_SO)


_ref = None
_port = None


def ref():
    global _ref
    if _ref is None:
        _ref = Ref()
    return _ref


def port():
    global _port
    if _port is None:
        _port = Port()
    return _port
