"""ctypes binding of the C-ABI in include/crass_b200.h (libcrass_b200.so, built in-tree).

This is the Python face of the product used by tests/, bench.py and the multi-GPU driver; every
compute call goes straight into the CUDA library.  There is no Python or CPU implementation of
the path here: if the shared library is missing the import fails loudly, and on a machine without
a CUDA device every compute entry point raises ``CrassB200Error`` (status ENODEVICE).
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("CRASS_B200_LIB") or os.path.join(_HERE, "libcrass_b200.so")   # override: an instrumented build


class CrassB200Error(RuntimeError):
    def __init__(self, status, msg):
        super().__init__("crass_b200 status %d: %s" % (status, msg))
        self.status = status


class Params(C.Structure):
    """The searched fields of the reference's ``options`` struct (crassDefines.h:140-170)."""
    _fields_ = [("low_dr", C.c_uint32), ("high_dr", C.c_uint32), ("low_spacer", C.c_uint32), ("high_spacer", C.c_uint32),
                ("window", C.c_uint32), ("min_repeats", C.c_uint32), ("kmer_clust", C.c_uint32), ("scan_range", C.c_uint32)]

    def __init__(self, **kw):
        super().__init__()
        lib().crass_b200_default_params(C.byref(self))
        for k, v in kw.items():
            if not hasattr(self, k):
                raise TypeError("unknown parameter %r" % k)
            setattr(self, k, v)

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class Hit(C.Structure):
    _fields_ = [("read_index", C.c_uint32), ("n_ss", C.c_uint32), ("ss_offset", C.c_uint32), ("repeat_len", C.c_uint32)]


HIT_DTYPE = np.dtype([("read_index", "<u4"), ("n_ss", "<u4"), ("ss_offset", "<u4"), ("repeat_len", "<u4")])

USS_JOB_DTYPE = np.dtype([("read", "<u4"), ("ss_offset", "<u4"), ("n_ss", "<u4"), ("front_offset", "<i4"), ("dr", "<u4"), ("out_offset", "<u4")])

EINVAL = -1
ENODEVICE = -2

_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("%s is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "or `make -C crass_b200/csrc` (there is no fallback implementation)" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, u8p, u32p, u64p, cp = C.c_void_p, C.POINTER(C.c_uint8), C.POINTER(C.c_uint32), C.POINTER(C.c_uint64), C.c_char_p
    sig = {
        "crass_b200_last_error": (cp, []),
        "crass_b200_abi_version": (C.c_int, []),
        "crass_b200_build_info": (cp, []),
        "crass_b200_device_count": (C.c_int, []),
        "crass_b200_default_params": (None, [C.POINTER(Params)]),
        "crass_b200_ctx_create": (C.c_int, [C.c_int, C.POINTER(vp)]),
        "crass_b200_ctx_destroy": (None, [vp]),
        "crass_b200_ctx_device": (C.c_int, [vp]),
        "crass_b200_ctx_launch_count": (C.c_uint64, [vp]),
        "crass_b200_ctx_last_candidates": (C.c_uint64, [vp]),
        "crass_b200_ctx_set_token_output": (C.c_int, [vp, vp, C.c_uint32]),
        "crass_b200_ctx_last_dr_list": (cp, [vp]),
        "crass_b200_dr_list_from_tokens": (vp, [vp, C.c_uint32, vp, C.c_uint32]),
        "crass_b200_unique_tokens_dev": (C.c_int, [vp, vp, C.c_uint32, vp, C.c_uint32, vp, vp, vp, vp]),
        "crass_b200_dr_list_from_unique": (vp, [vp, C.c_uint32, vp, C.c_uint32]),
        "crass_b200_dr_search_dev": (C.c_int, [vp, vp, vp, C.c_uint32, C.c_uint32, C.POINTER(Params), vp, vp, C.c_uint32, vp, C.c_uint32, vp, vp]),
        "crass_b200_dr_search": (C.c_int, [vp, vp, vp, C.c_uint32, C.POINTER(Params), vp, C.POINTER(vp), u32p, C.POINTER(vp), u32p]),
        "crass_b200_batch_upload": (C.c_int, [vp, vp, vp, C.c_uint32]),
        "crass_b200_dr_search_resident": (C.c_int, [vp, C.POINTER(Params), vp, C.POINTER(vp), u32p, C.POINTER(vp), u32p]),
        "crass_b200_ac_scan_resident": (C.c_int, [vp, vp, C.c_int, vp, C.POINTER(vp), u32p, C.POINTER(vp), u32p]),
        "crass_b200_dr_list_from_hits": (vp, [vp, vp, C.c_uint32, vp, C.c_uint32, vp]),
        "crass_b200_merge_dr_lists": (vp, [cp]),
        "crass_b200_ac_build": (C.c_int, [vp, vp, C.c_uint32, C.POINTER(vp)]),
        "crass_b200_ac_destroy": (None, [vp]),
        "crass_b200_ac_upload": (C.c_int, [vp, vp]),
        "crass_b200_ac_num_states": (C.c_uint32, [vp]),
        "crass_b200_ac_num_symbols": (C.c_uint32, [vp]),
        "crass_b200_ac_table_bytes": (C.c_uint64, [vp]),
        "crass_b200_ac_pattern_text": (vp, [vp, C.POINTER(C.c_uint32)]),
        "crass_b200_comm_unique_id": (C.c_int, [vp]),
        "crass_b200_ctx_comm_init": (C.c_int, [vp, vp, C.c_int, C.c_int]),
        "crass_b200_ctx_comm_world": (C.c_int, [vp]),
        "crass_b200_exchange_tokens_dev": (C.c_int, [vp, vp, C.c_uint32, vp, C.c_uint32, vp, C.c_uint32, vp, C.c_uint32, vp, C.c_uint32, vp]),
        "crass_b200_ksw_align": (C.c_int, [vp, vp, C.c_uint64, vp, C.c_uint32, vp]),
        "crass_b200_consensus_groups": (C.c_int, [vp, vp, vp, C.c_uint32, vp, vp, vp, vp, vp, C.c_uint32, vp, C.c_uint32, C.c_uint32, vp, vp, vp, vp, vp, vp, C.POINTER(C.c_uint32)]),
        "crass_b200_ac_scan_dev": (C.c_int, [vp, vp, vp, vp, C.c_uint32, C.c_uint32, vp, vp, vp, C.c_uint32, vp, C.c_uint32, vp, vp]),
        "crass_b200_ac_scan": (C.c_int, [vp, vp, vp, vp, C.c_uint32, vp, vp, C.POINTER(vp), u32p, C.POINTER(vp), u32p]),
        "crass_b200_edit_distance_batch": (C.c_int, [vp, vp, C.c_uint64, vp, vp, vp, vp, C.c_uint32, vp, vp]),
        "crass_b200_update_start_stops_dev": (C.c_int, [vp, vp, vp, vp, vp, vp, C.c_uint32, vp, C.c_uint32, vp, vp, vp, vp]),
        "crass_b200_update_start_stops": (C.c_int, [vp, vp, vp, C.c_uint32, vp, vp, C.c_uint32, vp, C.c_uint32, vp, C.c_uint32, C.c_uint32, vp, C.c_uint32, vp, vp]),
        "crass_b200_scan_right": (C.c_int, [vp, cp, C.c_uint32, u32p, u32p, C.c_uint32, cp, C.c_uint32, C.c_uint32, C.c_uint32]),
        "crass_b200_extend_pre_repeat": (C.c_int, [vp, cp, C.c_uint32, u32p, C.c_uint32, C.c_uint32, C.c_uint32, u32p]),
        "crass_b200_qc_found_repeats": (C.c_int, [vp, cp, C.c_uint32, u32p, C.c_uint32, C.c_int, C.c_int, C.POINTER(C.c_int)]),
        "crass_b200_parse_file": (C.c_int, [cp, C.POINTER(vp)]),
        "crass_b200_parse_stream_open": (C.c_int, [cp, C.c_uint64, C.POINTER(vp)]),
        "crass_b200_parse_stream_next": (C.c_int, [vp, C.POINTER(vp)]),
        "crass_b200_parse_stream_close": (None, [vp]),
        "crass_b200_batch_from_memory": (C.c_int, [vp, vp, C.c_uint32, vp, C.POINTER(vp)]),
        "crass_b200_batch_destroy": (None, [vp]),
        "crass_b200_batch_num_reads": (C.c_uint32, [vp]),
        "crass_b200_batch_max_read_len": (C.c_uint32, [vp]),
        "crass_b200_batch_parse_status": (C.c_int, [vp]),
        "crass_b200_batch_bases": (vp, [vp]),
        "crass_b200_batch_read": (vp, [vp, C.c_uint32, C.POINTER(C.c_uint32)]),
        "crass_b200_batch_offsets": (vp, [vp]),
        "crass_b200_batch_name": (cp, [vp, C.c_uint32]),
        "crass_b200_batch_comment": (cp, [vp, C.c_uint32, C.POINTER(C.c_int)]),
        "crass_b200_batch_qual": (cp, [vp, C.c_uint32, C.POINTER(C.c_int)]),
        "crass_b200_results_create": (C.c_int, [C.POINTER(vp)]),
        "crass_b200_results_destroy": (None, [vp]),
        "crass_b200_results_add_phase1": (C.c_int, [vp, vp, vp, C.c_uint32, vp]),
        "crass_b200_results_add_phase2": (C.c_int, [vp, vp, vp, C.c_uint32, vp]),
        "crass_b200_results_num_tokens": (C.c_uint32, [vp]),
        "crass_b200_results_num_reads": (C.c_uint32, [vp]),
        "crass_b200_results_dr_list": (vp, [vp]),
        "crass_b200_results_adopt_tokens": (C.c_int, [vp, cp]),
        "crass_b200_results_non_redundant": (vp, [vp, C.c_uint32, u32p]),
        "crass_b200_results_dump": (vp, [vp, C.c_int]),
        "crass_b200_ctx_keep_packed": (C.c_int, [vp, C.c_int]),
        "crass_b200_ctx_keep_packed_bases": (C.c_int, [vp, C.c_uint64]),
        "crass_b200_cluster_block_dev": (C.c_int, [vp, vp, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(vp), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), vp]),
        "crass_b200_cluster_block_patterns_dev": (vp, [vp, vp, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), vp]),
        "crass_b200_sort_hits": (None, [vp, C.c_uint32]),
        "crass_b200_sort_hits_dev": (C.c_int, [vp, vp, C.c_uint32, vp, vp, C.c_uint32, vp, vp]),
        "crass_b200_token_block_bytes": (C.c_size_t, [C.c_uint32, C.c_uint32]),
        "crass_b200_unique_tokens_block_dev": (C.c_int, [vp, vp, C.c_uint32, vp, C.c_uint32, vp, C.c_uint32, vp]),
        "crass_b200_merge_token_blocks_dev": (C.c_int, [vp, vp, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, vp, C.c_uint32, vp]),
        "crass_b200_dr_list_from_block": (vp, [vp, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
        "crass_b200_non_redundant_set": (vp, [cp, C.c_uint32]),
        "crass_b200_ac_build_from_dr_list": (C.c_int, [cp, C.c_uint32, C.POINTER(vp), C.POINTER(C.c_uint32)]),
        "crass_b200_non_redundant_patterns": (vp, [cp, C.c_uint32, C.POINTER(C.c_uint32)]),
        "crass_b200_non_redundant_patterns_from_block": (vp, [vp, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
        "crass_b200_ac_build_from_block": (C.c_int, [vp, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(vp), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]),
        "crass_b200_ac_build_from_pattern_list": (C.c_int, [cp, C.POINTER(vp), C.POINTER(C.c_uint32)]),
        "crass_b200_run_files": (C.c_int, [vp, C.POINTER(cp), C.c_uint32, C.POINTER(Params), C.c_int, C.POINTER(vp), C.POINTER(C.c_int)]),
        "crass_b200_free": (None, [vp]),
        "crass_b200_engine_create": (C.c_int, [C.POINTER(C.c_int), C.c_uint32, C.POINTER(vp)]),
        "crass_b200_engine_destroy": (None, [vp]),
        "crass_b200_engine_num_devices": (C.c_uint32, [vp]),
        "crass_b200_engine_uses_nccl": (C.c_int, [vp]),
        "crass_b200_engine_search_file": (C.c_int, [vp, cp, C.POINTER(Params), C.POINTER(vp), C.POINTER(vp), u32p, C.POINTER(vp), u32p]),
        "crass_b200_engine_exchange": (C.c_int, [vp, cp, C.c_uint32, C.POINTER(vp), u32p, u32p]),
        "crass_b200_engine_find_singletons": (C.c_int, [vp, cp, vp, C.c_int, C.POINTER(vp), C.POINTER(vp), u32p, C.POINTER(vp), u32p]),
        "crass_b200_engine_release_file": (None, [vp, cp]),
        "crass_b200_engine_run_files": (C.c_int, [vp, C.POINTER(cp), C.c_uint32, C.POINTER(Params), C.c_int, C.POINTER(vp), C.POINTER(C.c_int)]),
        "crass_b200_run_files_multi": (C.c_int, [C.POINTER(C.c_int), C.c_uint32, C.POINTER(cp), C.c_uint32, C.POINTER(Params), C.c_int, C.POINTER(vp), C.POINTER(C.c_int)]),
        "crass_b200_engine_transfer_bytes": (None, [vp, u64p, u64p]),
        "crass_b200_engine_launch_count": (C.c_uint64, [vp]),
        "crass_b200_engine_stage_ms": (None, [vp, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)          # AttributeError here == the library does not export what the header declares
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


EXPORTED_SYMBOLS = None  # filled lazily by exported_symbols()


def _check(status):
    if status != 0:
        raise CrassB200Error(status, lib().crass_b200_last_error().decode("utf-8", "replace"))


def _take_str(ptr):
    if not ptr:
        return None
    s = C.string_at(ptr)
    lib().crass_b200_free(C.c_void_p(ptr))
    return s


def device_count():
    return lib().crass_b200_device_count()


def _np_ptr(a):
    return C.c_void_p(a.ctypes.data)


def pack_reads(reads):
    """list of bytes -> (bases uint8[n_bases], offsets uint64[n+1]) in the byte-packed batch layout."""
    offs = np.zeros(len(reads) + 1, dtype=np.uint64)
    if reads:
        offs[1:] = np.cumsum([len(r) for r in reads], dtype=np.uint64)
    bases = np.frombuffer(b"".join(reads), dtype=np.uint8).copy() if reads else np.zeros(0, dtype=np.uint8)
    return bases, offs


class Batch:
    """A parsed read set (crass_b200_batch): what kseq_read hands to searchFile."""

    def __init__(self, handle):
        self.h = handle

    @classmethod
    def from_file(cls, path):
        h = C.c_void_p()
        _check(lib().crass_b200_parse_file(path.encode(), C.byref(h)))
        return cls(h)

    @classmethod
    def stream_file(cls, path, range_bytes):
        """The file's records as a sequence of batches of about range_bytes of input each (the streamed feed)."""
        s = C.c_void_p()
        _check(lib().crass_b200_parse_stream_open(path.encode(), range_bytes, C.byref(s)))
        try:
            while True:
                h = C.c_void_p()
                got = lib().crass_b200_parse_stream_next(s, C.byref(h))
                if got < 0:
                    _check(got)
                if got == 0:
                    return
                yield cls(h)
        finally:
            lib().crass_b200_parse_stream_close(s)

    @classmethod
    def from_arrays(cls, bases, offsets, names=None):
        h = C.c_void_p()
        n = len(offsets) - 1
        arr = None
        if names is not None:
            arr = (C.c_char_p * n)(*[x if isinstance(x, bytes) else x.encode() for x in names])
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        _check(lib().crass_b200_batch_from_memory(_np_ptr(bases), _np_ptr(offsets), n, C.cast(arr, C.c_void_p) if arr is not None else None, C.byref(h)))
        return cls(h)

    def __del__(self):
        if getattr(self, "h", None):
            lib().crass_b200_batch_destroy(self.h)
            self.h = None

    def __len__(self):
        return lib().crass_b200_batch_num_reads(self.h)

    @property
    def max_read_len(self):
        return lib().crass_b200_batch_max_read_len(self.h)

    @property
    def parse_status(self):
        return lib().crass_b200_batch_parse_status(self.h)

    @property
    def offsets(self):
        n = len(self)
        p = lib().crass_b200_batch_offsets(self.h)
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint64)), shape=(n + 1,))

    @property
    def bases(self):
        n = int(self.offsets[-1])
        p = lib().crass_b200_batch_bases(self.h)
        if n == 0:
            return np.zeros(0, dtype=np.uint8)
        return np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(n,))

    def read(self, i):
        """bases of record i (no copy of the whole batch: a range of a streamed file keeps its reads in segments)"""
        ln = C.c_uint32(0)
        p = lib().crass_b200_batch_read(self.h, i, C.byref(ln))
        return C.string_at(p, ln.value) if p else b""

    def name(self, i):
        return lib().crass_b200_batch_name(self.h, i)

    def comment(self, i):
        has = C.c_int(0)
        s = lib().crass_b200_batch_comment(self.h, i, C.byref(has))
        return s if has.value else None

    def qual(self, i):
        has = C.c_int(0)
        s = lib().crass_b200_batch_qual(self.h, i, C.byref(has))
        return s if has.value else None

    def record_stream(self):
        """Same text as the checkers' kseq_dump (tests compare them)."""
        out = []
        offs, bases = self.offsets, self.bases
        for i in range(len(self)):
            c, q = self.comment(i), self.qual(i)
            seq = bases[int(offs[i]):int(offs[i + 1])].tobytes()
            out.append(b"\t".join([self.name(i), c if c is not None else b"\x01", seq, q if q is not None else b"\x01"]))
        out.append(b"#ret=%d" % self.parse_status)
        return b"\n".join(out) + b"\n"


class Automaton:
    def __init__(self, patterns):
        pats = [p if isinstance(p, bytes) else p.encode() for p in patterns]
        data = np.frombuffer(b"".join(pats), dtype=np.uint8).copy() if pats else np.zeros(1, dtype=np.uint8)
        offs = np.zeros(len(pats) + 1, dtype=np.uint32)
        if pats:
            offs[1:] = np.cumsum([len(p) for p in pats], dtype=np.uint32)
        self.h = C.c_void_p()
        _check(lib().crass_b200_ac_build(_np_ptr(data), _np_ptr(offs), len(pats), C.byref(self.h)))
        self.num_patterns = len(pats)

    @classmethod
    def from_dr_list(cls, drs, kmer_clust=6):
        """createNonRedundantSet + matcher build in one native call; `drs` in token order (list or '\\n'-joined bytes)."""
        text = drs if isinstance(drs, (bytes, bytearray)) else b"".join((d if isinstance(d, bytes) else d.encode()) + b"\n" for d in drs)
        self = cls.__new__(cls)
        self.h = C.c_void_p()
        n = C.c_uint32(0)
        _check(lib().crass_b200_ac_build_from_dr_list(bytes(text), kmer_clust, C.byref(self.h), C.byref(n)))
        self.num_patterns = n.value
        return self

    @classmethod
    def from_block(cls, block, cap, stride, kmer_clust=6):
        """createNonRedundantSet + matcher straight from a host copy of a token block -> (matcher or None, count, flags)."""
        self = cls.__new__(cls)
        self.h = C.c_void_p()
        n, cnt, fl = C.c_uint32(0), C.c_uint32(0), C.c_uint32(0)
        _check(lib().crass_b200_ac_build_from_block(C.c_void_p(_addr(block)), cap, stride, kmer_clust, C.byref(self.h), C.byref(cnt), C.byref(fl), C.byref(n)))
        self.num_patterns = n.value
        return (self if self.h else None), cnt.value, fl.value

    @classmethod
    def from_pattern_text(cls, text):
        """Matcher from a '\\n'-separated pattern set (what non_redundant_patterns returns)."""
        self = cls.__new__(cls)
        self.h = C.c_void_p()
        n = C.c_uint32(0)
        _check(lib().crass_b200_ac_build_from_pattern_list(bytes(text), C.byref(self.h), C.byref(n)))
        self.num_patterns = n.value
        return self

    def __del__(self):
        if getattr(self, "h", None):
            lib().crass_b200_ac_destroy(self.h)
            self.h = None

    @property
    def num_states(self):
        return lib().crass_b200_ac_num_states(self.h)

    @property
    def table_bytes(self):
        return lib().crass_b200_ac_table_bytes(self.h)

    def pattern_text(self):
        """The patterns the matcher was built from, '\\n'-separated in build order."""
        n = C.c_uint32(0)
        return _take_str(lib().crass_b200_ac_pattern_text(self.h, C.byref(n)))


class Results:
    """Mirror of the containers the reference fills (ReadMap / StringCheck / lookupTables)."""

    def __init__(self, handle=None):
        if handle is None:
            handle = C.c_void_p()
            _check(lib().crass_b200_results_create(C.byref(handle)))
        self.h = handle

    def __del__(self):
        if getattr(self, "h", None):
            lib().crass_b200_results_destroy(self.h)
            self.h = None

    def add_phase1(self, batch, hits, pool):
        _check(lib().crass_b200_results_add_phase1(self.h, batch.h, _np_ptr(hits), len(hits), _np_ptr(pool)))

    def add_phase2(self, batch, hits, pool):
        _check(lib().crass_b200_results_add_phase2(self.h, batch.h, _np_ptr(hits), len(hits), _np_ptr(pool)))

    @property
    def num_tokens(self):
        return lib().crass_b200_results_num_tokens(self.h)

    @property
    def num_reads(self):
        return lib().crass_b200_results_num_reads(self.h)

    def dr_list(self):
        s = _take_str(lib().crass_b200_results_dr_list(self.h))
        return [x for x in s.split(b"\n") if x]

    def adopt_tokens(self, all_drs):
        _check(lib().crass_b200_results_adopt_tokens(self.h, b"".join(d + b"\n" for d in all_drs)))

    def non_redundant(self, kmer_clust=6):
        n = C.c_uint32(0)
        s = _take_str(lib().crass_b200_results_non_redundant(self.h, kmer_clust, C.byref(n)))
        return [x for x in s.split(b"\n") if x]

    def dump(self, max_read_len):
        return _take_str(lib().crass_b200_results_dump(self.h, max_read_len)).decode("latin-1")


def _addr(x):
    """host address of a numpy array or a (pinned) CPU torch tensor"""
    return x.ctypes.data if isinstance(x, np.ndarray) else x.data_ptr()


def dr_list_from_hits(bases, offsets, hits, pool):
    n = len(offsets) - 1
    s = _take_str(lib().crass_b200_dr_list_from_hits(C.c_void_p(_addr(bases)), C.c_void_p(_addr(offsets)), n, _np_ptr(hits), len(hits), _np_ptr(pool)))
    if s is None:
        _check(-1)
    return [x for x in s.split(b"\n") if x]


def dr_list_from_tokens(records, stride, hits):
    """records: uint8 numpy array [n_hits*stride] copied back from the device; hits: unsorted HIT_DTYPE array."""
    s = _take_str(lib().crass_b200_dr_list_from_tokens(_np_ptr(records), stride, _np_ptr(hits), len(hits)))
    return [x for x in s.split(b"\n") if x]


def dr_list_from_unique(records, stride, first_read, raw=False):
    """records/first_read: the outputs of Context.unique_tokens_dev copied to numpy -> DRs in first-appearance order
    (a list, or with raw=True the '\\n'-terminated text the C-ABI hands out)."""
    s = _take_str(lib().crass_b200_dr_list_from_unique(_np_ptr(records), stride, _np_ptr(first_read), len(first_read)))
    return s if raw else [x for x in s.split(b"\n") if x]


def sort_hits(hits):
    """In-place read-order sort of a host copy of device hit records (HIT_DTYPE numpy array or pinned torch view)."""
    lib().crass_b200_sort_hits(C.c_void_p(_addr(hits)), len(hits))
    return hits


def token_block_bytes(cap, stride):
    return lib().crass_b200_token_block_bytes(cap, stride)


def dr_list_from_block(block, cap, stride):
    """Host copy of a token block (numpy / pinned torch uint8) -> (DR list text in first-appearance order, count, flags)."""
    n, fl = C.c_uint32(0), C.c_uint32(0)
    s = _take_str(lib().crass_b200_dr_list_from_block(C.c_void_p(_addr(block)), cap, stride, C.byref(n), C.byref(fl)))
    if s is None:
        _check(-1)
    return s, n.value, fl.value


def merge_dr_lists(drs):
    """First-appearance de-duplication; a list gives a list, '\\n'-terminated text gives text."""
    if isinstance(drs, (bytes, bytearray)):
        return _take_str(lib().crass_b200_merge_dr_lists(bytes(drs)))
    s = _take_str(lib().crass_b200_merge_dr_lists(b"".join(d + b"\n" for d in drs)))
    return [x for x in s.split(b"\n") if x]


def non_redundant_patterns(dr_text, kmer_clust=6):
    """createNonRedundantSet on a '\\n'-terminated DR list in token order -> '\\n'-terminated pattern text."""
    n = C.c_uint32(0)
    s = _take_str(lib().crass_b200_non_redundant_patterns(bytes(dr_text), kmer_clust, C.byref(n)))
    if s is None:
        _check(-1)
    return s


def non_redundant_patterns_from_block(block, cap, stride, kmer_clust=6):
    """Host copy of a token block -> (pattern text, count, flags) without the intermediate DR text."""
    n, cnt, fl = C.c_uint32(0), C.c_uint32(0), C.c_uint32(0)
    s = _take_str(lib().crass_b200_non_redundant_patterns_from_block(C.c_void_p(_addr(block)), cap, stride, kmer_clust, C.byref(cnt), C.byref(fl), C.byref(n)))
    if s is None:
        _check(-1)
    return s, cnt.value, fl.value


def non_redundant_list(drs, kmer_clust=6):
    """createNonRedundantSet on an ordered DR list -> pattern list (survivors + reverse complements)."""
    out = non_redundant_set(drs, kmer_clust)
    return [l[2:].encode() for l in out.split("\n") if l.startswith("P\t")]


def non_redundant_set(drs, kmer_clust=6):
    s = _take_str(lib().crass_b200_non_redundant_set(b"".join(d + b"\n" for d in drs), kmer_clust))
    return s.decode()


class Context:
    """One per GPU (crass_b200_ctx): stream, staging buffers, workspaces."""

    def __init__(self, device=0):
        self.h = C.c_void_p()
        _check(lib().crass_b200_ctx_create(device, C.byref(self.h)))
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            lib().crass_b200_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    @property
    def launch_count(self):
        return int(lib().crass_b200_ctx_launch_count(self.h))

    @property
    def last_candidates(self):
        return int(lib().crass_b200_ctx_last_candidates(self.h))

    def set_token_output(self, d_tokens, stride=64):
        """K4: let dr_search_dev also write the low-lexi DR token of every hit (torch uint8 tensor, stride bytes per hit slot)."""
        _check(lib().crass_b200_ctx_set_token_output(self.h, d_tokens.data_ptr() if d_tokens is not None else None, stride))

    def unique_tokens_dev(self, d_hits, n_hits, d_tokens, stride, d_out_tokens, d_out_first_read, d_out_count, stream=0):
        """K4b: device-side de-duplication of the token records of the first n_hits hit slots (torch tensors)."""
        _check(lib().crass_b200_unique_tokens_dev(self.h, d_hits.data_ptr(), n_hits, d_tokens.data_ptr(), stride, d_out_tokens.data_ptr(),
                                                  d_out_first_read.data_ptr(), d_out_count.data_ptr(), stream))

    def cluster_block_dev(self, d_block, cap, stride, kmer_clust=6, stream=0):
        """K5 + host passes + matcher build from a token block on the device -> (Automaton or None, count, flags)."""
        ac = Automaton.__new__(Automaton)
        ac.h = C.c_void_p()
        n, cnt, fl = C.c_uint32(0), C.c_uint32(0), C.c_uint32(0)
        _check(lib().crass_b200_cluster_block_dev(self.h, d_block.data_ptr(), cap, stride, kmer_clust, C.byref(ac.h), C.byref(cnt), C.byref(fl), C.byref(n), stream))
        ac.num_patterns = n.value
        return (ac if ac.h else None), cnt.value, fl.value

    def cluster_block_patterns_dev(self, d_block, cap, stride, kmer_clust=6, stream=0):
        """The same, returning the pattern set as '\\n'-terminated text -> (text, count, flags)."""
        n, cnt, fl = C.c_uint32(0), C.c_uint32(0), C.c_uint32(0)
        s = _take_str(lib().crass_b200_cluster_block_patterns_dev(self.h, d_block.data_ptr(), cap, stride, kmer_clust, C.byref(cnt), C.byref(fl), C.byref(n), stream))
        if s is None:
            _check(-1)
        return s, cnt.value, fl.value

    def keep_packed(self, on=True):
        """Let ac_scan_dev reuse the 2-bit stream dr_search_dev wrote for the same (unchanged) batch."""
        _check(lib().crass_b200_ctx_keep_packed(self.h, 1 if on else 0))

    def sort_hits_dev(self, d_found, n_reads, d_hits, d_counters, max_hits, d_sorted, stream=0):
        """Hit records of a *_dev search into read order on the device (d_counters: the launch's counter tensor)."""
        _check(lib().crass_b200_sort_hits_dev(self.h, d_found.data_ptr(), n_reads, d_hits.data_ptr(), d_counters.data_ptr(), max_hits,
                                              d_sorted.data_ptr(), stream))

    def unique_tokens_block_dev(self, d_hits, n_hits, d_tokens, stride, d_block, cap, stream=0):
        """K4b in block form (see include/crass_b200.h): distinct tokens of the hit list -> one token block."""
        _check(lib().crass_b200_unique_tokens_block_dev(self.h, d_hits.data_ptr(), n_hits, d_tokens.data_ptr(), stride, d_block.data_ptr(), cap, stream))

    # -- one process per GPU: the library's own NCCL communicator ------------------------------------
    @staticmethod
    def comm_unique_id():
        buf = (C.c_uint8 * 128)()
        _check(lib().crass_b200_comm_unique_id(buf))
        return bytes(buf)

    def comm_init(self, unique_id, rank, world):
        buf = (C.c_uint8 * 128).from_buffer_copy(unique_id)
        _check(lib().crass_b200_ctx_comm_init(self.h, buf, rank, world))

    @property
    def comm_world(self):
        return lib().crass_b200_ctx_comm_world(self.h)

    def exchange_tokens_dev(self, d_hits, n_hits, d_tokens, stride, d_send, cap, d_recv, shard_reads, d_merged, out_cap, stream=0):
        _check(lib().crass_b200_exchange_tokens_dev(self.h, d_hits.data_ptr(), n_hits, d_tokens.data_ptr(), stride, d_send.data_ptr(), cap,
                                                    d_recv.data_ptr(), shard_reads, d_merged.data_ptr(), out_cap, stream))

    def merge_token_blocks_dev(self, d_blocks, n_ranks, cap, stride, shard_reads, d_out_block, out_cap, stream=0):
        """K4c: the blocks of all ranks (rank order, back to back) -> one block with global first-appearance keys."""
        _check(lib().crass_b200_merge_token_blocks_dev(self.h, d_blocks.data_ptr(), n_ranks, cap, stride, shard_reads, d_out_block.data_ptr(), out_cap, stream))

    def last_dr_list(self):
        s = lib().crass_b200_ctx_last_dr_list(self.h)
        return [x for x in s.split(b"\n") if x]

    # -- host-buffer entry points (the reference-facing calls; copies inside) -----------------------
    def _collect(self, call, n_reads, want_found):
        found = np.zeros(n_reads, dtype=np.uint8) if want_found else None
        hp, pp = C.c_void_p(), C.c_void_p()
        nh, npool = C.c_uint32(0), C.c_uint32(0)
        _check(call(_np_ptr(found) if found is not None else None, C.byref(hp), C.byref(nh), C.byref(pp), C.byref(npool)))
        try:
            hits = np.frombuffer(C.string_at(hp, nh.value * 16), dtype=HIT_DTYPE).copy() if nh.value else np.zeros(0, dtype=HIT_DTYPE)
            pool = np.frombuffer(C.string_at(pp, npool.value * 4), dtype=np.uint32).copy() if npool.value else np.zeros(0, dtype=np.uint32)
        finally:
            lib().crass_b200_free(hp)
            lib().crass_b200_free(pp)
        return hits, pool, found

    def dr_search(self, bases, offsets, params=None, want_found=True):
        """searchCore over a host batch -> (hits sorted by read index, start/stop pool, found flags)."""
        params = params or Params()
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        n = len(offsets) - 1
        return self._collect(lambda f, hp, nh, pp, npl: lib().crass_b200_dr_search(
            self.h, _np_ptr(bases), _np_ptr(offsets), n, C.byref(params), f, hp, nh, pp, npl), n, want_found)

    def ac_upload(self, ac):
        _check(lib().crass_b200_ac_upload(self.h, ac.h))

    def upload(self, bases, offsets):
        """H2D of a host batch into the context (pinned sources copy at full PCIe rate)."""
        n = len(offsets) - 1
        _check(lib().crass_b200_batch_upload(self.h, C.c_void_p(_addr(bases)), C.c_void_p(_addr(offsets)), n))
        self._resident_n = n

    def dr_search_resident(self, params=None, want_found=False):
        params = params or Params()
        return self._collect(lambda f, hp, nh, pp, npl: lib().crass_b200_dr_search_resident(self.h, C.byref(params), f, hp, nh, pp, npl),
                             self._resident_n, want_found)

    def ac_scan_resident(self, ac, skip_found=True, want_found=False):
        return self._collect(lambda f, hp, nh, pp, npl: lib().crass_b200_ac_scan_resident(self.h, ac.h, 1 if skip_found else 0, f, hp, nh, pp, npl),
                             self._resident_n, want_found)

    def ac_scan(self, ac, bases, offsets, skip=None, want_found=True):
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        n = len(offsets) - 1
        if skip is not None:
            skip = np.ascontiguousarray(skip, dtype=np.uint8)
        return self._collect(lambda f, hp, nh, pp, npl: lib().crass_b200_ac_scan(
            self.h, ac.h, _np_ptr(bases), _np_ptr(offsets), n, _np_ptr(skip) if skip is not None else None, f, hp, nh, pp, npl), n, want_found)

    def edit_distance_batch(self, pairs):
        blob = bytearray()
        a_off, a_len, b_off, b_len = [], [], [], []
        for a, b in pairs:
            a_off.append(len(blob)); a_len.append(len(a)); blob += a
            b_off.append(len(blob)); b_len.append(len(b)); blob += b
        blob += b"\0"
        data = np.frombuffer(bytes(blob), dtype=np.uint8).copy()
        arrs = [np.asarray(x, dtype=np.uint32) for x in (a_off, a_len, b_off, b_len)]
        dist = np.zeros(len(pairs), dtype=np.int32)
        sim = np.zeros(len(pairs), dtype=np.float32)
        _check(lib().crass_b200_edit_distance_batch(self.h, _np_ptr(data), len(data), *[_np_ptr(x) for x in arrs], len(pairs), _np_ptr(dist), _np_ptr(sim)))
        return dist, sim

    @staticmethod
    def pack_uss_jobs(jobs):
        """jobs: (read index, start/stop list, front offset, DR index) -> (USS_JOB_DTYPE array, ss_in, ss_out capacity)"""
        rec = np.zeros(len(jobs), dtype=USS_JOB_DTYPE)
        ss_in = []
        out_off = 0
        for i, (read, ss, front, dr) in enumerate(jobs):
            rec[i] = (read, len(ss_in), len(ss), front, dr, out_off)
            ss_in.extend(ss)
            out_off += len(ss) + 4
        return rec, np.asarray(ss_in if ss_in else [0], dtype=np.uint32), max(out_off, 1)

    def update_start_stops(self, bases, offsets, drs, jobs, low_spacer=26):
        """K6, host buffers: ReadHolder::updateStartStops for every job -> [(status, new start/stop list)]."""
        rec, ss_in, cap = self.pack_uss_jobs(jobs)
        dr_bytes, dr_offs = pack_reads(drs if drs else [b""])
        dr_offs = dr_offs.astype(np.uint32)
        bases = np.ascontiguousarray(bases, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        if not jobs:
            return []
        ss_out = np.zeros(cap, dtype=np.uint32)
        n_out = np.zeros(max(len(jobs), 1), dtype=np.uint32)
        status = np.zeros(max(len(jobs), 1), dtype=np.uint8)
        _check(lib().crass_b200_update_start_stops(self.h, _np_ptr(bases), _np_ptr(offsets), len(offsets) - 1, _np_ptr(dr_bytes),
                                                   _np_ptr(dr_offs), len(drs), _np_ptr(rec), len(jobs), _np_ptr(ss_in), len(ss_in),
                                                   low_spacer, _np_ptr(ss_out), cap, _np_ptr(n_out), _np_ptr(status)))
        return [(int(status[i]), ss_out[rec["out_offset"][i]: rec["out_offset"][i] + n_out[i]].tolist()) for i in range(len(jobs))]

    def update_start_stops_dev(self, d_bases, d_offsets, d_dr_bytes, d_dr_offsets, d_jobs, n_jobs, d_ss_in, low_spacer, d_ss_out, d_n_out, d_status, stream=0):
        _check(lib().crass_b200_update_start_stops_dev(self.h, d_bases.data_ptr(), d_offsets.data_ptr(), d_dr_bytes.data_ptr(), d_dr_offsets.data_ptr(),
                                                       d_jobs.data_ptr(), n_jobs, d_ss_in.data_ptr(), low_spacer, d_ss_out.data_ptr(),
                                                       d_n_out.data_ptr(), d_status.data_ptr(), stream))

    # -- K7: consensus DR of DR groups ------------------------------------------------------------
    def ksw_align(self, pairs, xtra=0x80000 | 0x40000 | 5):
        """pairs: [(query letters, target letters, reverse-complement the query?)] -> [(score, te, qe, score2, te2, tb, qb)]"""
        pool = bytearray()
        jobs = np.zeros((len(pairs), 6), dtype=np.uint32)
        for i, (q, t, rc) in enumerate(pairs):
            jobs[i] = (len(pool), len(q), len(pool) + len(q), len(t), 1 if rc else 0, xtra)
            pool += q + t
        poolb = np.frombuffer(bytes(pool) or b"\0", dtype=np.uint8).copy()
        out = np.zeros((len(pairs), 8), dtype=np.int32)
        _check(lib().crass_b200_ksw_align(self.h, _np_ptr(poolb), len(pool), _np_ptr(jobs), len(pairs), _np_ptr(out)))
        assert not out[:, 7].any(), "a job was not taken (query longer than 128?)"
        return [tuple(int(x) for x in r[:7]) for r in out]

    def consensus_groups(self, cases):
        """cases: [dict(reads=[(seq, ss, dr index inside the group)], drs=[master, slaves...], array_len)] (one array_len for all)
        -> [dict(place, reversed, flags, zone, consensus, conservation (uint32 bit patterns), coverage)], status bits"""
        n = cases[0]["array_len"]
        assert all(c["array_len"] == n for c in cases)
        reads, drs, gfirst = [], [], [0]
        for c in cases:
            base = len(drs)
            drs += c["drs"]
            reads += [(r[0], r[1], base + r[2]) for r in c["reads"]]
            gfirst.append(len(drs))
        bases = np.frombuffer(b"".join(r[0] for r in reads) or b"\0", dtype=np.uint8).copy()
        offs = np.zeros(len(reads) + 1, dtype=np.uint64)
        offs[1:] = np.cumsum([len(r[0]) for r in reads])
        ss_offs = np.zeros(len(reads) + 1, dtype=np.uint32)
        ss_offs[1:] = np.cumsum([len(r[1]) for r in reads])
        pool = np.array([x for r in reads for x in r[1]] or [0], dtype=np.uint32)
        read_dr = np.array([r[2] for r in reads] or [0], dtype=np.uint32)
        dr_bytes = np.frombuffer(b"".join(drs), dtype=np.uint8).copy()
        dr_offs = np.zeros(len(drs) + 1, dtype=np.uint32)
        dr_offs[1:] = np.cumsum([len(d) for d in drs])
        gf = np.array(gfirst, dtype=np.uint32)
        G = len(cases)
        place = np.zeros(len(drs), dtype=np.int32)
        flags = np.zeros(len(drs), dtype=np.uint8)
        zone = np.zeros(2 * G, dtype=np.int32)
        cons = np.zeros(G * n, dtype=np.uint8)
        conserv = np.zeros(G * n, dtype=np.float32)
        cov = np.zeros(G * 4 * n, dtype=np.int32)
        status = C.c_uint32(0)
        _check(lib().crass_b200_consensus_groups(self.h, _np_ptr(bases), _np_ptr(offs), len(reads), _np_ptr(read_dr), _np_ptr(ss_offs), _np_ptr(pool),
                                                 _np_ptr(dr_bytes), _np_ptr(dr_offs), len(drs), _np_ptr(gf), G, n, _np_ptr(place), _np_ptr(flags),
                                                 _np_ptr(zone), _np_ptr(cons), _np_ptr(conserv), _np_ptr(cov), C.byref(status)))
        out = []
        for g in range(G):
            a, b = gfirst[g], gfirst[g + 1]
            out.append(dict(place=place[a:b].tolist(), reversed=[int(f) & 1 for f in flags[a:b]], flags=flags[a:b].tolist(),
                            zone=zone[2 * g:2 * g + 2].tolist(), consensus=cons[g * n:(g + 1) * n].tobytes(),
                            conservation=conserv[g * n:(g + 1) * n].view(np.uint32).tolist(), coverage=cov[g * 4 * n:(g + 1) * 4 * n].tolist()))
        return out, status.value

    def scan_right(self, seq, ss, pattern, min_spacer, scan_range=24):
        cap = 2 * (len(seq) // 4 + 8)
        arr = (C.c_uint32 * cap)(*ss)
        n = C.c_uint32(len(ss))
        _check(lib().crass_b200_scan_right(self.h, seq, len(seq), arr, C.byref(n), cap, pattern, len(pattern), min_spacer, scan_range))
        return list(arr[: n.value])

    def extend_pre_repeat(self, seq, ss, window, min_spacer):
        arr = (C.c_uint32 * len(ss))(*ss)
        rl = C.c_uint32(0)
        _check(lib().crass_b200_extend_pre_repeat(self.h, seq, len(seq), arr, len(ss), window, min_spacer, C.byref(rl)))
        return rl.value, list(arr)

    def run_files(self, paths, params=None, phases=2):
        """searchFile* -> createNonRedundantSet -> findSingletons*; returns (Results, max_read_len)."""
        params = params or Params()
        arr = (C.c_char_p * len(paths))(*[p.encode() for p in paths])
        out = C.c_void_p()
        ml = C.c_int(0)
        _check(lib().crass_b200_run_files(self.h, arr, len(paths), C.byref(params), phases, C.byref(out), C.byref(ml)))
        return Results(out), ml.value

    # -- device-resident entry points (torch tensors; only enqueue) ------------------------------------
    def dr_search_dev(self, d_bases, d_offsets, n_reads, max_read_len, params, d_found, d_hits, d_pool, d_counters, stream=0):
        _check(lib().crass_b200_dr_search_dev(self.h, d_bases.data_ptr(), d_offsets.data_ptr(), n_reads, max_read_len, C.byref(params),
                                              d_found.data_ptr() if d_found is not None else None, d_hits.data_ptr(), d_hits.numel() // 4,
                                              d_pool.data_ptr(), d_pool.numel(), d_counters.data_ptr(), stream))

    def ac_scan_dev(self, ac, d_bases, d_offsets, n_reads, max_read_len, d_skip, d_found, d_hits, d_pool, d_counters, stream=0):
        _check(lib().crass_b200_ac_scan_dev(self.h, ac.h, d_bases.data_ptr(), d_offsets.data_ptr(), n_reads, max_read_len,
                                            d_skip.data_ptr() if d_skip is not None else None,
                                            d_found.data_ptr() if d_found is not None else None, d_hits.data_ptr(), d_hits.numel() // 4,
                                            d_pool.data_ptr(), d_pool.numel(), d_counters.data_ptr(), stream))


class Engine:
    """The whole path on one or more GPUs of one box behind the C-ABI (crass_b200_engine_*): one caller, one set of
    containers; reads are sharded contiguously over `devices` (a device may be named more than once)."""

    def __init__(self, devices=(0,)):
        devs = (C.c_int * len(devices))(*devices)
        h = C.c_void_p()
        _check(lib().crass_b200_engine_create(devs, len(devices), C.byref(h)))
        self.h = h

    def close(self):
        if self.h:
            lib().crass_b200_engine_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def num_devices(self):
        return int(lib().crass_b200_engine_num_devices(self.h))

    @property
    def uses_nccl(self):
        return bool(lib().crass_b200_engine_uses_nccl(self.h))

    @property
    def launch_count(self):
        return int(lib().crass_b200_engine_launch_count(self.h))

    def transfer_bytes(self):
        a, b = C.c_uint64(0), C.c_uint64(0)
        lib().crass_b200_engine_transfer_bytes(self.h, C.byref(a), C.byref(b))
        return a.value, b.value

    def stage_ms(self):
        v = [C.c_double(0) for _ in range(4)]
        lib().crass_b200_engine_stage_ms(self.h, *[C.byref(x) for x in v])
        return dict(zip(("parse", "phase1_h2d_k1_d2h", "exchange_cluster", "phase2"), (x.value for x in v)))

    def run_files(self, paths, params=None, phases=2):
        """searchFile* -> createNonRedundantSet -> findSingletons* on all devices; returns (Results, max_read_len)."""
        params = params or Params()
        arr = (C.c_char_p * len(paths))(*[p.encode() for p in paths])
        out = C.c_void_p()
        ml = C.c_int(0)
        _check(lib().crass_b200_engine_run_files(self.h, arr, len(paths), C.byref(params), phases, C.byref(out), C.byref(ml)))
        return Results(out), ml.value

    def _take_hits(self, hp, nh, pp, npool):
        hits = np.empty(nh.value, dtype=HIT_DTYPE)
        pool = np.empty(npool.value, dtype=np.uint32)
        if nh.value:
            C.memmove(hits.ctypes.data, hp.value, hits.nbytes)
        if npool.value:
            C.memmove(pool.ctypes.data, pp.value, pool.nbytes)
        lib().crass_b200_free(hp)
        lib().crass_b200_free(pp)
        return hits, pool

    def search_file(self, path, params=None):
        """searchFile on all devices -> (hits in global read order, ss_pool); the parsed file stays with the engine."""
        params = params or Params()
        b, hp, pp = C.c_void_p(), C.c_void_p(), C.c_void_p()
        nh, npool = C.c_uint32(0), C.c_uint32(0)
        _check(lib().crass_b200_engine_search_file(self.h, path.encode(), C.byref(params), C.byref(b), C.byref(hp), C.byref(nh), C.byref(pp), C.byref(npool)))
        return self._take_hits(hp, nh, pp, npool)

    def exchange(self, path, kmer_clust=6):
        """token blocks -> all-gather -> merge -> createNonRedundantSet -> (Automaton or None, n_variants, n_patterns)"""
        ac = C.c_void_p()
        nv, npat = C.c_uint32(0), C.c_uint32(0)
        _check(lib().crass_b200_engine_exchange(self.h, path.encode(), kmer_clust, C.byref(ac), C.byref(nv), C.byref(npat)))
        a = None
        if ac.value:
            a = Automaton.__new__(Automaton)
            a.h = ac
            a.num_patterns = npat.value
        return a, nv.value, npat.value

    def find_singletons(self, path, ac, skip_found=True):
        b, hp, pp = C.c_void_p(), C.c_void_p(), C.c_void_p()
        nh, npool = C.c_uint32(0), C.c_uint32(0)
        _check(lib().crass_b200_engine_find_singletons(self.h, path.encode(), ac.h, 1 if skip_found else 0, C.byref(b), C.byref(hp), C.byref(nh), C.byref(pp), C.byref(npool)))
        return self._take_hits(hp, nh, pp, npool)

    def release_file(self, path):
        lib().crass_b200_engine_release_file(self.h, path.encode())


def run_files_multi(devices, paths, params=None, phases=2):
    """crass_b200_run_files_multi: a one-shot engine over `devices`; returns (Results, max_read_len)."""
    params = params or Params()
    devs = (C.c_int * len(devices))(*devices)
    arr = (C.c_char_p * len(paths))(*[p.encode() for p in paths])
    out = C.c_void_p()
    ml = C.c_int(0)
    _check(lib().crass_b200_run_files_multi(devs, len(devices), arr, len(paths), C.byref(params), phases, C.byref(out), C.byref(ml)))
    return Results(out), ml.value
