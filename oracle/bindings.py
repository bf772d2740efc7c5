"""TEST INFRASTRUCTURE — ctypes bindings to the CPU oracle (oracle/_build/liboracle.so) and,
when present, to the unmodified reference build (oracle/_ref/libshasta_ref.so).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module. The product (shasta_b200/) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(_HERE, "_build", "liboracle.so")
REF_SO = os.path.join(_HERE, "_ref", "libshasta_ref.so")


def build(quiet=True):
    """Compile the oracle (always) and the reference build (only where /root/reference exists)."""
    out = subprocess.DEVNULL if quiet else None
    subprocess.check_call(["make", "-C", _HERE, "oracle"], stdout=out)
    if os.path.isdir("/root/reference/src"):
        subprocess.check_call(["make", "-C", _HERE, "-j8", "ref"], stdout=out)


_oracle = None
_ref = None


def oracle_lib():
    global _oracle
    if _oracle is None:
        if not os.path.exists(ORACLE_SO):
            build()
        _oracle = C.CDLL(ORACLE_SO)
        _oracle.orc_murmurhash64a.restype = C.c_uint64
        _oracle.orc_murmurhash64a.argtypes = [C.c_void_p, C.c_int, C.c_uint64]
        _oracle.orc_murmurhash2.restype = C.c_uint32
        _oracle.orc_murmurhash2.argtypes = [C.c_void_p, C.c_int, C.c_uint32]
        _oracle.orc_reverse_complement_kmer.restype = C.c_uint32
        _oracle.orc_reverse_complement_kmer.argtypes = [C.c_uint32, C.c_uint32]
        _oracle.orc_kmer_downsampling_hash.restype = C.c_uint32
        _oracle.orc_kmer_downsampling_hash.argtypes = [C.c_uint32, C.c_uint32]
        _oracle.orc_lowhash0.restype = C.c_int
        _oracle.orc_lowhash0.argtypes = [
            C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p,
            C.c_uint64, C.c_double, C.c_uint64, C.c_double, C.c_uint64,
            C.c_uint64, C.c_uint64, C.c_uint64,
            C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p]
        _oracle.orc_free.argtypes = [C.c_void_p]
    return _oracle


def have_ref():
    return os.path.exists(REF_SO)


def ref_lib():
    global _ref
    if _ref is None:
        _ref = C.CDLL(REF_SO)
        _ref.ref_lowhash0.restype = C.c_int
        _ref.ref_lowhash0.argtypes = [
            C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p,
            C.c_uint64, C.c_double, C.c_uint64, C.c_double, C.c_uint64,
            C.c_uint64, C.c_uint64, C.c_uint64, C.c_uint64,
            C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
            C.c_void_p, C.c_uint64, C.c_void_p]
        _ref.ref_markers_from_fasta.restype = C.c_int
        _ref.ref_markers_from_fasta.argtypes = [
            C.c_char_p, C.c_uint64, C.c_double, C.c_int, C.c_uint64, C.c_uint64,
            C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        _ref.ref_murmurhash64a.restype = C.c_uint64
        _ref.ref_murmurhash64a.argtypes = [C.c_void_p, C.c_int, C.c_uint64]
        _ref.ref_murmurhash2.restype = C.c_uint32
        _ref.ref_murmurhash2.argtypes = [C.c_void_p, C.c_int, C.c_uint32]
        _ref.ref_free.argtypes = [C.c_void_p]
    return _ref


class LowHashParams(dict):
    """m, hashFraction, minHashIterationCount, alignmentCandidatesPerRead, log2MinHashBucketCount,
    minBucketSize, maxBucketSize, minFrequency  (src/Assembler.hpp:688-699)."""

    DEFAULTS = dict(m=4, hashFraction=0.01, minHashIterationCount=10, alignmentCandidatesPerRead=20.0,
                    log2MinHashBucketCount=0, minBucketSize=0, maxBucketSize=10, minFrequency=2)

    def __init__(self, **kw):
        d = dict(self.DEFAULTS)
        d.update(kw)
        super().__init__(d)


def _common(toc, data, flags):
    toc = np.ascontiguousarray(toc, dtype=np.uint64)
    data = np.ascontiguousarray(data, dtype=np.uint8)
    flags = np.ascontiguousarray(flags, dtype=np.uint8)
    R = (len(toc) - 1) // 2
    # 8 bytes of slack: the 7-byte records are read with 4-byte loads.
    pad = np.zeros(len(data) + 8, np.uint8)
    pad[:len(data)] = data
    return R, toc, pad, flags


def oracle_lowhash0(toc, data, flags, p: LowHashParams, max_iters=4096):
    """Returns (candidates uint32[n,3], stats uint64[R,3], iterSummary uint64[iters,2])."""
    lib = oracle_lib()
    R, toc, data, flags = _common(toc, data, flags)
    cand = C.c_void_p()
    n = C.c_uint64()
    iters = C.c_uint64()
    stats = np.zeros((R, 3), np.uint64)
    summ = np.zeros((max_iters, 2), np.uint64)
    rc = lib.orc_lowhash0(R, toc.ctypes.data, data.ctypes.data, flags.ctypes.data,
                          p["m"], p["hashFraction"], p["minHashIterationCount"],
                          p["alignmentCandidatesPerRead"], p["log2MinHashBucketCount"],
                          p["minBucketSize"], p["maxBucketSize"], p["minFrequency"],
                          C.byref(cand), C.byref(n), stats.ctypes.data, summ.ctypes.data, max_iters,
                          C.byref(iters))
    if rc != 0:
        raise RuntimeError("log2MinHashBucketCount is unreasonably small.")
    out = np.ctypeslib.as_array(C.cast(cand, C.POINTER(C.c_uint32)), (n.value, 3)).copy() if n.value else np.zeros((0, 3), np.uint32)
    lib.orc_free(cand)
    return out, stats, summ[:iters.value].copy()


def ref_lowhash0(toc, data, flags, p: LowHashParams, threads=0, max_iters=4096, quiet=True):
    """Same, through the unmodified reference LowHash0. Also returns the constructor's wall seconds."""
    lib = ref_lib()
    R, toc, data, flags = _common(toc, data, flags)
    cand = C.c_void_p()
    n = C.c_uint64()
    iters = C.c_uint64()
    sec = C.c_double()
    stats = np.zeros((R, 3), np.uint64)
    summ = np.zeros((max_iters, 2), np.uint64)
    rc = lib.ref_lowhash0(R, toc.ctypes.data, data.ctypes.data, flags.ctypes.data,
                          p["m"], p["hashFraction"], p["minHashIterationCount"],
                          p["alignmentCandidatesPerRead"], p["log2MinHashBucketCount"],
                          p["minBucketSize"], p["maxBucketSize"], p["minFrequency"], threads,
                          C.byref(cand), C.byref(n), stats.ctypes.data, C.byref(sec), 1 if quiet else 0,
                          summ.ctypes.data, max_iters, C.byref(iters))
    if rc != 0:
        raise RuntimeError("reference LowHash0 failed")
    out = np.ctypeslib.as_array(C.cast(cand, C.POINTER(C.c_uint32)), (n.value, 3)).copy() if n.value else np.zeros((0, 3), np.uint32)
    lib.ref_free(cand)
    return out, stats, summ[:iters.value].copy(), sec.value


def ref_markers_from_fasta(path, k=10, probability=0.1, seed=231, min_read_length=10000, threads=4, want_hash=False):
    lib = ref_lib()
    R = C.c_uint64()
    toc = C.c_void_p()
    data = C.c_void_p()
    flags = C.c_void_p()
    kh = C.c_void_p()
    rc = lib.ref_markers_from_fasta(path.encode(), k, probability, seed, min_read_length, threads,
                                    C.byref(R), C.byref(toc), C.byref(data), C.byref(flags),
                                    C.byref(kh) if want_hash else None)
    if rc != 0:
        raise RuntimeError("reference marker finding failed")
    R = R.value
    tocn = np.ctypeslib.as_array(C.cast(toc, C.POINTER(C.c_uint64)), (2 * R + 1,)).copy()
    M = int(tocn[-1])
    datan = np.ctypeslib.as_array(C.cast(data, C.POINTER(C.c_uint8)), (M * 7,)).copy()
    flagsn = np.ctypeslib.as_array(C.cast(flags, C.POINTER(C.c_uint8)), (R,)).copy()
    out = dict(toc=tocn, data=datan, flags=flagsn, k=k)
    if want_hash:
        out["kmerHash"] = np.ctypeslib.as_array(C.cast(kh, C.POINTER(C.c_uint32)), (1 << (2 * k),)).copy()
        lib.ref_free(kh)
    for ptr in (toc, data, flags):
        lib.ref_free(ptr)
    return out


def candidate_digest(c):
    """FNV-1a style digest over (readId0, readId1, isSameStrand) rows."""
    h = 1469598103934665603
    for a, b, s in np.asarray(c).tolist():
        for v in (a, b, s):
            h = ((h ^ v) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h
