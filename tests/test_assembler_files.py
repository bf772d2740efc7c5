"""The `shasta.Assembler`-shaped facade (shasta_b200/assembler.py) and the reference's Data/ file formats.

CPU part: files written by the facade's writer are opened by the reference's own MemoryMapped::Vector code
(oracle/_ref) with the right record types, and files written by the reference (Markers, ReadFlags, Kmers from
TinyTest through ReadLoader + MarkerFinder) are read back by the facade's reader.
GPU part: the two entry points run on a reference-written Data/ directory with the reference's Python-script call
sequence (scripts/FindAlignmentCandidatesLowHash0.py, scripts/ComputeAlignments.py) and reproduce the TinyTest pin."""
import gzip
import os

import numpy as np
import pytest

from oracle import bindings as B
from shasta_b200 import assembler as A


def fnv(raw):
    h = 1469598103934665603
    for b in bytes(raw):
        h = ((h ^ b) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h


@pytest.mark.skipif(not B.have_ref(), reason="reference build absent")
def test_writer_is_readable_by_the_reference(tmp_path):
    rng = np.random.default_rng(1)
    cases = [(12, rng.integers(0, 1000, (1000, 3)).astype(np.uint32)), (64, rng.integers(0, 2**31, (77, 16)).astype(np.uint32)),
             (24, rng.integers(0, 2**40, (50, 3)).astype(np.uint64)), (1, rng.integers(0, 255, 5000).astype(np.uint8)),
             (8, rng.integers(0, 2**50, 4097).astype(np.uint64)), (4, np.zeros(0, np.uint32))]
    for i, (size, arr) in enumerate(cases):
        path = str(tmp_path / f"v{i}")
        A.mm_write_vector(path, arr, object_size=size)
        n, h = B.ref_open_vector(path, size)
        assert n == arr.nbytes // size
        assert h == fnv(arr.tobytes())
        back = A.mm_read_vector(path, arr.dtype, object_size=size)
        assert np.array_equal(np.asarray(back).reshape(arr.shape), arr)
    # wrong record size is rejected by the reference ("unexpected object size")
    with pytest.raises(RuntimeError):
        B.ref_open_vector(str(tmp_path / "v0"), 64)


@pytest.fixture(scope="module")
def tiny_data_dir(tmp_path_factory):
    if not B.have_ref() or not os.path.exists("/root/reference/tests/TinyTest.fasta.gz"):
        pytest.skip("needs the reference tree")
    d = tmp_path_factory.mktemp("run")
    fasta = str(d / "TinyTest.fasta")
    with open(fasta, "wb") as f:
        f.write(gzip.open("/root/reference/tests/TinyTest.fasta.gz").read())
    os.makedirs(d / "Data")
    B.ref_write_data_dir(fasta, str(d / "Data") + "/", k=10)
    return str(d / "Data") + "/"


def test_reader_on_reference_written_files(tiny_data_dir, golden_dir):
    a = A.Assembler(largeDataFileNamePrefix=tiny_data_dir)
    a.accessKmers()
    assert a.k == 10
    a.accessMarkers()
    toc, data, flags = a._markers
    z = np.load(os.path.join(golden_dir, "tinytest_markers.npz"))
    assert np.array_equal(toc, z["toc"]) and np.array_equal(data, z["data"]) and np.array_equal(flags, z["flags"])
    with pytest.raises(RuntimeError, match="not accessible"):
        A.Assembler(largeDataFileNamePrefix=tiny_data_dir).checkMarkersAreOpen()


@pytest.mark.gpu
def test_script_sequence_on_reference_data_dir(tmp_path, golden_dir):
    # scripts/FindAlignmentCandidatesLowHash0.py + scripts/ComputeAlignments.py call sequence. The Data/ inputs are written
    # with the facade's writer from the golden TinyTest markers (the reference tree is not on the GPU box).
    z = np.load(os.path.join(golden_dir, "tinytest_markers.npz"))
    prefix = str(tmp_path / "Data") + "/"
    os.makedirs(prefix)
    A.mm_write_vector_of_vectors(prefix + "Markers", z["toc"], z["data"], data_object_size=7)
    A.mm_write_vector(prefix + "ReadFlags", z["flags"], object_size=1)
    A.mm_write_vector(prefix + "Kmers", np.zeros((1 << 20) * 24, np.uint8), object_size=24)
    a = A.Assembler(largeDataFileNamePrefix=prefix)
    a.accessKmers()
    a.accessMarkers()
    a.findAlignmentCandidatesLowHash0(m=4, hashFraction=0.01, minHashIterationCount=10, alignmentCandidatesPerRead=20.,
                                      minBucketSize=0, maxBucketSize=10, minFrequency=2, threadCount=0)
    b = A.Assembler(largeDataFileNamePrefix=prefix)
    b.accessKmers()
    b.accessMarkers()
    b.accessAlignmentCandidates()
    cands = b.getAlignmentCandidates()
    assert len(cands) == 186 and cands[0].readIds == [0, 2] and cands[0].isSameStrand is False
    assert B.candidate_digest(b._candidates) == 0x3fc2c96e354f8733
    opts = A.AlignOptions()
    opts.minAlignedMarkerCount = 100
    opts.minAlignedFraction = 0.4
    b.computeAlignments(opts, 0)
    c = A.Assembler(largeDataFileNamePrefix=prefix)
    c.accessAlignmentData()
    c.accessCompressedAlignments()
    rec = c._alignment_data
    toc, data = c._compressed
    assert len(rec) > 20 and len(toc) == len(rec) + 1
    # against the oracle
    kmer = np.ascontiguousarray(np.asarray(z["data"]).reshape(-1, 7)[:, :4]).view("<u4").reshape(-1)
    oo = B.make_align_options(alignMethod=3, k=10, maxSkip=30, maxDrift=30, maxTrim=30, minAlignedMarkerCount=100,
                              minAlignedFraction=0.4, downsamplingFactor=0.1, bandExtend=10, maxBand=1000)
    orec, otoc, odata, _ = B.oracle_compute_alignments(z["toc"], kmer, b._candidates, oo, threads=4)
    assert np.array_equal(rec, orec) and np.array_equal(toc, otoc) and np.array_equal(data, odata)
    table_toc = A.mm_read_vector(prefix + "AlignmentTable.toc", np.uint32, object_size=4)
    assert len(table_toc) == 2 * 20 + 1 and int(table_toc[-1]) == 4 * len(rec)
    with pytest.raises(RuntimeError, match="unreasonably small"):
        b.findAlignmentCandidatesLowHash0(m=4, hashFraction=0.01, minHashIterationCount=10, alignmentCandidatesPerRead=20.,
                                          minBucketSize=0, maxBucketSize=10, minFrequency=2, log2MinHashBucketCount=3)
