"""Generates tinytest_reads.npz: the INPUTS of the reference's MarkerFinder for /root/reference/tests/TinyTest.fasta.gz as the
reference itself stores them (ReadLoader with the RLE read representation -> LongBaseSequences words + base counts, and
kmerTable[].isMarker for k = 10, p = 0.1, seed 231), through the unmodified reference TUs in oracle/_ref. The expected
OUTPUT of MarkerFinder on these inputs is tinytest_markers.npz (make_golden.py).

Run in the build container only (needs /root/reference):  python tests/golden/make_marker_golden.py
"""
import gzip
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import bindings as B  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    with tempfile.TemporaryDirectory() as tmp:
        fasta = os.path.join(tmp, "TinyTest.fasta")
        with open(fasta, "wb") as f:
            f.write(gzip.open("/root/reference/tests/TinyTest.fasta.gz").read())
        r = B.ref_reads_from_fasta(fasta, k=10, probability=0.1, seed=231, min_read_length=10000, threads=4)
        m = B.ref_markers_from_fasta(fasta, k=10, probability=0.1, seed=231, min_read_length=10000, threads=4)
    z = np.load(os.path.join(HERE, "tinytest_markers.npz"))
    assert np.array_equal(m["toc"], z["toc"]) and np.array_equal(m["data"], z["data"]), "tinytest_markers.npz is stale"
    bitmap = np.packbits(r["is_marker"].astype(np.uint8), bitorder="little").view(np.uint32)
    np.savez_compressed(os.path.join(HERE, "tinytest_reads.npz"), word_offsets=r["word_offsets"], words=r["words"],
                        base_counts=r["base_counts"], is_marker_bitmap=bitmap, k=np.array(10))
    print(f"{len(r['base_counts'])} reads, {int(r['base_counts'].sum())} RLE bases, {len(r['words'])} words, "
          f"{int(r['is_marker'].sum())} marker k-mers of {len(r['is_marker'])}")


if __name__ == "__main__":
    main()
