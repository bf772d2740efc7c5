"""Generates the golden fixtures in this directory from the UNMODIFIED reference
(oracle/_ref/libshasta_ref.so, built from /root/reference/src by oracle/Makefile).

Run in the build container only (needs /root/reference):  python tests/golden/make_golden.py

Outputs
  tinytest_markers.npz   markers of /root/reference/tests/TinyTest.fasta.gz produced by the reference's
                         ReadLoader(RLE) + MarkerFinder with k=10, p=0.1, seed 231 (toc, data, flags)
  lowhash_golden.npz     reference LowHash0 outputs (candidates, stats, per-iteration summary) for the
                         cases in LOWHASH_CASES; synthetic inputs are regenerated from their SynthParams.
"""
import gzip
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import bindings as B  # noqa: E402
from shasta_b200 import synth  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

# name -> (input spec, LowHash parameters)
LOWHASH_CASES = {
    # SURVEY.md Appendix D pin: 186 candidates, digest 0x3fc2c96e354f8733
    "tiny_default": ("tinytest", dict(m=4, hashFraction=0.01, minHashIterationCount=10, minBucketSize=0, maxBucketSize=10, minFrequency=2)),
    # conf/Nanopore-Dec2019.conf / Nanopore-May2022.conf [MinHash] values on the tiny input
    "tiny_dec2019": ("tinytest", dict(m=4, hashFraction=0.01, minHashIterationCount=10, minBucketSize=5, maxBucketSize=30, minFrequency=5)),
    # odd m (MurmurHash64A tail path), explicit bucket count, candidate-driven iteration count
    "tiny_m3_auto": ("tinytest", dict(m=3, hashFraction=0.02, minHashIterationCount=0, alignmentCandidatesPerRead=17.5, log2MinHashBucketCount=14, minBucketSize=2, maxBucketSize=8, minFrequency=3)),
    "tiny_m5": ("tinytest", dict(m=5, hashFraction=0.05, minHashIterationCount=4, minBucketSize=0, maxBucketSize=1000, minFrequency=1)),
    "synth600": (dict(reads=600, k=10, genome_markers=60000, n50_bases=20000, drop=0.10, ins=0.04, seed=7, palindromic_every=37),
                 dict(m=4, hashFraction=0.01, minHashIterationCount=10, minBucketSize=2, maxBucketSize=30, minFrequency=2)),
    "synth2000_k14": (dict(reads=2000, k=14, genome_markers=150000, n50_bases=15000, min_bases=8000, drop=0.08, ins=0.03, seed=11, repeat_period=3000, repeat_len=300),
                      dict(m=4, hashFraction=0.01, minHashIterationCount=10, minBucketSize=5, maxBucketSize=30, minFrequency=5)),
    "synth300_hifi": (dict(reads=300, k=14, genome_markers=30000, n50_bases=15000, min_bases=8000, drop=0.01, ins=0.005, seed=3),
                      dict(m=4, hashFraction=0.05, minHashIterationCount=20, minBucketSize=10, maxBucketSize=60, minFrequency=3)),
}


def load_input(spec):
    if spec == "tinytest":
        z = np.load(os.path.join(HERE, "tinytest_markers.npz"))
        return dict(toc=z["toc"], data=z["data"], flags=z["flags"], k=10)
    return synth.generate(synth.SynthParams(**spec))


def main():
    B.build()
    fasta = "/tmp/TinyTest.fasta"
    with open(fasta, "wb") as f:
        f.write(gzip.open("/root/reference/tests/TinyTest.fasta.gz").read())
    tt = B.ref_markers_from_fasta(fasta, k=10, probability=0.1, seed=231, min_read_length=10000)
    np.savez_compressed(os.path.join(HERE, "tinytest_markers.npz"), toc=tt["toc"], data=tt["data"], flags=tt["flags"])
    out = {}
    meta = {}
    for name, (spec, params) in LOWHASH_CASES.items():
        d = load_input(spec)
        p = B.LowHashParams(**params)
        c1, s1, it1, _ = B.ref_lowhash0(d["toc"], d["data"], d["flags"], p, threads=1)
        c4, s4, it4, _ = B.ref_lowhash0(d["toc"], d["data"], d["flags"], p, threads=4)
        assert np.array_equal(c1, c4) and np.array_equal(s1, s4) and np.array_equal(it1, it4), name
        out[name + "/candidates"] = c1
        out[name + "/stats"] = s1
        out[name + "/summary"] = it1
        meta[name] = dict(candidates=int(len(c1)), digest=hex(B.candidate_digest(c1)), iterations=int(len(it1)),
                          markers=int(d["toc"][-1]), reads=int(len(d["flags"])))
        print(name, meta[name])
    np.savez_compressed(os.path.join(HERE, "lowhash_golden.npz"), **out)
    with open(os.path.join(HERE, "lowhash_golden.json"), "w") as f:
        json.dump(meta, f, indent=1)


if __name__ == "__main__":
    main()
