"""Generates align_golden.npz from the UNMODIFIED reference alignment TUs
(oracle/_ref/libshasta_ref.so: Alignment.cpp, compressAlignment.cpp, Align4.cpp compiled where they lie; the absent SeqAn is
replaced by the shim that forwards to the oracle's overlap DP, so the DP tie-break stays 'parity unpinned' while everything
around it — the Align4 cell/component front end, its filters, AlignmentInfo and the compress codec — is the reference's own code).

Run in the build container only (needs /root/reference):  python tests/golden/make_align_golden.py

For every (case, pair): the two marker k-mer id sequences are regenerated from the case's SynthParams, and the fixture holds
  align4_*   the ordinals the reference's Align4 returns (concatenated, with a toc), only for pairs without a component tie
  info_*     the reference's AlignmentInfo::create words for those ordinals
  comp_*     the reference's shasta::compress bytes for those ordinals (concatenated, with a toc)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import bindings as B  # noqa: E402
from shasta_b200 import synth  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

LOWHASH = dict(m=4, hashFraction=0.02, minHashIterationCount=6, minBucketSize=2, maxBucketSize=30, minFrequency=2)

# name -> (SynthParams, pairs taken from the oracle's candidate list, Align4 options)
ALIGN_CASES = {
    "default_cells": (dict(reads=150, k=10, genome_markers=9000, n50_bases=9000, min_bases=5000, seed=5), 50,
                      dict(alignMethod=4, k=10, maxSkip=100, maxDrift=100, maxTrim=100, minAlignedMarkerCount=10, minAlignedFraction=0.1)),
    "small_cells_strict": (dict(reads=150, k=10, genome_markers=9000, n50_bases=9000, min_bases=5000, seed=5), 50,
                           dict(alignMethod=4, k=10, maxSkip=30, maxDrift=30, maxTrim=30, minAlignedMarkerCount=60, minAlignedFraction=0.4,
                                align4DeltaX=100, align4DeltaY=5, align4MinEntryCountPerCell=4, align4MaxDistanceFromBoundary=50, maxBand=300)),
    "noisy_k14": (dict(reads=120, k=14, genome_markers=8000, n50_bases=12000, min_bases=6000, seed=29, drop=0.12, ins=0.05), 40,
                  dict(alignMethod=4, k=14, maxSkip=100, maxDrift=100, maxTrim=100, minAlignedMarkerCount=10, minAlignedFraction=0.1)),
}


def case_pairs(spec, count):
    d = synth.generate(synth.SynthParams(**spec))
    cand, _, _ = B.oracle_lowhash0(d["toc"], d["data"], d["flags"], B.LowHashParams(**LOWHASH))
    toc = d["toc"].astype(np.int64)
    km = d["kmer"]
    pairs = []
    for r0, r1, same in cand[:count].tolist():
        o0, o1 = 2 * r0, 2 * r1 + (0 if same else 1)
        pairs.append((km[toc[o0]:toc[o0 + 1]], km[toc[o1]:toc[o1 + 1]]))
    return pairs


def main():
    B.build()
    assert B.have_ref(), "oracle/_ref is required"
    out = {}
    for name, (spec, count, opts) in ALIGN_CASES.items():
        o4 = B.make_align_options(**opts)
        ords, otoc, infos, comp, ctoc, used = [], [0], [], [], [0], []
        for p, (a, b) in enumerate(case_pairs(spec, count)):
            _, _, tie = B.oracle_align_pair(a, b, o4)
            if tie:         # the reference's pick among tied components depends on unordered_map iteration order
                continue
            ra = B.ref_align4(a, b, o4)
            used.append(p)
            ords.append(ra.reshape(-1, 2))
            otoc.append(otoc[-1] + len(ra))
            infos.append(B.ref_alignment_info(ra, len(a), len(b)) if len(ra) else np.zeros(12, np.uint32))
            cb = B.ref_compress(ra) if len(ra) else np.zeros(0, np.uint8)
            comp.append(cb)
            ctoc.append(ctoc[-1] + len(cb))
        out[name + "_pairs"] = np.array(used, np.uint32)
        out[name + "_align4"] = np.concatenate(ords).astype(np.uint32) if ords else np.zeros((0, 2), np.uint32)
        out[name + "_align4_toc"] = np.array(otoc, np.uint64)
        out[name + "_info"] = np.stack(infos).astype(np.uint32)
        out[name + "_comp"] = np.concatenate(comp).astype(np.uint8)
        out[name + "_comp_toc"] = np.array(ctoc, np.uint64)
        print(name, "pairs", len(used), "non-empty", sum(1 for x in ords if len(x)), "ordinals", otoc[-1], "compressed bytes", ctoc[-1])
    np.savez_compressed(os.path.join(HERE, "align_golden.npz"), **out)


if __name__ == "__main__":
    main()
