"""createReadGraph (ReadGraph.creationMethod 0) oracle against a direct statement of src/AssemblerReadGraph.cpp:35-175:
per read, the maxAlignmentCount largest (markerCount, alignmentId) pairs; edges and connectivity in the reference's order."""
import numpy as np

from oracle import bindings as B


def _records(rng, n, reads):
    rec = np.zeros((n, 16), np.uint32)
    a = rng.integers(0, reads, n)
    b = rng.integers(0, reads, n)
    same = a == b
    b[same] = (a[same] + 1) % reads
    rec[:, 0] = np.minimum(a, b)
    rec[:, 1] = np.maximum(a, b)
    rec[:, 2] = rng.integers(0, 2, n)
    rec[:, 9] = rng.integers(10, 14, n)         # few distinct marker counts: ties are decided by the alignment id
    rec[:, 15] = rng.integers(0, 2, n)          # stale flags must be overwritten
    return rec


def _direct(rec, reads, k):
    n = len(rec)
    keep = np.zeros(n, np.uint8)
    for r in range(reads):
        ids = [a for a in range(n) if rec[a, 0] == r or rec[a, 1] == r]
        best = sorted(((int(rec[a, 9]), a) for a in ids), reverse=True)[:k]
        for _, a in best:
            keep[a] = 1
    edges = []
    for a in range(n):
        if keep[a]:
            o0, o1 = 2 * int(rec[a, 0]), 2 * int(rec[a, 1]) + (0 if rec[a, 2] else 1)
            edges.append((o0, o1, a, 0))
            edges.append((o0 ^ 1, o1 ^ 1, a, 0))
    conn = [[] for _ in range(2 * reads)]
    for i, e in enumerate(edges):
        conn[e[0]].append(i)
        conn[e[1]].append(i)
    toc = np.zeros(2 * reads + 1, np.uint32)
    toc[1:] = np.cumsum([len(c) for c in conn])
    data = np.array([i for c in conn for i in c], np.uint32)
    return keep, np.array(edges, np.uint32).reshape(-1, 4), toc, data


def test_against_direct_statement():
    rng = np.random.default_rng(3)
    for n, reads, k in [(0, 5, 3), (1, 4, 6), (60, 12, 3), (300, 25, 6), (300, 25, 1), (200, 10, 0), (150, 8, 1000)]:
        rec = _records(rng, n, reads)
        out, keep, edges, toc, data = B.oracle_create_read_graph(rec, reads, k)
        dkeep, dedges, dtoc, ddata = _direct(rec, reads, k)
        assert np.array_equal(keep, dkeep), (n, reads, k)
        assert np.array_equal(edges, dedges) and np.array_equal(toc, dtoc) and np.array_equal(data, ddata)
        assert np.array_equal(out[:, 15] & 1, keep) and np.array_equal(out[:, :15], rec[:, :15])
        # an edge joins (r0, 0) with (r1, 0 / 1) and is ordered (src/AssemblerReadGraph.cpp:129,137)
        assert np.all(edges[:, 0] < edges[:, 1]) if len(edges) else True


def _quality_records(rng, n, reads):
    """AlignmentData with plausible AlignmentInfo words: Data{markerCount, firstOrdinal, lastOrdinal} x 2, markerCount, offsets,
    maxSkip, maxDrift."""
    rec = _records(rng, n, reads)
    for side in (0, 1):
        total = rng.integers(200, 5000, n)
        first = rng.integers(0, 150, n)
        last = total - 1 - rng.integers(0, 150, n)
        rec[:, 3 + 3 * side] = total
        rec[:, 4 + 3 * side] = first
        rec[:, 5 + 3 * side] = np.maximum(last, first)
    span = np.minimum(rec[:, 5] - rec[:, 4], rec[:, 8] - rec[:, 7]) + 1
    rec[:, 9] = np.maximum(1, (span * rng.uniform(0.2, 1.0, n)).astype(np.uint32))       # markerCount: up to ~4800 (beyond 3000)
    rec[:, 13] = rng.integers(0, 140, n)        # maxSkip: some beyond the histogram's 100
    rec[:, 14] = rng.integers(0, 120, n)        # maxDrift
    return rec


import pytest


@pytest.mark.skipif(not B.have_ref(), reason="reference build absent")
def test_histogram2_and_indicators_against_the_reference_classes():
    rng = np.random.default_rng(5)
    for (start, stop, bins) in [(0, 1, 100), (0, 3000, 300), (0, 100, 100)]:
        for trial in range(20):
            n = int(rng.integers(0, 400))
            # in range, exactly on the upper edge, and well beyond it (the dynamic-bounds growth path)
            x = rng.uniform(start, stop * rng.choice([0.5, 1.0, 1.7]), n)
            if n and trial % 3 == 0:
                x[rng.integers(0, n)] = stop
            x = np.round(x, 2) if stop > 1 else x
            for fraction in (0.015, 0.12, 0.5, 0.88, 0.985, 1.0):
                want = B.ref_histogram2_threshold(x, start, stop, bins, fraction)
                got = B.oracle_histogram2_threshold(x, start, stop, bins, fraction)
                assert want == got or (np.isnan(want) and np.isnan(got)), (start, stop, bins, trial, fraction)
    rec = _quality_records(rng, 200, 30)
    for r in rec:
        d0, d1 = r[3:6].astype(np.int64), r[6:9].astype(np.int64)
        frac = min(r[9] / (d0[2] + 1 - d0[1]), r[9] / (d1[2] + 1 - d1[1]))
        trim = max(min(d0[1], d1[1]), min(d0[0] - 1 - d0[2], d1[0] - 1 - d1[2]))
        assert np.array_equal(B.ref_alignment_indicators(r), np.array([frac, r[9], r[14], r[13], trim], np.float64))


def test_creation_method_2_against_direct_statement():
    rng = np.random.default_rng(8)
    pc = (0.015, 0.12, 0.12, 0.12, 0.015)      # ReadGraph.*Percentile defaults (src/AssemblerOptions.cpp)
    for n, reads, k in [(0, 4, 6), (400, 30, 6), (3000, 120, 6), (3000, 120, 2)]:
        rec = _quality_records(rng, n, reads)
        crit, out, keep, edges, toc, data = B.oracle_create_read_graph2(rec, reads, k, pc)
        ok = np.zeros(n, bool)
        for a in range(n):
            r = rec[a]
            d0, d1 = r[3:6].astype(np.int64), r[6:9].astype(np.int64)
            frac = min(r[9] / (d0[2] + 1 - d0[1]), r[9] / (d1[2] + 1 - d1[1]))
            trim = max(min(d0[1], d1[1]), min(d0[0] - 1 - d0[2], d1[0] - 1 - d1[2]))
            ok[a] = not (frac < crit["minAlignedFraction"] or r[9] < crit["minAlignedMarkerCount"] or r[14] > crit["maxDrift"]
                         or r[13] > crit["maxSkip"] or trim > crit["maxTrim"])
        if n:
            assert 0 < ok.sum() < n
        # the selection of method 0 restricted to the alignments that pass: compare through the sub-table
        sub = rec[ok]
        _, skeep, _, _, _ = B.oracle_create_read_graph(sub, reads, k)
        want = np.zeros(n, np.uint8)
        want[np.flatnonzero(ok)[skeep.astype(bool)]] = 1
        assert np.array_equal(keep, want)
        dkeep, dedges, dtoc, ddata = _direct(rec[:, :], reads, 0)      # edges / connectivity from the keep flags
        kept = np.flatnonzero(keep)
        assert np.array_equal(edges[0::2, 2], kept) and np.array_equal(edges[1::2, 2], kept)
        assert np.array_equal(out[:, 15] & 1, keep)
