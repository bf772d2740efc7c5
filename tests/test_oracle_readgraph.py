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
