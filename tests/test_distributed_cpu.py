"""CPU test of the N>1 path (world_size 2, gloo): the sharded LowHash0 orchestration of shasta_b200/distributed.py —
bucket ownership, per-iteration all-to-all-v, final pair exchange by readId0 owner, statistics all-reduce, rank-order
concatenation — driven with a numpy stand-in for the per-rank device stages, checked against the CPU oracle on the
unsharded input. (On GPUs the same orchestration drives the C ABI: tests/test_gpu_distributed.py, bench.py --gpus N.)"""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

U64 = np.uint64
M64 = U64(0xc6a4a7935bd1e995)


def murmur64a_features(kmer, m, seed):
    """Vectorised MurmurHash64A (src/MurmurHash2.cpp:96-137) of every m-marker feature of one row."""
    n = len(kmer) - m + 1
    if n <= 0:
        return np.zeros(0, U64)
    k = kmer.astype(U64)
    with np.errstate(over="ignore"):
        h = np.full(n, U64(seed) ^ (U64(4 * m) * M64), U64)
        for b in range(m // 2):
            w = k[2 * b:2 * b + n] | (k[2 * b + 1:2 * b + 1 + n] << U64(32))
            w = w * M64
            w ^= w >> U64(47)
            w = w * M64
            h ^= w
            h = h * M64
        if m % 2:
            h ^= k[m - 1:m - 1 + n]
            h = h * M64
        h ^= h >> U64(47)
        h = h * M64
        h ^= h >> U64(47)
    return h


class NumpyStages:
    """Test double with the interface of distributed.CudaStages."""
    max_fused_iterations = 4

    def __init__(self, toc, kmer, flags, read_begin, read_end, read_count_total, total_markers):
        self.toc, self.kmer, self.flags = toc.astype(np.int64), kmer, flags
        self.rb, self.re, self.R, self.M = read_begin, read_end, read_count_total, total_markers

    def begin(self, p):
        self.p = p
        est = int(p["hashFraction"] * float(self.M))
        log2est = est.bit_length()
        b = p.get("log2MinHashBucketCount", 0) or (5 + log2est)
        self.log2b = min(b, 31)
        self.mask = U64((1 << self.log2b) - 1)
        self.threshold = U64(int(np.float64(p["hashFraction"]) * np.float64(np.iinfo(np.uint64).max)))
        self.stats = np.zeros(3 * self.R, np.int64)
        self.acc = {}
        return self.log2b

    def sweep(self, it0, count):
        self.slabs = []
        for it in range(it0, it0 + count):
            keys, vals = [], []
            for r in range(self.rb, self.re):
                if self.flags[r] & 1:
                    continue
                for s in (0, 1):
                    o = 2 * (r - self.rb) + s
                    row = self.kmer[self.toc[o]:self.toc[o + 1]]
                    h = murmur64a_features(row, self.p["m"], 37 * it)
                    h = h[h < self.threshold]
                    keys.append(((h & self.mask) << U64(32)) | (h >> U64(32)))
                    vals.append(np.full(len(h), 2 * r + s, np.uint32))
            self.slabs.append((np.concatenate(keys) if keys else np.zeros(0, U64), np.concatenate(vals) if vals else np.zeros(0, np.uint32)))
        return [len(k) for k, _ in self.slabs]

    def slab(self, s, n):
        k, v = self.slabs[s]
        return torch.from_numpy(k.view(np.int64).copy()), torch.from_numpy(v.view(np.int32).copy())

    def partition(self, keys, vals, shift, bits):
        k = keys.numpy().view(U64)
        d = ((k >> U64(shift)) & U64((1 << bits) - 1)).astype(np.int64)
        order = np.argsort(d, kind="stable")
        counts = np.bincount(d, minlength=1 << bits).tolist()
        return keys[torch.from_numpy(order)], vals[torch.from_numpy(order)], counts

    def process_entries(self, keys, vals):
        k = keys.numpy().view(U64)
        v = vals.numpy().view(np.uint32)
        order = np.argsort(k >> U64(32), kind="stable")
        k, v = k[order], v[order]
        bucket = (k >> U64(32)).astype(np.int64)
        lo = max(2, self.p["minBucketSize"])
        i = 0
        while i < len(k):
            j = i
            while j < len(k) and bucket[j] == bucket[i]:
                j += 1
            size = j - i
            cls = 0 if size < self.p["minBucketSize"] else (2 if size > self.p["maxBucketSize"] else 1)
            for t in range(i, j):
                self.stats[3 * (int(v[t]) >> 1) + cls] += 1
            if lo <= size <= self.p["maxBucketSize"]:
                for a in range(i, j):
                    for b in range(i, j):
                        if (int(k[a]) & 0xffffffff) != (int(k[b]) & 0xffffffff):
                            continue
                        r0, r1 = int(v[a]) >> 1, int(v[b]) >> 1
                        if r1 <= r0:
                            continue
                        key = (r0 << 32) | (r1 << 1) | ((int(v[a]) ^ int(v[b])) & 1)
                        self.acc[key] = self.acc.get(key, 0) + 1
            i = j

    def local_pairs(self):
        keys = np.array(sorted(self.acc), dtype=U64)
        cnts = np.array([self.acc[int(x)] for x in keys], dtype=np.uint32)
        return torch.from_numpy(keys.view(np.int64).copy()), torch.from_numpy(cnts.view(np.int32).copy())

    def set_pairs(self, keys, counts):
        self.acc = {}
        for kk, cc in zip(keys.numpy().view(U64).tolist(), counts.numpy().view(np.uint32).tolist()):
            self.acc[kk] = self.acc.get(kk, 0) + cc

    def emit(self):
        out = [(key >> 32, (key & 0xffffffff) >> 1, 0 if (key & 1) else 1) for key in sorted(self.acc)
               if (self.acc[key] & 0xffff) >= self.p["minFrequency"]]
        return np.array(out, np.uint32).reshape(-1, 3)

    def stats_tensor(self):
        return torch.from_numpy(self.stats)


def _worker(rank, world, port, case, result_path):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import bindings as B
        from shasta_b200 import distributed as D
        from shasta_b200 import synth
        d = synth.generate(synth.SynthParams(**case["synth"]))
        toc = d["toc"].astype(np.int64)
        R = len(d["flags"])
        weights = toc[2::2] - toc[0:-2:2]
        bounds = D.balanced_read_ranges(weights, world)
        rb, re = bounds[rank], bounds[rank + 1]
        local_toc = toc[2 * rb:2 * re + 1] - toc[2 * rb]
        local_kmer = d["kmer"][toc[2 * rb]:toc[2 * re]]
        stages = NumpyStages(local_toc, local_kmer, d["flags"], rb, re, R, int(toc[-1]))
        cand, stats, info = D.lowhash0_sharded(stages, case["params"], R)
        # the ranks own disjoint, increasing readId0 ranges (cut for equal pair mass, not equal width)
        spans = [None] * world
        dist.all_gather_object(spans, (int(cand[:, 0].min()), int(cand[:, 0].max())) if len(cand) else None)
        spans = [x for x in spans if x is not None]
        assert all(spans[g][1] < spans[g + 1][0] for g in range(len(spans) - 1)), spans
        allc = D.gather_candidates(cand)
        # rebalancing keeps the global order and evens the slice sizes
        rbc = D.rebalance_candidates(cand)
        allr = D.gather_candidates(rbc)
        sizes = [None] * world
        dist.all_gather_object(sizes, len(rbc))
        assert max(sizes) - min(sizes) <= 1
        if rank == 0:
            assert np.array_equal(allr, allc)
            oc, os_, _ = B.oracle_lowhash0(d["toc"], d["data"], d["flags"], B.LowHashParams(**case["params"]))
            ok = np.array_equal(allc, oc) and np.array_equal(stats.numpy().reshape(-1, 3).astype(np.uint64), os_)
            with open(result_path, "w") as f:
                f.write("OK %d %d" % (len(oc), info["entriesReceived"]) if ok else "MISMATCH %d %d" % (len(allc), len(oc)))
    finally:
        dist.destroy_process_group()


CASES = [
    dict(synth=dict(reads=90, k=10, genome_markers=6000, n50_bases=9000, min_bases=4000, seed=17, palindromic_every=11),
         params=dict(m=4, hashFraction=0.02, minHashIterationCount=6, alignmentCandidatesPerRead=20.0, log2MinHashBucketCount=0,
                     minBucketSize=2, maxBucketSize=20, minFrequency=2)),
    dict(synth=dict(reads=70, k=10, genome_markers=5000, n50_bases=9000, min_bases=4000, seed=23),
         params=dict(m=3, hashFraction=0.03, minHashIterationCount=5, alignmentCandidatesPerRead=20.0, log2MinHashBucketCount=12,
                     minBucketSize=0, maxBucketSize=8, minFrequency=1)),
]


@pytest.mark.parametrize("case_index", range(len(CASES)))
def test_sharded_lowhash_world2_gloo(case_index, tmp_path):
    _run_world(2, case_index, tmp_path)


def test_sharded_lowhash_world4_gloo(tmp_path):
    # four ranks: two owner bits for the buckets, pair owners cut on the fine histogram into four unequal readId0 ranges
    _run_world(4, 0, tmp_path)


def _run_world(world, case_index, tmp_path):
    result = tmp_path / "result.txt"
    port = 29600 + 10 * world + case_index + (os.getpid() % 200)
    mp.spawn(_worker, args=(world, port, CASES[case_index], str(result)), nprocs=world, join=True)
    text = result.read_text()
    assert text.startswith("OK"), text
    assert int(text.split()[1]) > 20 and int(text.split()[2]) > 0


def test_balanced_read_ranges():
    from shasta_b200 import distributed as D
    b = D.balanced_read_ranges([1] * 10, 4)
    assert b[0] == 0 and b[-1] == 10 and all(b[i] <= b[i + 1] for i in range(4))
    b = D.balanced_read_ranges([100, 1, 1, 1, 100], 2)
    w = [100, 1, 1, 1, 100]
    assert b[0] == 0 and b[2] == 5 and abs(sum(w[:b[1]]) - sum(w[b[1]:])) <= 100
    assert D.read_bits(1) == 1 and D.read_bits(2) == 1 and D.read_bits(3) == 2 and D.read_bits(1 << 20) == 20
