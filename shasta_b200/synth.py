"""Synthetic Nanopore-shaped read sets in MARKER space (numpy; counter-based RNG).

The hot path consumes marker rows (`CompressedMarker`, /root/reference src/Marker.hpp:56-69:
uint32 kmerId + uint24 position, 7 bytes) laid out as a CSR indexed by
`OrientedReadId = (readId << 1) | strand` (src/ReadId.hpp:35-155), with the strand-1 row equal
to the strand-0 row reversed and reverse-complemented (src/MarkerFinder.cpp:92-100).
This generator produces exactly that layout without going through bases:

* a random "genome" of markers (k-mer ids uniform in [0, 4^k), gaps uniform in
  [1, 2*meanGap-1] bases, mean 1/0.0738 ~ 13.55 bases as measured in SURVEY.md section 8),
  with periodic copied blocks (repeats) so that crowded LowHash buckets exist;
* reads = windows of the genome, log-normal length, random strand, each genome marker dropped
  with probability `drop` and a spurious marker inserted with probability `ins`
  (marker-level image of sequencing errors).

Every random draw is a pure function of (seed, stream, i, j) (splitmix64 finaliser), so the
CUDA generator used by bench.py (csrc/synth.cu) reproduces the same data bit for bit.
"""
from __future__ import annotations

import numpy as np

U64 = np.uint64
_C1 = U64(0x9E3779B97F4A7C15)
_C2 = U64(0xBF58476D1CE4E5B9)
_C3 = U64(0x94D049BB133111EB)


def mix64(seed, stream, a, b=0):
    """splitmix64 finaliser of (seed, stream, a, b); vectorised over a and b."""
    with np.errstate(over="ignore"):
        x = (U64(seed) ^ (U64(stream) * _C1)) + np.asarray(a, dtype=U64) * _C2 + np.asarray(b, dtype=U64) * _C3
        x ^= x >> U64(30)
        x *= _C2
        x ^= x >> U64(27)
        x *= _C3
        x ^= x >> U64(31)
    return x


def unit(x):
    """uint64 -> float64 in [0,1) using the top 53 bits."""
    return (x >> U64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def reverse_complement_kmer(kmer, k):
    """Bit-plane reverse complement, src/ShortBaseSequence.hpp:92-118 (vectorised)."""
    kmer = np.asarray(kmer, dtype=np.uint32)
    mask = np.uint32((1 << k) - 1)
    lsb = ~kmer & mask
    msb = ~(kmer >> np.uint32(k)) & mask
    rl = np.zeros_like(kmer)
    rm = np.zeros_like(kmer)
    for i in range(k):
        rl |= ((lsb >> np.uint32(i)) & np.uint32(1)) << np.uint32(k - 1 - i)
        rm |= ((msb >> np.uint32(i)) & np.uint32(1)) << np.uint32(k - 1 - i)
    return (rm << np.uint32(k)) | rl


class SynthParams:
    def __init__(self, reads=1000, k=10, genome_markers=200_000, mean_gap=13.55,
                 n50_bases=20_000, sigma=0.5, min_bases=10_000, max_bases=1 << 23,
                 drop=0.12, ins=0.05, repeat_period=5000, repeat_len=200,
                 palindromic_every=0, seed=1):
        self.reads = reads
        self.k = k
        self.genome_markers = genome_markers
        self.mean_gap = mean_gap
        self.n50_bases = n50_bases
        self.sigma = sigma
        self.min_bases = min_bases
        self.max_bases = max_bases
        self.drop = drop
        self.ins = ins
        self.repeat_period = repeat_period
        self.repeat_len = repeat_len
        self.palindromic_every = palindromic_every
        self.seed = seed


def genome(p: SynthParams):
    """Returns (kmerId uint32[G], pos uint64[G]) of the marker-space genome."""
    G = p.genome_markers
    g = np.arange(G, dtype=U64)
    # Repeats: the first repeat_len markers of every period copy the k-mers of a block
    # chosen pseudo-randomly among the earlier part of the genome.
    src = g.copy()
    if p.repeat_period and p.repeat_len:
        period = g // U64(p.repeat_period)
        off = g % U64(p.repeat_period)
        isrep = (off < U64(p.repeat_len)) & (period > 0)
        srcblock = mix64(p.seed, 9, period) % np.maximum(period, U64(1))
        src = np.where(isrep, srcblock * U64(p.repeat_period) + U64(p.repeat_len) + off, g)
    kmer = (mix64(p.seed, 1, src) % U64(1 << (2 * p.k))).astype(np.uint32)
    maxgap = int(2 * p.mean_gap) - 1
    gap = U64(1) + mix64(p.seed, 2, g) % U64(maxgap)
    pos = np.cumsum(gap) - gap
    return kmer, pos


def read_windows(p: SynthParams):
    """Per-read genome window: (start int64[R], span int64[R], rev uint8[R]) in genome-marker units."""
    G = p.genome_markers
    R = p.reads
    r = np.arange(R, dtype=U64)
    # Log-normal read length (bases) with the requested N50: for a log-normal, N50 = exp(mu + sigma^2).
    u1 = unit(mix64(p.seed, 3, r, 0))
    u2 = unit(mix64(p.seed, 3, r, 1))
    z = np.sqrt(-2.0 * np.log(1.0 - u1)) * np.cos(2.0 * np.pi * u2)
    mu = np.log(p.n50_bases) - p.sigma ** 2
    length = np.clip(np.exp(mu + p.sigma * z), p.min_bases, p.max_bases)
    span = np.minimum(np.maximum((length / p.mean_gap).astype(np.int64), 8), G - 1)
    start = (mix64(p.seed, 4, r) % (U64(G) - span.astype(U64))).astype(np.int64)
    rev = (mix64(p.seed, 8, r) & U64(1)).astype(np.uint8)
    return start, span, rev


def generate(p: SynthParams):
    """Returns dict(toc uint64[2R+1], data uint8[M*7], flags uint8[R], kmer uint32[M], pos uint32[M])."""
    gk, gpos = genome(p)
    R = p.reads
    start, span, rev = read_windows(p)
    k4 = U64(1 << (2 * p.k))
    dropT = p.drop
    insT = p.ins

    rows = []
    for i in range(R):
        gi = np.arange(start[i], start[i] + span[i], dtype=U64)
        keep = unit(mix64(p.seed, 5, U64(i), gi)) >= dropT
        insm = unit(mix64(p.seed, 6, U64(i), gi)) < insT
        insk = (mix64(p.seed, 7, U64(i), gi) % k4).astype(np.uint32)
        base = gpos[start[i]]
        pg = (gpos[start[i]:start[i] + span[i]] - base).astype(np.int64)
        # interleave: genome marker g (if kept) then inserted marker (if any), position pg+1
        n = int(keep.sum() + insm.sum())
        km = np.empty(n, np.uint32)
        ps = np.empty(n, np.int64)
        slot = np.cumsum(keep.astype(np.int64) + insm.astype(np.int64)) - (keep.astype(np.int64) + insm.astype(np.int64))
        km[slot[keep]] = gk[start[i]:start[i] + span[i]][keep]
        ps[slot[keep]] = pg[keep]
        islot = slot + keep.astype(np.int64)
        km[islot[insm]] = insk[insm]
        ps[islot[insm]] = pg[insm] + 1
        total_len = int(pg[-1]) + p.k + 2
        if rev[i]:
            km = reverse_complement_kmer(km[::-1], p.k)
            ps = (total_len - p.k - ps[::-1])
        rows.append((km, ps, total_len))

    counts = np.array([len(x[0]) for x in rows], dtype=np.int64)
    toc = np.zeros(2 * R + 1, dtype=U64)
    toc[1:] = np.cumsum(np.repeat(counts, 2)).astype(U64)
    M = int(toc[-1])
    kmer = np.empty(M, np.uint32)
    pos = np.empty(M, np.uint32)
    for i, (km, ps, total_len) in enumerate(rows):
        a, b, c = int(toc[2 * i]), int(toc[2 * i + 1]), int(toc[2 * i + 2])
        kmer[a:b] = km
        pos[a:b] = ps
        kmer[b:c] = reverse_complement_kmer(km[::-1], p.k)
        pos[b:c] = total_len - p.k - ps[::-1]
    flags = read_flags(p)
    return dict(toc=toc, data=pack_markers(kmer, pos), flags=flags, kmer=kmer, pos=pos, k=p.k)


def read_flags(p: SynthParams):
    flags = np.zeros(p.reads, np.uint8)
    if p.palindromic_every:
        flags[p.palindromic_every - 1::p.palindromic_every] = 1
    return flags


def pack_markers(kmer, pos):
    """(uint32 kmerId, uint24 position) -> 7-byte little-endian records (src/Marker.hpp:56-69)."""
    M = len(kmer)
    data = np.empty((M, 7), np.uint8)
    data[:, 0:4] = np.asarray(kmer, dtype="<u4").view(np.uint8).reshape(M, 4)
    data[:, 4:7] = np.asarray(pos, dtype="<u4").view(np.uint8).reshape(M, 4)[:, 0:3]
    return data.reshape(-1)


def unpack_markers(data):
    d = np.asarray(data, np.uint8).reshape(-1, 7)
    kmer = np.ascontiguousarray(d[:, 0:4]).view("<u4").reshape(-1)
    p = np.zeros((d.shape[0], 4), np.uint8)
    p[:, 0:3] = d[:, 4:7]
    return kmer.copy(), p.view("<u4").reshape(-1).copy()
