/*
 * TEST INFRASTRUCTURE — NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this file's library.
 *
 * CPU restatement (plain C, single thread) of the LowHash0 half of the hot path of
 * chanzuckerberg/shasta. It is written for clarity and determinism, not speed; every
 * function cites the reference lines it follows (paths relative to /root/reference).
 *
 * Parity status: PINNED. tests/test_oracle_lowhash.py checks this restatement against
 *   (a) the MurmurHash64A / MurmurHash2 known answers of SURVEY.md Appendix B,
 *   (b) golden fixtures produced by the UNMODIFIED reference LowHash0 (oracle/_ref,
 *       generator: tests/golden/make_golden.py) on tests/TinyTest.fasta.gz and on
 *       synthetic marker sets, including the per-iteration (highFrequency,total) lines.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

/* ---------------------------------------------------------------------------------------
 * MurmurHash64A — src/MurmurHash2.cpp:96-137. 64-bit blocks are read little-endian at
 * 4-byte alignment; the tail switch falls through; final avalanche h^=h>>47; h*=m; h^=h>>47.
 */
uint64_t orc_murmurhash64a(const void* key, int len, uint64_t seed)
{
    const uint64_t m = 0xc6a4a7935bd1e995ULL;
    const int r = 47;
    uint64_t h = seed ^ ((uint64_t)len * m);
    const uint8_t* p = (const uint8_t*)key;
    const int nblocks = len / 8;
    for(int i = 0; i < nblocks; i++) {
        uint64_t k;
        memcpy(&k, p + 8*i, 8);
        k *= m; k ^= k >> r; k *= m;
        h ^= k; h *= m;
    }
    const uint8_t* t = p + 8*nblocks;
    switch(len & 7) {
    case 7: h ^= (uint64_t)t[6] << 48;  /* fall through */
    case 6: h ^= (uint64_t)t[5] << 40;  /* fall through */
    case 5: h ^= (uint64_t)t[4] << 32;  /* fall through */
    case 4: h ^= (uint64_t)t[3] << 24;  /* fall through */
    case 3: h ^= (uint64_t)t[2] << 16;  /* fall through */
    case 2: h ^= (uint64_t)t[1] << 8;   /* fall through */
    case 1: h ^= (uint64_t)t[0];
            h *= m;
    }
    h ^= h >> r; h *= m; h ^= h >> r;
    return h;
}

/* MurmurHash2 (32 bit) — src/MurmurHash2.cpp:37-88. Used for the method-3 downsampling hash. */
uint32_t orc_murmurhash2(const void* key, int len, uint32_t seed)
{
    const uint32_t m = 0x5bd1e995u;
    const int r = 24;
    uint32_t h = seed ^ (uint32_t)len;
    const uint8_t* data = (const uint8_t*)key;
    while(len >= 4) {
        uint32_t k;
        memcpy(&k, data, 4);
        k *= m; k ^= k >> r; k *= m;
        h *= m; h ^= k;
        data += 4; len -= 4;
    }
    switch(len) {
    case 3: h ^= (uint32_t)data[2] << 16;   /* fall through */
    case 2: h ^= (uint32_t)data[1] << 8;    /* fall through */
    case 1: h ^= (uint32_t)data[0];
            h *= m;
    }
    h ^= h >> 13; h *= m; h ^= h >> 15;
    return h;
}

/* ---------------------------------------------------------------------------------------
 * K-mer id helpers — src/ShortBaseSequence.hpp:92-118 (id layout = (msbPlane<<k)|lsbPlane,
 * base 0 at the most significant bit of each k-bit plane; reverse complement = reversed
 * order, complemented bases; complement of base b is 3-b, i.e. both planes inverted).
 */
uint32_t orc_reverse_complement_kmer(uint32_t kmerId, uint32_t k)
{
    const uint32_t mask = (k == 16) ? 0xffffu : ((1u << k) - 1u);
    uint32_t lsb = kmerId & mask;
    uint32_t msb = (kmerId >> k) & mask;
    uint32_t rl = 0, rm = 0;
    for(uint32_t i = 0; i < k; i++) {
        rl |= ((~lsb >> i) & 1u) << (k - 1 - i);
        rm |= ((~msb >> i) & 1u) << (k - 1 - i);
    }
    return (rm << k) | rl;
}

/* kmerTable[kmerId].hash — src/AssemblerKmers.cpp:182-186. */
uint32_t orc_kmer_downsampling_hash(uint32_t kmerId, uint32_t k)
{
    const uint64_t n = (uint64_t)kmerId + (uint64_t)orc_reverse_complement_kmer(kmerId, k);
    return orc_murmurhash2(&n, 8, 13477u);
}


/* ---------------------------------------------------------------------------------------
 * LowHash0 — src/LowHash0.cpp:23-257.
 */
typedef struct { uint64_t key; uint32_t oread; } orc_entry;       /* key = bucketId<<32 | hashHigh */
typedef struct { uint64_t key; uint32_t count; } orc_pair;        /* key = r0<<32 | r1<<1 | strand */

static int cmp_entry(const void* a, const void* b)
{
    const orc_entry* x = (const orc_entry*)a; const orc_entry* y = (const orc_entry*)b;
    if(x->key != y->key) return x->key < y->key ? -1 : 1;
    if(x->oread != y->oread) return x->oread < y->oread ? -1 : 1;
    return 0;
}
static int cmp_u64(const void* a, const void* b)
{
    const uint64_t x = *(const uint64_t*)a, y = *(const uint64_t*)b;
    return x < y ? -1 : (x > y ? 1 : 0);
}

static inline uint32_t kmer_of(const uint8_t* data7, uint64_t i)
{
    uint32_t v; memcpy(&v, data7 + 7*i, 4); return v;      /* src/Marker.hpp:56-69 */
}

/*
 * Inputs:  R reads; toc uint64[2R+1] (row = readId*2+strand); data7 = 7-byte CompressedMarker
 *          records; flags = 1 byte per read (bit0 = isPalindromic, src/ReadFlags.hpp:10-30).
 * Outputs: *candidatesOut = malloc'ed uint32[3*n] {readId0, readId1, isSameStrand} in the
 *          reference order (src/LowHash0.cpp:204-214); stats uint64[R][3] (caller allocated);
 *          iterSummary (optional) uint64[2*maxIters] receives (highFrequency,total) per
 *          iteration (src/LowHash0.cpp:185-196); *iterationsOut = iterations executed.
 * Returns 0, or 1 for the reference's "log2MinHashBucketCount is unreasonably small" throw.
 */
int orc_lowhash0(
    uint64_t R, const uint64_t* toc, const uint8_t* data7, const uint8_t* flags,
    uint64_t m, double hashFraction, uint64_t minHashIterationCount,
    double alignmentCandidatesPerRead, uint64_t log2MinHashBucketCount,
    uint64_t minBucketSize, uint64_t maxBucketSize, uint64_t minFrequency,
    uint32_t** candidatesOut, uint64_t* candidateCountOut, uint64_t* stats,
    uint64_t* iterSummary, uint64_t maxIters, uint64_t* iterationsOut)
{
    const uint64_t M = toc[2*R];

    /* Bucket-count rule, src/LowHash0.cpp:69-98. (__builtin_clzl(0) is undefined in the
       reference; we define the zero estimate as log2 = 0.) */
    const uint64_t est = (uint64_t)(hashFraction * (double)M);
    const uint32_t log2Est = est ? (uint32_t)(64 - __builtin_clzll(est)) : 0;
    if(log2MinHashBucketCount == 0) log2MinHashBucketCount = 5 + log2Est;
    else if(log2MinHashBucketCount < log2Est) return 1;
    if(log2MinHashBucketCount > 31) log2MinHashBucketCount = 31;
    const uint64_t mask = (1ULL << log2MinHashBucketCount) - 1ULL;

    /* createKmerIds, src/LowHash0.cpp:261-308. */
    uint32_t* kmerIds = (uint32_t*)malloc(sizeof(uint32_t) * (M + 1));
    for(uint64_t i = 0; i < M; i++) kmerIds[i] = kmer_of(data7, i);

    /* src/LowHash0.cpp:109 */
    const uint64_t hashThreshold = (uint64_t)(hashFraction * (double)UINT64_MAX);

    memset(stats, 0, sizeof(uint64_t) * 3 * R);

    orc_entry* entries = NULL; uint64_t entryCap = 0;
    uint64_t* newPairs = NULL; uint64_t newCap = 0;
    orc_pair* acc = NULL; uint64_t accCount = 0;          /* sorted by key, unique */

    uint64_t highFrequency = 0;
    uint64_t iteration = 0;
    for(;; iteration++) {
        /* Iteration control, src/LowHash0.cpp:136-157. */
        if(minHashIterationCount == 0) {
            const double current = 2. * (double)highFrequency / (double)R;
            if(current >= alignmentCandidatesPerRead) break;
            /* The reference spins forever when the target cannot be reached; the oracle gives up
               (return code 2) once maxIters iterations have run. */
            if(iteration >= maxIters) {
                free(kmerIds); free(entries); free(newPairs); free(acc);
                *candidatesOut = NULL; *candidateCountOut = 0;
                return 2;
            }
        } else if(iteration == minHashIterationCount) break;

        /* pass1, src/LowHash0.cpp:314-360: low hashes of every non-palindromic oriented read. */
        const uint64_t seed = iteration * 37;
        uint64_t n = 0;
        for(uint64_t readId = 0; readId < R; readId++) {
            if(flags[readId] & 1) continue;
            for(uint32_t strand = 0; strand < 2; strand++) {
                const uint64_t o = 2*readId + strand;
                const uint64_t count = toc[o+1] - toc[o];
                if(count < m) continue;
                const uint32_t* p = kmerIds + toc[o];
                for(uint64_t j = 0; j + m <= count; j++) {
                    const uint64_t h = orc_murmurhash64a(p + j, (int)(4*m), seed);
                    if(h < hashThreshold) {
                        if(n == entryCap) {
                            entryCap = entryCap ? 2*entryCap : 1024;
                            entries = (orc_entry*)realloc(entries, entryCap * sizeof(orc_entry));
                        }
                        entries[n].key = ((h & mask) << 32) | (h >> 32);   /* LowHash0.hpp:96-106 */
                        entries[n].oread = (uint32_t)o;
                        n++;
                    }
                }
            }
        }

        /* pass2 = bucket fill (src/LowHash0.cpp:365-398); a sort gives the same buckets. */
        qsort(entries, n, sizeof(orc_entry), cmp_entry);

        /* pass2 statistics + pass3 pair generation (src/LowHash0.cpp:386-393, 403-458). */
        uint64_t np = 0;
        for(uint64_t b0 = 0; b0 < n; ) {
            uint64_t b1 = b0;
            const uint64_t bucketId = entries[b0].key >> 32;
            while(b1 < n && (entries[b1].key >> 32) == bucketId) b1++;
            const uint64_t size = b1 - b0;
            const int cls = (size < minBucketSize) ? 0 : ((size > maxBucketSize) ? 2 : 1);
            for(uint64_t i = b0; i < b1; i++) stats[3*(entries[i].oread >> 1) + cls]++;
            const uint64_t lo = minBucketSize > 2 ? minBucketSize : 2;
            if(size >= lo && size <= maxBucketSize) {
                for(uint64_t i = b0; i < b1; i++) {
                    const uint32_t r0 = entries[i].oread >> 1, s0 = entries[i].oread & 1;
                    for(uint64_t j = b0; j < b1; j++) {
                        if((uint32_t)entries[j].key != (uint32_t)entries[i].key) continue;
                        const uint32_t r1 = entries[j].oread >> 1, s1 = entries[j].oread & 1;
                        if(r1 <= r0) continue;
                        if(np == newCap) {
                            newCap = newCap ? 2*newCap : 1024;
                            newPairs = (uint64_t*)realloc(newPairs, newCap * sizeof(uint64_t));
                        }
                        newPairs[np++] = ((uint64_t)r0 << 32) | ((uint64_t)r1 << 1) | (uint64_t)(s0 != s1);
                    }
                }
            }
            b0 = b1;
        }

        /* sort + merge with uint16 wrap-around sums (src/LowHash0.cpp:462-472, 493-562). */
        qsort(newPairs, np, sizeof(uint64_t), cmp_u64);
        orc_pair* merged = (orc_pair*)malloc(sizeof(orc_pair) * (accCount + np + 1));
        uint64_t i0 = 0, i1 = 0, k = 0;
        while(i0 < accCount || i1 < np) {
            uint64_t key; uint32_t c;
            if(i1 == np || (i0 < accCount && acc[i0].key < newPairs[i1])) { key = acc[i0].key; c = acc[i0].count; i0++; }
            else { key = newPairs[i1]; c = 1; i1++; }
            if(k && merged[k-1].key == key) merged[k-1].count = (uint16_t)(merged[k-1].count + c);
            else { merged[k].key = key; merged[k].count = c; k++; }
        }
        free(acc); acc = merged; accCount = k;

        /* Per-iteration summary, src/LowHash0.cpp:185-196. */
        highFrequency = 0;
        for(uint64_t i = 0; i < accCount; i++) if(acc[i].count >= minFrequency) highFrequency++;
        if(iterSummary && iteration < maxIters) {
            iterSummary[2*iteration] = highFrequency;
            iterSummary[2*iteration+1] = accCount;
        }
    }

    /* Final emission, src/LowHash0.cpp:204-214. */
    uint64_t nOut = 0;
    for(uint64_t i = 0; i < accCount; i++) if(acc[i].count >= minFrequency) nOut++;
    uint32_t* out = (uint32_t*)malloc(12 * (nOut ? nOut : 1));
    uint64_t w = 0;
    for(uint64_t i = 0; i < accCount; i++) {
        if(acc[i].count >= minFrequency) {
            out[3*w+0] = (uint32_t)(acc[i].key >> 32);
            out[3*w+1] = (uint32_t)((acc[i].key & 0xffffffffULL) >> 1);
            out[3*w+2] = (acc[i].key & 1ULL) ? 0u : 1u;    /* strand 0 = same strand */
            w++;
        }
    }
    *candidatesOut = out; *candidateCountOut = nOut;
    if(iterationsOut) *iterationsOut = iteration;
    free(kmerIds); free(entries); free(newPairs); free(acc);
    return 0;
}

void orc_free(void* p) { free(p); }
