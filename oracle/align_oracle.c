/*
 * TEST INFRASTRUCTURE — NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this file's library.
 *
 * CPU restatement (plain C) of the alignment half of the hot path of chanzuckerberg/shasta:
 *   Assembler::computeAlignments          src/AssemblerAlign.cpp:208-495
 *   Assembler::alignOrientedReads3        src/AssemblerAlign3.cpp:23-313      (method 3)
 *   Align4::align / Aligner               src/Align4.cpp:30-1087              (method 4)
 *   AlignmentInfo::create, filters        src/Alignment.cpp:67-113, src/Alignment.hpp:105-297
 *   shasta::compress                      src/compressAlignment.cpp:11-70
 *
 * Parity status
 *   - Everything AROUND the dynamic programme is PINNED: oracle/_ref compiles the reference's own
 *     Alignment.cpp, compressAlignment.cpp and Align4.cpp unmodified (Align4.cpp against shim headers
 *     whose globalAlignment() calls orc_overlap_align below) and tests/test_oracle_align.py checks this
 *     restatement against them and against the testAlignmentCompression vectors
 *     (src/compressAlignment.cpp:161-192).
 *   - The dynamic programme itself (orc_overlap_align) is "PARITY UNPINNED": in the reference it is
 *     seqan::globalAlignment from SeqAn 2.x (libseqan2-dev, not vendored, absent from the build
 *     container, and no reference test pins its output). Scores are those of any correct overlap
 *     alignment; the TIE-BREAK rule among equal-score paths follows SURVEY.md Appendix A
 *     (diagonal > vertical > horizontal; first strict maximum in column-major order over last-row /
 *     last-column cells), lives in include/shb_dp_policy.h (shared with the CUDA kernels) and is applied in this one
 *     function, which also records whether the chosen path was exposed to a tie at all.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#include <limits.h>
#include <math.h>
#include <pthread.h>

uint32_t orc_murmurhash2(const void* key, int len, uint32_t seed);
uint32_t orc_kmer_downsampling_hash(uint32_t kmerId, uint32_t k);

#define ORC_MIN_VALUE INT_MIN     /* seqan::MinValue<int>::VALUE */

/* ---------------------------------------------------------------------------------------------
 * Tie-break policy of the DP: compile-time default from include/shb_dp_policy.h (shared with the CUDA
 * kernels), switchable at run time so that the exposure of a data set to the choice can be measured.
 * bit 0: diagonal wins ties against a gap move; bit 1: vertical before horizontal; bit 2: first maximum
 * (column-major) is the end cell.
 */
#include "../include/shb_dp_policy.h"
static int orc_dp_policy = SHB_DP_POLICY_BITS;
void orc_set_dp_policy(int bits) { orc_dp_policy = bits & 7; }
int orc_get_dp_policy(void) { return orc_dp_policy; }

/* Tie exposure of the DP calls made by the calling thread since the last reset:
 * bit 0: some cell ON THE CHOSEN PATH had two or more predecessors giving its score (a co-optimal
 *        path branches off there); bit 1: another end-cell candidate had the best score. */
static __thread int orc_tls_tie = 0;
void orc_reset_tie_flags(void) { orc_tls_tie = 0; }
int orc_get_tie_flags(void) { return orc_tls_tie; }

/* Per-thread scratch of the DP (grow-only; one DP call per candidate and stage would otherwise malloc/free
 * a trace of nx*W bytes each time). */
static __thread uint8_t* tlsTrace = NULL; static __thread size_t tlsTraceCap = 0;
static __thread int32_t* tlsRows = NULL;  static __thread size_t tlsRowsCap = 0;
void orc_release_thread_scratch(void)
{
    free(tlsTrace); tlsTrace = NULL; tlsTraceCap = 0;
    free(tlsRows); tlsRows = NULL; tlsRowsCap = 0;
}

/* ---------------------------------------------------------------------------------------------
 * Overlap alignment, linear gaps, all four end gaps free, optional band lo <= i - j <= hi where
 * i indexes sequence a (horizontal, read 0) and j sequence b (vertical, read 1).
 * Call sites restated: src/AssemblerAlign3.cpp:117-122,254-260; src/Align4.cpp:1027-1033.
 * Output: the diagonal steps of the chosen optimal path, in path order, as (x,y) = (i-1,j-1)
 * pairs, regardless of match/mismatch. Returns the score, or ORC_MIN_VALUE if the band misses the
 * matrix. *pathOut is malloc'ed.
 */
int orc_overlap_align(const uint32_t* a, int64_t nx, const uint32_t* b, int64_t ny,
                      int match, int mismatch, int gap, int banded, int64_t lo, int64_t hi,
                      uint32_t** pathOut, uint64_t* pathLen)
{
    *pathOut = NULL; *pathLen = 0;
    if(!banded) { lo = -ny; hi = nx; }
    if(lo > hi || hi < -ny || lo > nx) return ORC_MIN_VALUE;
    if(lo < -ny) lo = -ny;
    if(hi > nx) hi = nx;
    const int diagWins = orc_dp_policy & 1, vertFirst = (orc_dp_policy >> 1) & 1, endFirst = (orc_dp_policy >> 2) & 1;
    const int64_t W = hi - lo + 1;                 /* cells per column at most */
    const size_t traceBytes = (size_t)(nx + 1) * (size_t)W;
    if(traceBytes > tlsTraceCap) { free(tlsTrace); tlsTraceCap = traceBytes + traceBytes / 4 + 4096; tlsTrace = (uint8_t*)malloc(tlsTraceCap); }
    if((size_t)(2 * (W + 2)) > tlsRowsCap) { free(tlsRows); tlsRowsCap = (size_t)(2 * (W + 2)) + 1024; tlsRows = (int32_t*)malloc(sizeof(int32_t) * tlsRowsCap); }
    uint8_t* trace = tlsTrace;
    int32_t* prev = tlsRows;
    int32_t* cur = tlsRows + (W + 2);
    int64_t prevLo = 0, prevHi = -1;
    int best = ORC_MIN_VALUE; int64_t bi = -1, bj = -1; int haveBest = 0, endTie = 0;
    enum { T_NONE = 0, T_DIAG = 1, T_VERT = 2, T_HORZ = 3, T_TIE = 4 };

    for(int64_t i = 0; i <= nx; i++) {
        int64_t jlo = i - hi; if(jlo < 0) jlo = 0;
        int64_t jhi = i - lo; if(jhi > ny) jhi = ny;
        uint8_t* tcol = trace + (size_t)i * (size_t)W;
        for(int64_t j = jlo; j <= jhi; j++) {
            int32_t s; uint8_t t;
            if(i == 0 || j == 0) { s = 0; t = T_NONE; }
            else {
                /* diagonal predecessor (i-1,j-1) is always in band */
                const int32_t d = prev[j - 1 - prevLo] + ((a[i-1] == b[j-1]) ? match : mismatch);
                const int haveV = (j - 1 >= jlo);                       /* vertical: (i, j-1) */
                const int haveH = (j <= prevHi && j >= prevLo);         /* horizontal: (i-1, j) */
                const int32_t v = haveV ? cur[j - 1 - jlo] + gap : 0;
                const int32_t h = haveH ? prev[j - prevLo] + gap : 0;
                /* the better gap move, then against the diagonal (include/shb_dp_policy.h) */
                int haveG = 0; int32_t g = 0; uint8_t tg = T_NONE; int ties = 0;
                if(haveV && haveH) {
                    haveG = 1;
                    if(v > h || (v == h && vertFirst)) { g = v; tg = T_VERT; } else { g = h; tg = T_HORZ; }
                    if(v == h) ties = 1;
                } else if(haveV) { haveG = 1; g = v; tg = T_VERT; }
                else if(haveH) { haveG = 1; g = h; tg = T_HORZ; }
                if(!haveG || d > g || (d == g && diagWins)) { s = d; t = T_DIAG; if(haveG && d == g) ties = 1; else ties = 0; }
                else { s = g; t = tg; if(d == g) ties = 1; }
                if(ties) t |= T_TIE;
            }
            cur[j - jlo] = s;
            tcol[j - jlo] = t;
            if(j == ny || i == nx) {
                if(haveBest && s == best) endTie = 1;
                if(!haveBest || s > best || (s == best && !endFirst)) {
                    if(!haveBest || s > best) endTie = 0;
                    best = s; bi = i; bj = j; haveBest = 1;
                }
            }
        }
        int32_t* tmp = prev; prev = cur; cur = tmp;
        prevLo = jlo; prevHi = jhi;
    }
    if(!haveBest) return ORC_MIN_VALUE;
    if(endTie) orc_tls_tie |= 2;

    /* Traceback. */
    uint64_t cap = 1024, n = 0;
    uint32_t* path = (uint32_t*)malloc(sizeof(uint32_t) * 2 * cap);
    int64_t i = bi, j = bj;
    while(i > 0 && j > 0) {
        int64_t jlo = i - hi; if(jlo < 0) jlo = 0;
        const uint8_t tt = trace[(size_t)i * (size_t)W + (size_t)(j - jlo)];
        const uint8_t t = tt & 3;
        if(tt & T_TIE) orc_tls_tie |= 1;
        if(t == T_DIAG) {
            if(n == cap) { cap *= 2; path = (uint32_t*)realloc(path, sizeof(uint32_t) * 2 * cap); }
            path[2*n] = (uint32_t)(i - 1); path[2*n+1] = (uint32_t)(j - 1); n++;
            i--; j--;
        } else if(t == T_VERT) j--;
        else if(t == T_HORZ) i--;
        else break;
    }
    for(uint64_t k = 0; k < n / 2; k++) {
        uint32_t t0 = path[2*k], t1 = path[2*k+1];
        path[2*k] = path[2*(n-1-k)]; path[2*k+1] = path[2*(n-1-k)+1];
        path[2*(n-1-k)] = t0; path[2*(n-1-k)+1] = t1;
    }
    *pathOut = path; *pathLen = n;
    return best;
}


/* ---------------------------------------------------------------------------------------------
 * Alignment = list of (ordinal0, ordinal1); AlignmentInfo — src/Alignment.hpp:86-200.
 * The 64-byte AlignmentData record (src/Alignment.hpp:419-447) is written as 16 uint32 words:
 *  [0] readId0 [1] readId1 [2] isSameStrand
 *  [3..5] data[0]{markerCount,firstOrdinal,lastOrdinal} [6..8] data[1]{...}
 *  [9] markerCount [10] minOrdinalOffset [11] maxOrdinalOffset [12] averageOrdinalOffset
 *  [13] maxSkip [14] maxDrift [15] flags (isInReadGraph:1 = 0)
 */
typedef struct {
    uint32_t markerCount0, first0, last0, markerCount1, first1, last1;
    uint32_t markerCount; int32_t minOff, maxOff, avgOff; uint32_t maxSkip, maxDrift;
} orc_info;

/* AlignmentInfo::create, src/Alignment.cpp:67-113. */
void orc_alignment_info(const uint32_t* ord, uint64_t n, uint32_t nx, uint32_t ny, orc_info* info)
{
    info->markerCount = (uint32_t)n;
    info->markerCount0 = nx; info->markerCount1 = ny;
    info->first0 = n ? ord[0] : 0; info->last0 = n ? ord[2*(n-1)] : 0;
    info->first1 = n ? ord[1] : 0; info->last1 = n ? ord[2*(n-1)+1] : 0;
    int32_t mn = INT_MAX, mx = INT_MIN; uint32_t maxSkip = 0, maxDrift = 0; double sum = 0.;
    for(uint64_t i = 0; i < n; i++) {
        const int32_t off = (int32_t)ord[2*i] - (int32_t)ord[2*i+1];
        if(off < mn) mn = off;
        if(off > mx) mx = off;
        sum += (double)off;
        if(i) {
            const uint32_t s0 = (uint32_t)abs((int32_t)ord[2*i] - (int32_t)ord[2*i-2]);
            const uint32_t s1 = (uint32_t)abs((int32_t)ord[2*i+1] - (int32_t)ord[2*i-1]);
            if(s0 > maxSkip) maxSkip = s0;
            if(s1 > maxSkip) maxSkip = s1;
            const int32_t poff = (int32_t)ord[2*i-2] - (int32_t)ord[2*i-1];
            const uint32_t d = (uint32_t)abs(off - poff);
            if(d > maxDrift) maxDrift = d;
        }
    }
    info->minOff = mn; info->maxOff = mx; info->maxSkip = maxSkip; info->maxDrift = maxDrift;
    /* int32_t(std::round(sum/double(markerCount))); for an empty alignment the reference converts
       NaN (implementation defined); empty alignments are never stored, we write 0. */
    info->avgOff = n ? (int32_t)round(sum / (double)n) : 0;
}

static double alignedFraction(const orc_info* f, int i)
{
    const uint32_t range = i ? (f->last1 + 1 - f->first1) : (f->last0 + 1 - f->first0);
    return (double)f->markerCount / (double)range;                 /* src/Alignment.hpp:266-270 */
}
static double minAlignedFractionOf(const orc_info* f)
{
    const double a = alignedFraction(f, 0), b = alignedFraction(f, 1);
    return a < b ? a : b;
}
static void computeTrim(const orc_info* f, uint32_t* left, uint32_t* right)   /* Alignment.hpp:279-284 */
{
    const uint32_t l0 = f->first0, l1 = f->first1;
    const uint32_t r0 = f->markerCount0 - 1 - f->last0, r1 = f->markerCount1 - 1 - f->last1;
    *left = l0 < l1 ? l0 : l1; *right = r0 < r1 ? r0 : r1;
}
static int isContaining(const orc_info* f, uint32_t maxTrim)                    /* Alignment.hpp:290-297 */
{
    if(f->first0 <= maxTrim && f->markerCount0 - 1 - f->last0 <= maxTrim) return 1;
    if(f->first1 <= maxTrim && f->markerCount1 - 1 - f->last1 <= maxTrim) return 1;
    return 0;
}


/* ---------------------------------------------------------------------------------------------
 * shasta::compress — src/compressAlignment.cpp:11-70; formats src/compressAlignment.hpp:102-320
 * (GCC little-endian bit-field layout: fields are allocated from the least significant bit).
 * out must hold 16 bytes per ordinal pair. Returns the number of bytes written.
 */
uint64_t orc_compress_alignment(const uint32_t* ord, uint64_t n, uint8_t* out)
{
    uint64_t w = 0;
    uint32_t ordinal0 = 0, ordinal1 = 0;
    for(uint64_t i = 0; i < n; ) {
        const int32_t skip0 = (int32_t)ord[2*i] - (int32_t)ordinal0;
        const int32_t skip1 = (int32_t)ord[2*i+1] - (int32_t)ordinal1;
        ordinal0 = ord[2*i]; ordinal1 = ord[2*i+1];
        uint32_t len = 1;
        for(uint64_t j = i + 1; j < n; j++, len++) {
            if(ord[2*j] != ordinal0 + 1) break;
            if(ord[2*j+1] != ordinal1 + 1) break;
            ++ordinal0; ++ordinal1;
        }
        i += len;
        const uint64_t nm1 = len - 1;
        if(skip0 >= 0 && skip0 <= 3 && skip1 >= 0 && skip1 <= 3 && len <= 8) {
            out[w++] = (uint8_t)(0u | ((uint32_t)skip0 << 1) | ((uint32_t)skip1 << 3) | (nm1 << 5));
        } else if(skip0 >= -8 && skip0 <= 7 && skip1 >= -8 && skip1 <= 7 && len <= 32) {
            const uint16_t v = (uint16_t)(1u | (((uint32_t)skip0 & 0xFu) << 3) | (((uint32_t)skip1 & 0xFu) << 7) | (nm1 << 11));
            memcpy(out + w, &v, 2); w += 2;
        } else if(skip0 >= -512 && skip0 <= 511 && skip1 >= -512 && skip1 <= 511 && len <= 512) {
            const uint32_t v = 3u | (((uint32_t)skip0 & 0x3FFu) << 3) | (((uint32_t)skip1 & 0x3FFu) << 13) | ((uint32_t)nm1 << 23);
            memcpy(out + w, &v, 4); w += 4;
        } else if(skip0 >= -524288 && skip0 <= 524287 && skip1 >= -524288 && skip1 <= 524287 && len <= 2097152) {
            const uint64_t v = 5ull | (((uint64_t)(int64_t)skip0 & 0xFFFFFull) << 3) | (((uint64_t)(int64_t)skip1 & 0xFFFFFull) << 23) | (nm1 << 43);
            memcpy(out + w, &v, 8); w += 8;
        } else {
            const uint32_t v[4] = {7u, (uint32_t)skip0, (uint32_t)skip1, (uint32_t)nm1};
            memcpy(out + w, v, 16); w += 16;
        }
    }
    return w;
}

/* shasta::decompress — src/compressAlignment.cpp:73-137. ordOut must hold the decoded pairs. */
uint64_t orc_decompress_alignment(const uint8_t* s, uint64_t bytes, uint32_t* ordOut, uint64_t cap)
{
    uint64_t pos = 0, n = 0;
    uint32_t ordinal0 = 0, ordinal1 = 0;
    while(pos < bytes) {
        int32_t skip0, skip1; uint32_t len;
        const uint8_t c = s[pos];
        if((c & 1u) == 0) {
            skip0 = (c >> 1) & 3; skip1 = (c >> 3) & 3; len = ((c >> 5) & 7u) + 1; pos += 1;
        } else if((c & 7u) == 1) {
            uint16_t v; memcpy(&v, s + pos, 2); pos += 2;
            skip0 = (int32_t)((v >> 3) & 0xF); if(skip0 & 8) skip0 -= 16;
            skip1 = (int32_t)((v >> 7) & 0xF); if(skip1 & 8) skip1 -= 16;
            len = (uint32_t)(v >> 11) + 1;
        } else if((c & 7u) == 3) {
            uint32_t v; memcpy(&v, s + pos, 4); pos += 4;
            skip0 = (int32_t)((v >> 3) & 0x3FF); if(skip0 & 0x200) skip0 -= 0x400;
            skip1 = (int32_t)((v >> 13) & 0x3FF); if(skip1 & 0x200) skip1 -= 0x400;
            len = (v >> 23) + 1;
        } else if((c & 7u) == 5) {
            uint64_t v; memcpy(&v, s + pos, 8); pos += 8;
            skip0 = (int32_t)((v >> 3) & 0xFFFFF); if(skip0 & 0x80000) skip0 -= 0x100000;
            skip1 = (int32_t)((v >> 23) & 0xFFFFF); if(skip1 & 0x80000) skip1 -= 0x100000;
            len = (uint32_t)(v >> 43) + 1;
        } else {
            uint32_t v[4]; memcpy(v, s + pos, 16); pos += 16;
            skip0 = (int32_t)v[1]; skip1 = (int32_t)v[2]; len = v[3] + 1;
        }
        ordinal0 += (uint32_t)skip0; ordinal1 += (uint32_t)skip1;
        for(uint32_t i = 0; i < len; i++) {
            if(n < cap) { ordOut[2*n] = ordinal0 + i; ordOut[2*n+1] = ordinal1 + i; }
            n++;
        }
        ordinal0 += len - 1; ordinal1 += len - 1;
    }
    return n;
}


/* ---------------------------------------------------------------------------------------------
 * Options: field for field AlignOptions (src/AssemblerOptions.hpp:177-199) as used on the path.
 */
typedef struct {
    uint32_t alignMethod;           /* 3 or 4 (1 = unbanded on all markers is also accepted) */
    uint32_t k;                     /* k-mer length: the method-3 downsampling hash is recomputed from it */
    uint64_t maxSkip, maxDrift, maxTrim, minAlignedMarkerCount;
    double minAlignedFraction;
    int32_t matchScore, mismatchScore, gapScore;
    double downsamplingFactor;
    int32_t bandExtend, maxBand;
    uint32_t suppressContainments;
    uint64_t align4DeltaX, align4DeltaY, align4MinEntryCountPerCell, align4MaxDistanceFromBoundary;
} orc_align_options;

typedef struct { uint32_t* ord; uint64_t n, cap; } orc_alignment;

static void alignmentPush(orc_alignment* al, uint32_t x, uint32_t y)
{
    if(al->n == al->cap) { al->cap = al->cap ? 2 * al->cap : 256; al->ord = (uint32_t*)realloc(al->ord, sizeof(uint32_t) * 2 * al->cap); }
    al->ord[2*al->n] = x; al->ord[2*al->n+1] = y; al->n++;
}

/* Keep the equal-kmer diagonal steps of a path (src/AssemblerAlign3.cpp:279-295, src/Align4.cpp:1052-1068). */
static void extractEqualSteps(const uint32_t* path, uint64_t len, const uint32_t* a, const uint32_t* b, orc_alignment* al)
{
    al->n = 0;
    for(uint64_t i = 0; i < len; i++) {
        const uint32_t x = path[2*i], y = path[2*i+1];
        if(a[x] == b[y]) alignmentPush(al, x, y);
    }
}

/* Method 1: SeqAn unbanded on all markers, src/AssemblerAlign1.cpp:129-148. status 0 ok. */
static int alignMethod1(const uint32_t* a, uint32_t nx, const uint32_t* b, uint32_t ny, const orc_align_options* o, orc_alignment* al)
{
    uint32_t* path; uint64_t len;
    const int score = orc_overlap_align(a, nx, b, ny, o->matchScore, o->mismatchScore, o->gapScore, 0, 0, 0, &path, &len);
    if(score == ORC_MIN_VALUE) { al->n = 0; return 1; }
    extractEqualSteps(path, len, a, b, al);
    free(path);
    return 0;
}

/* Method 3, src/AssemblerAlign3.cpp:23-313. Returns 0, or 1 when the reference throws
   ("SeqAn banded alignment computation failed", candidate skipped). */
int orc_align_method3(const uint32_t* a, uint32_t nx, const uint32_t* b, uint32_t ny,
                      const orc_align_options* o, orc_alignment* al)
{
    al->n = 0;
    const uint32_t hashThreshold = (uint32_t)(o->downsamplingFactor * (double)UINT32_MAX);    /* :71-72 */
    uint32_t* da = (uint32_t*)malloc(sizeof(uint32_t) * (nx + 1)); uint32_t* oa = (uint32_t*)malloc(sizeof(uint32_t) * (nx + 1));
    uint32_t* db = (uint32_t*)malloc(sizeof(uint32_t) * (ny + 1)); uint32_t* ob = (uint32_t*)malloc(sizeof(uint32_t) * (ny + 1));
    uint32_t na = 0, nb = 0;
    for(uint32_t i = 0; i < nx; i++) if(orc_kmer_downsampling_hash(a[i], o->k) < hashThreshold) { da[na] = a[i]; oa[na] = i; na++; }
    for(uint32_t i = 0; i < ny; i++) if(orc_kmer_downsampling_hash(b[i], o->k) < hashThreshold) { db[nb] = b[i]; ob[nb] = i; nb++; }
    int status = 0;
    if(na == 0 || nb == 0) goto done;                                           /* :100-106 */
    {
        uint32_t* path; uint64_t len;
        const int score = orc_overlap_align(da, na, db, nb, o->matchScore, o->mismatchScore, o->gapScore, 0, 0, 0, &path, &len);
        if(score == ORC_MIN_VALUE) { status = 1; goto done; }                   /* :127-129 */
        if(len == 0) { free(path); goto done; }                                 /* alignmentLength == n0+n1, :185-191 */
        int32_t offsetMin = INT_MAX, offsetMax = INT_MIN;                       /* :197-221 */
        for(uint64_t i = 0; i < len; i++) {
            const uint32_t x = path[2*i], y = path[2*i+1];
            if(da[x] == db[y]) {
                const int32_t off = (int32_t)oa[x] - (int32_t)ob[y];
                if(off < offsetMin) offsetMin = off;
                if(off > offsetMax) offsetMax = off;
            }
        }
        free(path);
        /* :222-239 (computed in 32-bit wrap-around arithmetic like the compiled reference) */
        const int32_t bandMin = (int32_t)((uint32_t)offsetMin - (uint32_t)o->bandExtend);
        const int32_t bandMax = (int32_t)((uint32_t)offsetMax + (uint32_t)o->bandExtend);
        if((int32_t)((uint32_t)bandMax - (uint32_t)bandMin) > o->maxBand) goto done;
        const int score2 = orc_overlap_align(a, nx, b, ny, o->matchScore, o->mismatchScore, o->gapScore, 1, bandMin, bandMax, &path, &len);
        if(score2 == ORC_MIN_VALUE) { status = 1; goto done; }                  /* :264-266 */
        extractEqualSteps(path, len, a, b, al);
        free(path);
    }
done:
    free(da); free(oa); free(db); free(ob);
    return status;
}


/* ---------------------------------------------------------------------------------------------
 * Method 4 — src/Align4.cpp. Cell grid in (X,Y) = (x+y, y+nx-1-x), cells of deltaX x deltaY.
 * Only the NUMBER of alignment-matrix entries per cell matters (createCells, :380-436), so the
 * sparse matrix is restated as a dense count grid.
 */
static void getxy(int32_t X, int32_t Y, int32_t nx, int32_t* x, int32_t* y)      /* :183-191 (C division truncates) */
{
    *x = (X - Y + nx - 1) / 2;
    *y = (X + Y - nx + 1) / 2;
}

typedef struct { uint32_t a, b; } orc_kp;          /* (kmerId, ordinal) */
static int cmp_kp(const void* p, const void* q)
{
    const orc_kp* x = (const orc_kp*)p; const orc_kp* y = (const orc_kp*)q;
    return x->a < y->a ? -1 : (x->a > y->a ? 1 : 0);
}

/* Returns 0. The chosen alignment is put in al (empty if none kept). If tieOut != NULL it is set to 1
   when two kept components tie on markerCount (the reference's choice then depends on
   std::unordered_map iteration order, src/Align4.cpp:798-866,132-140). */
int orc_align_method4(const uint32_t* a, uint32_t nx, const uint32_t* b, uint32_t ny,
                      const orc_align_options* o, orc_alignment* al, int* tieOut)
{
    al->n = 0;
    if(tieOut) *tieOut = 0;
    if(nx == 0 || ny == 0) return 0;
    const uint32_t deltaX = (uint32_t)o->align4DeltaX, deltaY = (uint32_t)o->align4DeltaY;
    const uint32_t sizeXY = nx + ny - 1;
    const uint32_t nIX = (sizeXY - 1) / deltaX + 1, nIY = (sizeXY - 1) / deltaY + 1;
    uint32_t* count = (uint32_t*)calloc((size_t)nIX * nIY, sizeof(uint32_t));

    /* createAlignmentMatrix :195-267 (merge join of the kmer-sorted markers). */
    orc_kp* sa = (orc_kp*)malloc(sizeof(orc_kp) * nx); orc_kp* sb = (orc_kp*)malloc(sizeof(orc_kp) * ny);
    for(uint32_t i = 0; i < nx; i++) { sa[i].a = a[i]; sa[i].b = i; }
    for(uint32_t i = 0; i < ny; i++) { sb[i].a = b[i]; sb[i].b = i; }
    qsort(sa, nx, sizeof(orc_kp), cmp_kp); qsort(sb, ny, sizeof(orc_kp), cmp_kp);
    for(uint32_t i0 = 0, i1 = 0; i0 < nx && i1 < ny; ) {
        if(sa[i0].a < sb[i1].a) i0++;
        else if(sb[i1].a < sa[i0].a) i1++;
        else {
            const uint32_t kmer = sa[i0].a;
            uint32_t e0 = i0, e1 = i1;
            while(e0 < nx && sa[e0].a == kmer) e0++;
            while(e1 < ny && sb[e1].a == kmer) e1++;
            for(uint32_t p = i0; p < e0; p++) for(uint32_t q = i1; q < e1; q++) {
                const uint32_t x = sa[p].b, y = sb[q].b;
                const uint32_t X = x + y, Y = nx + y - x - 1;
                count[(size_t)(Y / deltaY) * nIX + X / deltaX]++;
            }
            i0 = e0; i1 = e1;
        }
    }
    free(sa); free(sb);

    /* createCells :380-436. flags: 1 exists, 2 nearLeftOrTop, 4 nearRightOrBottom, 8 fwd, 16 bwd. */
    uint8_t* cell = (uint8_t*)calloc((size_t)nIX * nIY, 1);
    const uint64_t maxD = o->align4MaxDistanceFromBoundary;
    for(uint32_t iY = 0; iY < nIY; iY++) for(uint32_t iX = 0; iX < nIX; iX++) {
        const size_t idx = (size_t)iY * nIX + iX;
        if(count[idx] == 0 || (int64_t)count[idx] < (int64_t)o->align4MinEntryCountPerCell) continue;
        int32_t x, y; uint32_t dLeft, dRight, dTop, dBottom;
        getxy((int32_t)(iX * deltaX), (int32_t)((iY + 1) * deltaY), (int32_t)nx, &x, &y);      /* :536-553 */
        dLeft = x < 0 ? 0u : (uint32_t)x;
        getxy((int32_t)((iX + 1) * deltaX), (int32_t)(iY * deltaY), (int32_t)nx, &x, &y);      /* :562-578 */
        dRight = (x >= (int32_t)nx - 1) ? 0u : (uint32_t)(nx - 1 - (uint32_t)x);
        getxy((int32_t)(iX * deltaX), (int32_t)(iY * deltaY), (int32_t)nx, &x, &y);            /* :587-604 */
        dTop = y < 0 ? 0u : (uint32_t)y;
        getxy((int32_t)((iX + 1) * deltaX), (int32_t)((iY + 1) * deltaY), (int32_t)nx, &x, &y); /* :613-629 */
        dBottom = (y >= (int32_t)ny - 1) ? 0u : (uint32_t)(ny - 1 - (uint32_t)y);
        uint8_t f = 1;
        if(dLeft < maxD || dTop < maxD) f |= 2;
        if(dRight < maxD || dBottom < maxD) f |= 4;
        cell[idx] = f;
    }
    free(count);

    /* forwardSearch :682-729 and backwardSearch :736-787 (reachability; visiting order is irrelevant). */
    uint32_t* stack = (uint32_t*)malloc(sizeof(uint32_t) * 2 * ((size_t)nIX * nIY + 1));
    size_t sp = 0;
    for(uint32_t iY = 0; iY < nIY; iY++) for(uint32_t iX = 0; iX < nIX; iX++) {
        uint8_t* c = &cell[(size_t)iY * nIX + iX];
        if((*c & 1) && (*c & 2)) { *c |= 8; stack[2*sp] = iX; stack[2*sp+1] = iY; sp++; }
    }
    while(sp) {
        sp--; const uint32_t iX0 = stack[2*sp], iY0 = stack[2*sp+1];
        for(int dY = -1; dY <= 1; dY++) {
            const int64_t iY1 = (int64_t)iY0 + dY; if(iY1 < 0 || iY1 >= nIY) continue;
            for(uint32_t dX = 0; dX <= 1; dX++) {
                const uint32_t iX1 = iX0 + dX; if(iX1 >= nIX) continue;
                uint8_t* c = &cell[(size_t)iY1 * nIX + iX1];
                if((*c & 1) && !(*c & 8)) { *c |= 8; stack[2*sp] = iX1; stack[2*sp+1] = (uint32_t)iY1; sp++; }
            }
        }
    }
    for(uint32_t iY = 0; iY < nIY; iY++) for(uint32_t iX = 0; iX < nIX; iX++) {
        uint8_t* c = &cell[(size_t)iY * nIX + iX];
        if((*c & 1) && (*c & 4) && (*c & 8)) { *c |= 16; stack[2*sp] = iX; stack[2*sp+1] = iY; sp++; }
    }
    while(sp) {
        sp--; const uint32_t iX0 = stack[2*sp], iY0 = stack[2*sp+1];
        for(int dY = -1; dY <= 1; dY++) {
            const int64_t iY1 = (int64_t)iY0 + dY; if(iY1 < 0 || iY1 >= nIY) continue;
            for(int dX = -1; dX <= 0; dX++) {
                const int64_t iX1 = (int64_t)iX0 + dX; if(iX1 < 0) continue;
                uint8_t* c = &cell[(size_t)iY1 * nIX + iX1];
                if((*c & 1) && !(*c & 16)) { *c |= 16; stack[2*sp] = (uint32_t)iX1; stack[2*sp+1] = (uint32_t)iY1; sp++; }
            }
        }
    }

    /* findActiveCellsConnectedComponents :792-868: 8-neighbourhood components of active cells
       (active = forward and backward accessible). label[] = component id in raster order of first cell. */
    int32_t* label = (int32_t*)malloc(sizeof(int32_t) * (size_t)nIX * nIY);
    for(size_t i = 0; i < (size_t)nIX * nIY; i++) label[i] = -1;
    uint32_t nComp = 0;
    uint32_t* compYMin = NULL; uint32_t* compYMax = NULL; size_t compCap = 0;
    for(uint32_t iY = 0; iY < nIY; iY++) for(uint32_t iX = 0; iX < nIX; iX++) {
        const size_t idx = (size_t)iY * nIX + iX;
        if((cell[idx] & 24) != 24 || label[idx] >= 0) continue;
        if(nComp == compCap) { compCap = compCap ? 2 * compCap : 16; compYMin = (uint32_t*)realloc(compYMin, 4 * compCap); compYMax = (uint32_t*)realloc(compYMax, 4 * compCap); }
        compYMin[nComp] = iY; compYMax[nComp] = iY;
        label[idx] = (int32_t)nComp; stack[0] = iX; stack[1] = iY; sp = 1;
        while(sp) {
            sp--; const uint32_t iX0 = stack[2*sp], iY0 = stack[2*sp+1];
            if(iY0 < compYMin[nComp]) compYMin[nComp] = iY0;
            if(iY0 > compYMax[nComp]) compYMax[nComp] = iY0;
            for(int dY = -1; dY <= 1; dY++) for(int dX = -1; dX <= 1; dX++) {
                if(!dX && !dY) continue;
                const int64_t iX1 = (int64_t)iX0 + dX, iY1 = (int64_t)iY0 + dY;
                if(iX1 < 0 || iY1 < 0 || iX1 >= nIX || iY1 >= nIY) continue;
                const size_t j = (size_t)iY1 * nIX + iX1;
                if((cell[j] & 24) == 24 && label[j] < 0) { label[j] = (int32_t)nComp; stack[2*sp] = (uint32_t)iX1; stack[2*sp+1] = (uint32_t)iY1; sp++; }
            }
        }
        nComp++;
    }
    free(stack); free(label); free(cell);

    /* computeBandedAlignments :874-989 + best selection :128-147. Scores are hard-coded 6/-1/-1
       (src/Align4.hpp:159-161: the Options scores are never copied into the Aligner). */
    orc_alignment best = {0, 0, 0}; uint32_t bestCount = 0; int haveBestAl = 0; int tie = 0;
    orc_alignment cur = {0, 0, 0};
    for(uint32_t cI = 0; cI < nComp; cI++) {
        const uint32_t YMin = compYMin[cI] * deltaY, YMax = (compYMax[cI] + 1) * deltaY - 1;
        const int32_t bandMin = (int32_t)nx - 1 - (int32_t)YMax, bandMax = (int32_t)nx - 1 - (int32_t)YMin;
        const int32_t bandWidth = bandMax - bandMin + 1;
        if(bandWidth > (int64_t)o->maxBand) continue;
        uint32_t* path; uint64_t len;
        const int score = orc_overlap_align(a, nx, b, ny, 6, -1, -1, 1, bandMin, bandMax, &path, &len);
        cur.n = 0;
        if(score != ORC_MIN_VALUE) { extractEqualSteps(path, len, a, b, &cur); free(path); }
        orc_info info; orc_alignment_info(cur.ord, cur.n, nx, ny, &info);
        if(info.markerCount < o->minAlignedMarkerCount) continue;
        if(minAlignedFractionOf(&info) < o->minAlignedFraction) continue;
        if(info.maxSkip > o->maxSkip) continue;
        if(info.maxDrift > o->maxDrift) continue;
        uint32_t lt, rt; computeTrim(&info, &lt, &rt);
        if(lt > o->maxTrim) continue;
        if(rt > o->maxTrim) continue;
        if(!haveBestAl || info.markerCount > bestCount) {
            if(haveBestAl && info.markerCount == bestCount) tie = 1;
            orc_alignment t = best; best = cur; cur = t; bestCount = info.markerCount; haveBestAl = 1;
        } else if(info.markerCount == bestCount) tie = 1;
    }
    free(cur.ord); free(compYMin); free(compYMax);
    if(haveBestAl) { free(al->ord); *al = best; } else { free(best.ord); al->n = 0; }
    if(tieOut) *tieOut = tie;
    return 0;
}


/* ---------------------------------------------------------------------------------------------
 * computeAlignments — src/AssemblerAlign.cpp:208-495. Results are stored in CANDIDATE ORDER (a legal
 * order: the reference's order is thread-schedule dependent).
 */
typedef struct {
    const uint64_t* toc; const uint32_t* kmerIds;
    const uint32_t* candidates; uint64_t n; const orc_align_options* o;
    uint64_t begin, end;
    /* per candidate outputs */
    uint8_t* keep; uint32_t* records; uint8_t** compressed; uint64_t* compressedBytes; uint8_t* tie;
} orc_worker;

static void* workerMain(void* arg)
{
    orc_worker* w = (orc_worker*)arg;
    orc_alignment al = {0, 0, 0};
    for(uint64_t i = w->begin; i < w->end; i++) {
        const uint32_t r0 = w->candidates[3*i], r1 = w->candidates[3*i+1], same = w->candidates[3*i+2];
        const uint64_t o0 = 2ull * r0, o1 = 2ull * r1 + (same ? 0 : 1);        /* :381-382 */
        const uint32_t* a = w->kmerIds + w->toc[o0]; const uint32_t nx = (uint32_t)(w->toc[o0+1] - w->toc[o0]);
        const uint32_t* b = w->kmerIds + w->toc[o1]; const uint32_t ny = (uint32_t)(w->toc[o1+1] - w->toc[o1]);
        int status = 0, tie = 0;
        orc_reset_tie_flags();
        if(w->o->alignMethod == 3) status = orc_align_method3(a, nx, b, ny, w->o, &al);
        else if(w->o->alignMethod == 4) status = orc_align_method4(a, nx, b, ny, w->o, &al, &tie);
        else status = alignMethod1(a, nx, b, ny, w->o, &al);
        /* bit 0: Align4 component tie; bit 1: a DP path of this candidate branched on a tie; bit 2: end-cell tie */
        w->tie[i] = (uint8_t)((tie ? 1 : 0) | (orc_get_tie_flags() << 1));
        w->keep[i] = 0;
        if(status) continue;                                                    /* exception: candidate skipped, :419-434 */
        orc_info info; orc_alignment_info(al.ord, al.n, nx, ny, &info);
        if(al.n == 0) continue;      /* empty alignments are never stored (min 0 / fraction 0 corner) */
        if(al.n < w->o->minAlignedMarkerCount) continue;                        /* :439 */
        if(minAlignedFractionOf(&info) < w->o->minAlignedFraction) continue;    /* :445 */
        uint32_t lt, rt; computeTrim(&info, &lt, &rt);
        if(lt > w->o->maxTrim || rt > w->o->maxTrim) continue;                  /* :450-456 */
        if(info.maxSkip > w->o->maxSkip) continue;                              /* :460-467 */
        if(info.maxDrift > w->o->maxDrift) continue;
        if(w->o->suppressContainments && isContaining(&info, (uint32_t)w->o->maxTrim)) continue;   /* :470 */
        uint32_t* rec = w->records + 16 * i;
        rec[0] = r0; rec[1] = r1; rec[2] = same ? 1 : 0;
        rec[3] = info.markerCount0; rec[4] = info.first0; rec[5] = info.last0;
        rec[6] = info.markerCount1; rec[7] = info.first1; rec[8] = info.last1;
        rec[9] = info.markerCount; rec[10] = (uint32_t)info.minOff; rec[11] = (uint32_t)info.maxOff;
        rec[12] = (uint32_t)info.avgOff; rec[13] = info.maxSkip; rec[14] = info.maxDrift; rec[15] = 0;
        uint8_t* buf = (uint8_t*)malloc(16 * al.n + 16);
        w->compressedBytes[i] = orc_compress_alignment(al.ord, al.n, buf);
        w->compressed[i] = buf;
        w->keep[i] = 1;
    }
    free(al.ord);
    orc_release_thread_scratch();
    return NULL;
}

/*
 * toc uint64[2R+1], kmerIds uint32[M] (the kmerId column of the markers), candidates uint32[n][3].
 * Outputs (malloc'ed): records uint32[count][16], compressedToc uint64[count+1], compressedData bytes,
 * ties uint8[n] per candidate (may be NULL): bit 0 = two kept Align4 components tie on markerCount, bit 1 = a DP path of
 * the candidate passed through a cell with co-optimal predecessors, bit 2 = the DP end cell had a co-optimal rival.
 */
int orc_compute_alignments(const uint64_t* toc, const uint32_t* kmerIds, const uint32_t* candidates, uint64_t n,
                           const orc_align_options* o, uint32_t threads,
                           uint32_t** recordsOut, uint64_t* countOut,
                           uint64_t** compressedTocOut, uint8_t** compressedDataOut, uint8_t** tiesOut)
{
    if(threads == 0) threads = 1;
    if(threads > n) threads = (uint32_t)(n ? n : 1);
    uint8_t* keep = (uint8_t*)calloc(n + 1, 1);
    uint8_t* tie = (uint8_t*)calloc(n + 1, 1);
    uint32_t* rec = (uint32_t*)calloc(16 * (n + 1), sizeof(uint32_t));
    uint8_t** comp = (uint8_t**)calloc(n + 1, sizeof(uint8_t*));
    uint64_t* compBytes = (uint64_t*)calloc(n + 1, sizeof(uint64_t));
    pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * threads);
    orc_worker* w = (orc_worker*)malloc(sizeof(orc_worker) * threads);
    /* Interleaved blocks of 10 candidates per thread grab, like the reference's batch size (:243-249),
       here statically assigned in round-robin chunks to stay deterministic. */
    for(uint32_t t = 0; t < threads; t++) {
        w[t].toc = toc; w[t].kmerIds = kmerIds; w[t].candidates = candidates; w[t].n = n; w[t].o = o;
        w[t].begin = n * t / threads; w[t].end = n * (t + 1) / threads;
        w[t].keep = keep; w[t].records = rec; w[t].compressed = comp; w[t].compressedBytes = compBytes; w[t].tie = tie;
        pthread_create(&th[t], NULL, workerMain, &w[t]);
    }
    for(uint32_t t = 0; t < threads; t++) pthread_join(th[t], NULL);
    uint64_t count = 0, bytes = 0;
    for(uint64_t i = 0; i < n; i++) if(keep[i]) { count++; bytes += compBytes[i]; }
    uint32_t* outRec = (uint32_t*)malloc(64 * (count ? count : 1));
    uint64_t* outToc = (uint64_t*)malloc(8 * (count + 1));
    uint8_t* outData = (uint8_t*)malloc(bytes ? bytes : 1);
    uint64_t k = 0, off = 0;
    for(uint64_t i = 0; i < n; i++) if(keep[i]) {
        memcpy(outRec + 16 * k, rec + 16 * i, 64);
        outToc[k] = off; memcpy(outData + off, comp[i], compBytes[i]); off += compBytes[i];
        free(comp[i]); k++;
    }
    outToc[count] = off;
    *recordsOut = outRec; *countOut = count; *compressedTocOut = outToc; *compressedDataOut = outData;
    if(tiesOut) *tiesOut = tie; else free(tie);
    free(keep); free(rec); free(comp); free(compBytes); free(th); free(w);
    return 0;
}


/* ---------------------------------------------------------------------------------------------
 * computeAlignmentTable — src/AssemblerAlign.cpp:509-571 with OrientedReadPair::getOther
 * (src/OrientedReadPair.hpp:63-85). records: uint32[n][16] AlignmentData records.
 * Outputs (caller allocated): toc uint32[2R+1], table uint32[4n]: for each oriented read the indices of
 * the alignments it is involved in, sorted by (other OrientedReadId, alignment index).
 */
typedef struct { uint32_t other, index; } orc_te;
static int cmp_te(const void* a, const void* b)
{
    const orc_te* x = (const orc_te*)a; const orc_te* y = (const orc_te*)b;
    if(x->other != y->other) return x->other < y->other ? -1 : 1;
    return x->index < y->index ? -1 : (x->index > y->index ? 1 : 0);
}
void orc_compute_alignment_table(const uint32_t* records, uint64_t n, uint64_t R, uint32_t* toc, uint32_t* table)
{
    memset(toc, 0, sizeof(uint32_t) * (2 * R + 1));
    for(uint64_t i = 0; i < n; i++) {
        const uint32_t r0 = records[16*i], r1 = records[16*i+1], same = records[16*i+2] & 0xff;
        const uint32_t o0 = 2 * r0, o1 = 2 * r1 + (same ? 0 : 1);
        toc[o0 + 1]++; toc[o1 + 1]++; toc[(o0 ^ 1) + 1]++; toc[(o1 ^ 1) + 1]++;
    }
    for(uint64_t o = 0; o < 2 * R; o++) toc[o + 1] += toc[o];
    uint32_t* fill = (uint32_t*)calloc(2 * R + 1, sizeof(uint32_t));
    orc_te* e = (orc_te*)malloc(sizeof(orc_te) * (4 * n + 1));
    for(uint64_t i = 0; i < n; i++) {
        const uint32_t r0 = records[16*i], r1 = records[16*i+1], same = records[16*i+2] & 0xff;
        const uint32_t o0 = 2 * r0, o1 = 2 * r1 + (same ? 0 : 1);
        const uint32_t rows[4] = {o0, o1, o0 ^ 1, o1 ^ 1};
        const uint32_t others[4] = {o1, o0, o1 ^ 1, o0 ^ 1};
        for(int k = 0; k < 4; k++) {
            orc_te* slot = &e[toc[rows[k]] + fill[rows[k]]++];
            slot->other = others[k]; slot->index = (uint32_t)i;
        }
    }
    for(uint64_t o = 0; o < 2 * R; o++) qsort(e + toc[o], toc[o+1] - toc[o], sizeof(orc_te), cmp_te);
    for(uint64_t i = 0; i < 4 * n; i++) table[i] = e[i].index;
    free(fill); free(e);
}


/* ---------------------------------------------------------------------------------------------
 * AlignmentCandidates::computeCandidateTable — src/AssemblerAlignmentCandidates.cpp:379-448.
 * candidates: uint32[n][3] (readId0, readId1, isSameStrand). Outputs (caller allocated): toc uint64[2R+1],
 * table uint64[4n]: for each oriented read the indices of the candidates it is involved in (both reads, both
 * strands: :389-399), each row sorted by (other OrientedReadId, candidate index) (:417-440, sort of
 * pair<OrientedReadId, uint32_t>).
 */
void orc_compute_candidate_table(const uint32_t* candidates, uint64_t n, uint64_t R, uint64_t* toc, uint64_t* table)
{
    memset(toc, 0, sizeof(uint64_t) * (2 * R + 1));
    for(uint64_t i = 0; i < n; i++) {
        const uint32_t r0 = candidates[3*i], r1 = candidates[3*i+1], same = candidates[3*i+2] & 0xff;
        const uint32_t o0 = 2 * r0, o1 = 2 * r1 + (same ? 0 : 1);
        toc[o0 + 1]++; toc[o1 + 1]++; toc[(o0 ^ 1) + 1]++; toc[(o1 ^ 1) + 1]++;
    }
    for(uint64_t o = 0; o < 2 * R; o++) toc[o + 1] += toc[o];
    uint64_t* fill = (uint64_t*)calloc(2 * R + 1, sizeof(uint64_t));
    orc_te* e = (orc_te*)malloc(sizeof(orc_te) * (4 * n + 1));
    for(uint64_t i = 0; i < n; i++) {
        const uint32_t r0 = candidates[3*i], r1 = candidates[3*i+1], same = candidates[3*i+2] & 0xff;
        const uint32_t o0 = 2 * r0, o1 = 2 * r1 + (same ? 0 : 1);
        const uint32_t rows[4] = {o0, o1, o0 ^ 1, o1 ^ 1};
        const uint32_t others[4] = {o1, o0, o1 ^ 1, o0 ^ 1};
        for(int k = 0; k < 4; k++) {
            orc_te* slot = &e[toc[rows[k]] + fill[rows[k]]++];
            slot->other = others[k]; slot->index = (uint32_t)i;
        }
    }
    for(uint64_t o = 0; o < 2 * R; o++) qsort(e + toc[o], toc[o+1] - toc[o], sizeof(orc_te), cmp_te);
    for(uint64_t i = 0; i < 4 * n; i++) table[i] = e[i].index;
    free(fill); free(e);
}
