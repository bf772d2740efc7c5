// TEST INFRASTRUCTURE — NOT PRODUCT CODE.
// extern "C" access to the reference's own alignment-side translation units, compiled unmodified
// from /root/reference/src by oracle/Makefile:
//   Align4::align            src/Align4.cpp:30-42     (against ref_glue/shims/{seqan/align.h,boost/...,png.h})
//   AlignmentInfo::create    src/Alignment.cpp:67-113
//   shasta::compress         src/compressAlignment.cpp:11-70, testAlignmentCompression :160-220
#include "Align4.hpp"
#include "Alignment.hpp"
#include "compressAlignment.hpp"
#include "Marker.hpp"
#include "MemoryMappedAllocator.hpp"
#include "PngImage.hpp"
#include "orderPairs.hpp"

#include <algorithm>
#include <cstring>
#include <vector>

using namespace shasta;

// PngImage is only reachable from Align4's debug output; link-time stubs (libpng is absent).
PngImage::PngImage(int w, int h) : width(w), height(h) {}
void PngImage::setPixel(int, int, int, int, int) {}
void PngImage::write(const string&) const {}
void PngImage::writeGrid(int, int, int, int) {}

extern "C" {

// ordOut: malloc'ed uint32[2*n] ordinal pairs of the alignment Align4::align returns.
int ref_align4(const uint32_t* a, uint32_t nx, const uint32_t* b, uint32_t ny,
               uint64_t deltaX, uint64_t deltaY, uint64_t minEntryCountPerCell, uint64_t maxDistanceFromBoundary,
               uint64_t minAlignedMarkerCount, double minAlignedFraction, uint64_t maxSkip, uint64_t maxDrift,
               uint64_t maxTrim, uint64_t maxBand, uint32_t** ordOut, uint64_t* nOut)
{
    try {
        std::vector<CompressedMarker> m[2];
        array<vector< pair<KmerId, uint32_t> >, 2> sorted;
        const uint32_t* src[2] = {a, b};
        const uint32_t n[2] = {nx, ny};
        for(int i = 0; i < 2; i++) {
            m[i].resize(n[i]);
            sorted[i].resize(n[i]);
            for(uint32_t j = 0; j < n[i]; j++) {
                m[i][j].kmerId = src[i][j];
                m[i][j].position = j;
                sorted[i][j] = make_pair(KmerId(src[i][j]), j);
            }
            // src/AssemblerAlign4.cpp:172 / :242
            sort(sorted[i].begin(), sorted[i].end(), OrderPairsByFirstOnly<KmerId, uint32_t>());
        }
        array<Align4::CompressedMarkers, 2> cm = {
            Align4::CompressedMarkers(m[0].data(), m[0].data() + nx),
            Align4::CompressedMarkers(m[1].data(), m[1].data() + ny)};
        array<span< const pair<KmerId, uint32_t> >, 2> sm = {
            span< const pair<KmerId, uint32_t> >(sorted[0].data(), sorted[0].data() + nx),
            span< const pair<KmerId, uint32_t> >(sorted[1].data(), sorted[1].data() + ny)};
        Align4::Options o;
        o.deltaX = deltaX; o.deltaY = deltaY; o.minEntryCountPerCell = minEntryCountPerCell;
        o.maxDistanceFromBoundary = maxDistanceFromBoundary; o.minAlignedMarkerCount = minAlignedMarkerCount;
        o.minAlignedFraction = minAlignedFraction; o.maxSkip = maxSkip; o.maxDrift = maxDrift; o.maxTrim = maxTrim;
        o.maxBand = maxBand; o.matchScore = 6; o.mismatchScore = -1; o.gapScore = -1;
        MemoryMapped::ByteAllocator byteAllocator("", 4096, 256ULL * 1024 * 1024);
        Alignment alignment;
        AlignmentInfo info;
        Align4::align(cm, sm, o, byteAllocator, alignment, info, false);
        const uint64_t k = alignment.ordinals.size();
        uint32_t* out = (uint32_t*)malloc(8 * (k ? k : 1));
        for(uint64_t i = 0; i < k; i++) { out[2*i] = alignment.ordinals[i][0]; out[2*i+1] = alignment.ordinals[i][1]; }
        *ordOut = out; *nOut = k;
        return 0;
    } catch(const std::exception& e) {
        fprintf(stderr, "ref_align4: %s\n", e.what());
        return 1;
    }
}

// out: 12 uint32 = words [3..14] of the AlignmentData record layout used by the oracle.
void ref_alignment_info(const uint32_t* ord, uint64_t n, uint32_t nx, uint32_t ny, uint32_t* out)
{
    Alignment al;
    for(uint64_t i = 0; i < n; i++) al.ordinals.push_back(array<uint32_t, 2>{ord[2*i], ord[2*i+1]});
    AlignmentInfo info(al, nx, ny);
    out[0] = info.data[0].markerCount; out[1] = info.data[0].firstOrdinal; out[2] = info.data[0].lastOrdinal;
    out[3] = info.data[1].markerCount; out[4] = info.data[1].firstOrdinal; out[5] = info.data[1].lastOrdinal;
    out[6] = info.markerCount; out[7] = uint32_t(info.minOrdinalOffset); out[8] = uint32_t(info.maxOrdinalOffset);
    out[9] = uint32_t(info.averageOrdinalOffset); out[10] = info.maxSkip; out[11] = info.maxDrift;
}

// out must hold 16 bytes per pair; returns bytes written.
uint64_t ref_compress_alignment(const uint32_t* ord, uint64_t n, uint8_t* out)
{
    Alignment al;
    for(uint64_t i = 0; i < n; i++) al.ordinals.push_back(array<uint32_t, 2>{ord[2*i], ord[2*i+1]});
    string s;
    shasta::compress(al, s);
    memcpy(out, s.data(), s.size());
    return s.size();
}

uint64_t ref_decompress_alignment(const uint8_t* bytes, uint64_t nBytes, uint32_t* ordOut, uint64_t cap)
{
    Alignment al;
    shasta::decompress(span<const char>(reinterpret_cast<const char*>(bytes), reinterpret_cast<const char*>(bytes) + nBytes), al);
    for(uint64_t i = 0; i < al.ordinals.size() && i < cap; i++) { ordOut[2*i] = al.ordinals[i][0]; ordOut[2*i+1] = al.ordinals[i][1]; }
    return al.ordinals.size();
}

// The reference's own self-test (throws on failure).
int ref_test_alignment_compression()
{
    try { shasta::testAlignmentCompression(); return 0; } catch(...) { return 1; }
}

} // extern "C"
