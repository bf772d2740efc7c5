// TEST INFRASTRUCTURE: thin C entry points over the reference's own classes used by createReadGraph2
// (src/AssemblerReadGraph2.cpp): shasta::Histogram2 (src/Histogram.cpp, compiled unmodified) and the AlignmentInfo accessors
// (src/Alignment.hpp). Used to pin oracle/readgraph_oracle.c; nothing here is product code.
#include "Alignment.hpp"
#include "Histogram.hpp"
#include <cstring>
using namespace shasta;

extern "C" double ref_histogram2_threshold(const double* x, uint64_t n, double start, double stop, uint64_t binCount, double fraction)
{
    Histogram2 h(start, stop, binCount, false, false, true);
    for(uint64_t i = 0; i < n; i++) h.update(x[i]);
    return h.thresholdByCumulativeProportion(fraction);
}

// out: minAlignedFraction, markerCount, maxDrift, maxSkip, max(leftTrim, rightTrim) of a 64-byte AlignmentData record's info
extern "C" void ref_alignment_indicators(const uint32_t* record16, double* out5)
{
    static_assert(sizeof(AlignmentInfo) == 52, "AlignmentInfo is 13 words");
    AlignmentInfo info;
    std::memcpy(&info, record16 + 3, sizeof(info));
    const auto trims = info.computeTrim();
    out5[0] = info.minAlignedFraction();
    out5[1] = double(info.markerCount);
    out5[2] = double(info.maxDrift);
    out5[3] = double(info.maxSkip);
    out5[4] = double(std::max(trims.first, trims.second));
}
