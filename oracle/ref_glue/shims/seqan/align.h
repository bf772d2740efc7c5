// TEST INFRASTRUCTURE shim for the 12 SeqAn 2 names the reference's alignment code uses
// (src/Align4.cpp:993-1043, src/AssemblerAlign3.cpp:42-135). SeqAn is absent from the build
// container, so globalAlignment() forwards to the ONE function that defines the
// overlap DP and its tie-break rule for this repository: orc_overlap_align (oracle/align_oracle.c).
// "Parity unpinned" with respect to real SeqAn — see the header of that file.
#pragma once
#include <cstdint>
#include <vector>
#include <climits>

extern "C" int orc_overlap_align(const uint32_t* a, int64_t nx, const uint32_t* b, int64_t ny,
                                 int match, int mismatch, int gap, int banded, int64_t lo, int64_t hi,
                                 uint32_t** pathOut, uint64_t* pathLen);
extern "C" void orc_free(void*);

namespace seqan {

template<class T> struct String : public std::vector<T> {};
struct Owner_ {};
template<class T = void> struct Dependent {};
template<class TString, class TSpec = Owner_> struct StringSet : public std::vector<TString> {};

template<class T> inline void appendValue(String<T>& s, const T& v) { s.push_back(v); }
template<class T, class V> inline void appendValue(String<T>& s, const V& v) { s.push_back(T(v)); }
template<class TString, class TSpec> inline void appendValue(StringSet<TString, TSpec>& s, const TString& v) { s.push_back(v); }
template<class T> inline uint64_t length(const String<T>& s) { return s.size(); }

struct Simple {};
template<class TValue, class TSpec> struct Score {
    TValue match, mismatch, gap;
    Score(TValue m, TValue mm, TValue g) : match(m), mismatch(mm), gap(g) {}
};
template<bool A, bool B, bool C, bool D> struct AlignConfig {};
struct LinearGaps {};
template<class T> struct MinValue;
template<> struct MinValue<int> { static const int VALUE = INT_MIN; };

template<class TStringSet> struct Alignment {};
template<class TSpec> struct Graph;
template<class TStringSet> struct Graph< Alignment<TStringSet> > {
    std::vector<uint32_t> seq[2];
    std::vector<uint32_t> path;         // diagonal steps (x,y)
    template<class TOther> explicit Graph(const TOther& set)
    {
        seq[0].assign(set[0].begin(), set[0].end());
        seq[1].assign(set[1].begin(), set[1].end());
    }
};

template<class TGraph, class TScore, class TConfig>
inline int globalAlignmentImpl(TGraph& g, const TScore& s, int banded, int lo, int hi)
{
    uint32_t* p = nullptr; uint64_t n = 0;
    const int score = orc_overlap_align(g.seq[0].data(), (int64_t)g.seq[0].size(), g.seq[1].data(), (int64_t)g.seq[1].size(),
                                        s.match, s.mismatch, s.gap, banded, lo, hi, &p, &n);
    g.path.assign(p, p + 2 * n);
    if(p) orc_free(p);
    return score;
}
template<class TGraph, class TScore, class TConfig>
inline int globalAlignment(TGraph& g, const TScore& s, const TConfig&, const LinearGaps&)
{ return globalAlignmentImpl<TGraph, TScore, TConfig>(g, s, 0, 0, 0); }
template<class TGraph, class TScore, class TConfig>
inline int globalAlignment(TGraph& g, const TScore& s, const TConfig&, int lo, int hi, const LinearGaps&)
{ return globalAlignmentImpl<TGraph, TScore, TConfig>(g, s, 1, lo, hi); }

// Two gapped rows (gap symbol 45), concatenated; end gaps written explicitly.
template<class TGraph, class T> inline void convertAlignment(const TGraph& g, String<T>& out)
{
    const uint32_t gapValue = 45;
    std::vector<uint32_t> row0, row1;
    uint64_t i = 0, j = 0;
    const uint64_t n = g.path.size() / 2;
    for(uint64_t k = 0; k <= n; k++) {
        const uint64_t x = (k < n) ? g.path[2*k] : g.seq[0].size();
        const uint64_t y = (k < n) ? g.path[2*k+1] : g.seq[1].size();
        for(; i < x; i++) { row0.push_back(g.seq[0][i]); row1.push_back(gapValue); }
        for(; j < y; j++) { row0.push_back(gapValue); row1.push_back(g.seq[1][j]); }
        if(k < n) { row0.push_back(g.seq[0][i++]); row1.push_back(g.seq[1][j++]); }
    }
    out.clear();
    for(uint32_t v : row0) out.push_back(T(v));
    for(uint32_t v : row1) out.push_back(T(v));
}

} // namespace seqan
