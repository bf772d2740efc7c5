// LowHash0 on the GPU: host orchestration of the kernels in lowhash_kernels.cuh.
// Mirrors LowHash0::LowHash0 (src/LowHash0.cpp:23-257 of chanzuckerberg/shasta) step by step;
// see DESIGN.md for the data layout and the per-kernel roofline.
#include "context.cuh"
#include "lowhash_kernels.cuh"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <vector>

namespace shb {

thread_local uint64_t g_launchCount = 0;

namespace {

struct EventTimer {
    cudaEvent_t a = nullptr, b = nullptr;
    EventTimer() { cudaEventCreate(&a); cudaEventCreate(&b); }
    ~EventTimer() { if(a) cudaEventDestroy(a); if(b) cudaEventDestroy(b); }
};

uint32_t bitsFor(uint64_t maxValue)
{
    uint32_t b = 0;
    while(b < 64 && (maxValue >> b)) b++;
    return b ? b : 1;
}

template<class T> T readScalar(const T* dev, cudaStream_t stream)
{
    T v;
    SHB_CUDA(cudaMemcpyAsync(&v, dev, sizeof(T), cudaMemcpyDeviceToHost, stream));
    SHB_CUDA(cudaStreamSynchronize(stream));
    return v;
}

void launchSweep(const SweepArgs& a, uint32_t blocks, cudaStream_t stream)
{
    switch(a.m) {
    case 1: SHB_LAUNCH(lowhashSweepKernel<1>, blocks, kSweepThreads, 0, stream, a); break;
    case 2: SHB_LAUNCH(lowhashSweepKernel<2>, blocks, kSweepThreads, 0, stream, a); break;
    case 3: SHB_LAUNCH(lowhashSweepKernel<3>, blocks, kSweepThreads, 0, stream, a); break;
    case 4: SHB_LAUNCH(lowhashSweepKernel<4>, blocks, kSweepThreads, 0, stream, a); break;
    case 5: SHB_LAUNCH(lowhashSweepKernel<5>, blocks, kSweepThreads, 0, stream, a); break;
    case 6: SHB_LAUNCH(lowhashSweepKernel<6>, blocks, kSweepThreads, 0, stream, a); break;
    case 7: SHB_LAUNCH(lowhashSweepKernel<7>, blocks, kSweepThreads, 0, stream, a); break;
    case 8: SHB_LAUNCH(lowhashSweepKernel<8>, blocks, kSweepThreads, 0, stream, a); break;
    default: SHB_LAUNCH(lowhashSweepKernel<0>, blocks, kSweepThreads, 0, stream, a); break;
    }
}

// Sorted-by-key (keys[,vals]) -> head flags, exclusive segment index, segment starts.
// Returns the number of segments (one host sync).
uint32_t buildSegments(shb_context* c, const uint64_t* sortedKeys, uint32_t n, int shift)
{
    cudaStream_t st = c->stream;
    c->flagsBuf.reserve(n);
    c->indexBuf.reserve(n);
    c->segStartBuf.reserve(uint64_t(n) + 1);
    c->scanWs.reserve(scanWorkspaceElements(n));
    c->scalars.reserve(64);
    SHB_LAUNCH(headFlagsKernel, ceilDiv(n, 256), 256, 0, st, sortedKeys, n, shift, c->flagsBuf.get());
    uint32_t* total = reinterpret_cast<uint32_t*>(c->scalars.get() + 32);
    exclusiveScan<uint32_t>(c->flagsBuf.get(), c->indexBuf.get(), n, total, c->scanWs.get(), st);
    SHB_LAUNCH(segmentStartsKernel, ceilDiv(n, 256), 256, 0, st,
               (const uint32_t*)c->flagsBuf.get(), (const uint32_t*)c->indexBuf.get(), n, c->segStartBuf.get());
    return readScalar<uint32_t>(total, st);
}

struct Accumulator {
    uint64_t count = 0;
    bool inB = false;       // which of the acc ping-pong buffers holds the data
};

uint64_t* accKeys(shb_context* c, const Accumulator& a) { return a.inB ? c->accKeysB.get() : c->accKeysA.get(); }
uint32_t* accVals(shb_context* c, const Accumulator& a) { return a.inB ? c->accValsB.get() : c->accValsA.get(); }

void accReserve(shb_context* c, Accumulator& a, uint64_t n)
{
    // Keep both ping-pong buffers the same size; only the live one is preserved.
    if(a.inB) { c->accKeysB.reserve(n, true, c->stream); c->accValsB.reserve(n, true, c->stream); }
    else      { c->accKeysA.reserve(n, true, c->stream); c->accValsA.reserve(n, true, c->stream); }
}

// Sort the accumulated (pairKey,count) items by key and sum the counts of equal keys
// (the order-independent equivalent of LowHash0::merge, src/LowHash0.cpp:493-562; the uint16
// wrap-around is applied when the frequency is read, frequencyFlagsKernel).
void mergeAccumulator(shb_context* c, Accumulator& acc, uint32_t readBits)
{
    if(acc.count == 0) return;
    SHB_REQUIRE(acc.count < (1ull << 32), SHB_ERR_INVALID, "LowHash0: candidate accumulator exceeds 2^32-1 items.");
    const uint32_t n = uint32_t(acc.count);
    cudaStream_t st = c->stream;
    // The sort ping-pongs between A and B.
    uint64_t* kA = acc.inB ? c->accKeysB.get() : c->accKeysA.get();
    uint32_t* vA = acc.inB ? c->accValsB.get() : c->accValsA.get();
    DeviceBuffer<uint64_t>& otherK = acc.inB ? c->accKeysA : c->accKeysB;
    DeviceBuffer<uint32_t>& otherV = acc.inB ? c->accValsA : c->accValsB;
    otherK.reserve(n);
    otherV.reserve(n);
    const int ranges[2][2] = {{0, int(readBits) + 1}, {32, 32 + int(readBits)}};
    const bool flipped = radixSort<true>(kA, otherK.get(), vA, otherV.get(), n, ranges, 2, c->sortWs, st);
    if(flipped) acc.inB = !acc.inB;
    const uint64_t* sortedK = accKeys(c, acc);
    const uint32_t* sortedV = accVals(c, acc);
    const uint32_t numSeg = buildSegments(c, sortedK, n, 0);
    // Reduce into the other buffer pair.
    DeviceBuffer<uint64_t>& outK = acc.inB ? c->accKeysA : c->accKeysB;
    DeviceBuffer<uint32_t>& outV = acc.inB ? c->accValsA : c->accValsB;
    outK.reserve(numSeg);
    outV.reserve(numSeg);
    SHB_LAUNCH(segmentSumKernel, ceilDiv(numSeg, 256), 256, 0, st, sortedK, sortedV,
               (const uint32_t*)c->segStartBuf.get(), numSeg, outK.get(), outV.get());
    acc.inB = !acc.inB;
    acc.count = numSeg;
}

uint64_t countHighFrequency(shb_context* c, const Accumulator& acc, uint64_t minFrequency, bool keepOffsets)
{
    if(acc.count == 0) return 0;
    const uint32_t n = uint32_t(acc.count);
    cudaStream_t st = c->stream;
    c->flagsBuf.reserve(n);
    c->indexBuf.reserve(n);
    c->scanWs.reserve(scanWorkspaceElements(n));
    c->scalars.reserve(64);
    SHB_LAUNCH(frequencyFlagsKernel, ceilDiv(n, 256), 256, 0, st, (const uint32_t*)accVals(c, acc), n, minFrequency, c->flagsBuf.get());
    uint32_t* total = reinterpret_cast<uint32_t*>(c->scalars.get() + 32);
    exclusiveScan<uint32_t>(c->flagsBuf.get(), c->indexBuf.get(), n, total, c->scanWs.get(), st);
    (void)keepOffsets;
    return readScalar<uint32_t>(total, st);
}

} // namespace


// The whole LowHash0 computation on the markers held by the context (single GPU).
void lowhash0(shb_context* c, const shb_lowhash_params& p,
              void** candidatesOut, uint64_t* candidateCountOut,
              uint64_t* statsOut, uint64_t* iterSummary, uint64_t maxIterSummary,
              shb_lowhash_result* result)
{
    SHB_REQUIRE(c->haveMarkers, SHB_ERR_STATE, "Markers are not accessible.");
    SHB_REQUIRE(c->readBegin == 0 && c->readEnd == c->readCountTotal, SHB_ERR_STATE,
                "shb_lowhash0 needs all reads on this GPU (use the staged multi-GPU calls otherwise).");
    SHB_REQUIRE(p.m >= 1 && p.m <= 32, SHB_ERR_INVALID, "MinHash.m must be between 1 and 32 in this implementation.");
    SHB_REQUIRE(c->readCountTotal < (1ull << 31), SHB_ERR_INVALID, "Too many reads.");
    SHB_CUDA(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    g_launchCount = 0;

    const uint64_t R = c->readCountTotal;
    const uint64_t M = c->localMarkerCount;
    if(R == 0) {        // the reference would spin forever on 0/0 in its iteration control; return nothing
        *candidatesOut = malloc(1);
        *candidateCountOut = 0;
        if(result) memset(result, 0, sizeof(*result));
        return;
    }

    // Bucket-count rule, src/LowHash0.cpp:69-98.
    const uint64_t totalLowHashCountEstimate = uint64_t(p.hashFraction * double(c->totalMarkerCount));
    const uint32_t log2Estimate = totalLowHashCountEstimate ? uint32_t(64 - __builtin_clzll(totalLowHashCountEstimate)) : 0;
    uint64_t log2BucketCount = p.log2MinHashBucketCount;
    if(log2BucketCount == 0) log2BucketCount = 5 + log2Estimate;
    else SHB_REQUIRE(log2BucketCount >= log2Estimate, SHB_ERR_INVALID, "log2MinHashBucketCount is unreasonably small.");
    if(log2BucketCount > 31) log2BucketCount = 31;
    const uint64_t bucketMask = (1ull << log2BucketCount) - 1ull;

    // src/LowHash0.cpp:109
    const uint64_t hashThreshold = uint64_t(double(p.hashFraction) * double(std::numeric_limits<uint64_t>::max()));

    const bool perIteration = (p.perIterationMerge != 0) || (p.minHashIterationCount == 0);
    const uint32_t readBits = bitsFor(R ? R - 1 : 0);

    EventTimer totalTimer, sweepTimer;
    SHB_CUDA(cudaEventRecord(totalTimer.a, st));
    double sweepMs = 0.;
    uint64_t sweepLaunches = 0, lowHashCount = 0, pairCount = 0;

    c->stats.reserve(3 * R + 1);
    SHB_CUDA(cudaMemsetAsync(c->stats.get(), 0, (3 * R + 1) * sizeof(unsigned long long), st));
    c->scalars.reserve(64);

    // Capacity of one iteration's low-hash slab.
    uint64_t capacity = uint64_t(1.25 * p.hashFraction * double(M)) + 65536;
    if(capacity > M + 1) capacity = M + 1;

    Accumulator acc;
    uint64_t highFrequency = 0;
    uint64_t iteration = 0;
    bool done = false;

    while(!done) {
        // Iteration control, src/LowHash0.cpp:136-157.
        uint32_t group = 1;
        if(p.minHashIterationCount == 0) {
            const double current = 2. * double(highFrequency) / double(R);
            if(current >= p.alignmentCandidatesPerRead) break;
        } else {
            if(iteration == p.minHashIterationCount) break;
            if(!perIteration) group = uint32_t(std::min<uint64_t>(kMaxFusedIterations, p.minHashIterationCount - iteration));
        }

        // ---- pass 1: hash sweep for `group` iterations in one pass over the k-mer ids -----------
        std::vector<unsigned long long> counts(group, 0);
        for(;;) {
            c->sweepKeys.reserve(capacity * group);
            c->sweepVals.reserve(capacity * group);
            SHB_CUDA(cudaMemsetAsync(c->scalars.get(), 0, kMaxFusedIterations * sizeof(unsigned long long), st));
            SweepArgs a;
            a.kmerIds = c->kmerIds;
            a.markerCount = M;
            a.toc = c->toc.get();
            a.orientedReadCount = uint32_t(2 * (c->readEnd - c->readBegin));
            a.orientedReadBase = uint32_t(2 * c->readBegin);
            a.readFlags = c->readFlags.get();
            a.m = uint32_t(p.m);
            a.hashThreshold = hashThreshold;
            a.bucketMask = bucketMask;
            a.iterationBegin = uint32_t(iteration);
            a.iterationCount = group;
            a.keys = c->sweepKeys.get();
            a.vals = c->sweepVals.get();
            a.capacity = capacity;
            a.counts = c->scalars.get();
            if(M >= p.m) {
                SHB_CUDA(cudaEventRecord(sweepTimer.a, st));
                launchSweep(a, ceilDiv(M, kSweepTile), st);
                SHB_CUDA(cudaEventRecord(sweepTimer.b, st));
                sweepLaunches++;
            }
            SHB_CUDA(cudaMemcpyAsync(counts.data(), c->scalars.get(), group * sizeof(unsigned long long), cudaMemcpyDeviceToHost, st));
            SHB_CUDA(cudaStreamSynchronize(st));
            if(M >= p.m) {
                float ms = 0.f;
                SHB_CUDA(cudaEventElapsedTime(&ms, sweepTimer.a, sweepTimer.b));
                sweepMs += ms;
            }
            const unsigned long long worst = *std::max_element(counts.begin(), counts.end());
            if(worst <= capacity) break;
            capacity = worst + worst / 8 + 1024;        // slab overflow: grow and redo this group
        }

        // ---- passes 2 and 3 for each iteration of the group ---------------------------------------
        for(uint32_t s = 0; s < group; s++, iteration++) {
            const uint64_t n64 = counts[s];
            SHB_REQUIRE(n64 < (1ull << 32), SHB_ERR_INVALID, "LowHash0: more than 2^32-1 low hashes in one iteration.");
            const uint32_t n = uint32_t(n64);
            lowHashCount += n;
            if(n) {
                uint64_t* keysA = c->sweepKeys.get() + uint64_t(s) * capacity;
                uint32_t* valsA = c->sweepVals.get() + uint64_t(s) * capacity;
                c->entryKeysTmp.reserve(n);
                c->entryValsTmp.reserve(n);
                const int bucketRange[1][2] = {{32, 32 + int(log2BucketCount)}};
                const bool inTmp = radixSort<true>(keysA, c->entryKeysTmp.get(), valsA, c->entryValsTmp.get(), n,
                                                   bucketRange, 1, c->sortWs, st);
                const uint64_t* keys = inTmp ? c->entryKeysTmp.get() : keysA;
                const uint32_t* vals = inTmp ? c->entryValsTmp.get() : valsA;

                buildSegments(c, keys, n, 32);

                // Count pass (also per-read statistics), scan, emit pass.
                c->countsBuf.reserve(n);
                unsigned long long* pairTotal = c->scalars.get() + 40;
                SHB_CUDA(cudaMemsetAsync(pairTotal, 0, sizeof(unsigned long long), st));
                SHB_LAUNCH(bucketPairsKernel<false>, ceilDiv(n, 256), 256, 0, st, keys, vals, n,
                           (const uint32_t*)c->flagsBuf.get(), (const uint32_t*)c->indexBuf.get(),
                           (const uint32_t*)c->segStartBuf.get(), p.minBucketSize, p.maxBucketSize,
                           c->stats.get(), pairTotal, c->countsBuf.get(), (uint64_t*)nullptr);
                c->scanWs.reserve(scanWorkspaceElements(n));
                exclusiveScan<uint32_t>(c->countsBuf.get(), c->countsBuf.get(), n, (uint32_t*)nullptr, c->scanWs.get(), st);
                // The exact 64-bit total guards the 32-bit offsets.
                const unsigned long long np64 = readScalar<unsigned long long>(pairTotal, st);
                SHB_REQUIRE(np64 < (1ull << 32), SHB_ERR_INVALID,
                            "LowHash0: more than 2^32-1 candidate pair hits in one iteration (maxBucketSize too large).");
                const uint32_t np = uint32_t(np64);
                pairCount += np;
                if(np) {
                    c->pairsA.reserve(np);
                    c->pairsB.reserve(np);
                    SHB_LAUNCH(bucketPairsKernel<true>, ceilDiv(n, 256), 256, 0, st, keys, vals, n,
                               (const uint32_t*)c->flagsBuf.get(), (const uint32_t*)c->indexBuf.get(),
                               (const uint32_t*)c->segStartBuf.get(), p.minBucketSize, p.maxBucketSize,
                               (unsigned long long*)nullptr, (unsigned long long*)nullptr, c->countsBuf.get(), c->pairsA.get());
                    const int pairRanges[2][2] = {{0, int(readBits) + 1}, {32, 32 + int(readBits)}};
                    const bool inB = radixSort<false>(c->pairsA.get(), c->pairsB.get(), nullptr, nullptr, np,
                                                      pairRanges, 2, c->sortWs, st);
                    const uint64_t* sortedPairs = inB ? c->pairsB.get() : c->pairsA.get();
                    const uint32_t numUnique = buildSegments(c, sortedPairs, np, 0);
                    accReserve(c, acc, acc.count + numUnique);
                    SHB_LAUNCH(uniqueCountsKernel, ceilDiv(numUnique, 256), 256, 0, st, sortedPairs,
                               (const uint32_t*)c->segStartBuf.get(), numUnique,
                               accKeys(c, acc) + acc.count, accVals(c, acc) + acc.count);
                    acc.count += numUnique;
                }
            }
            if(perIteration) {
                mergeAccumulator(c, acc, readBits);
                highFrequency = countHighFrequency(c, acc, p.minFrequency, false);
                if(iterSummary && iteration < maxIterSummary) {
                    iterSummary[2 * iteration] = highFrequency;
                    iterSummary[2 * iteration + 1] = acc.count;
                }
            } else if(acc.count > (1ull << 30)) {
                mergeAccumulator(c, acc, readBits);     // keep the deferred accumulator below 2^32 items
            }
        }
    }

    // ---- final merge + emission, src/LowHash0.cpp:204-214 ----------------------------------------
    if(!perIteration) mergeAccumulator(c, acc, readBits);
    const uint64_t nOut = countHighFrequency(c, acc, p.minFrequency, true);
    void* host = malloc(nOut ? nOut * 12 : 1);
    SHB_REQUIRE(host != nullptr, SHB_ERR_OOM, "Out of host memory for the alignment candidates.");
    if(nOut) {
        c->candidatesDev.reserve(3 * nOut);
        SHB_LAUNCH(emitCandidatesKernel, ceilDiv(acc.count, 256), 256, 0, st, (const uint64_t*)accKeys(c, acc),
                   (const uint32_t*)c->flagsBuf.get(), (const uint32_t*)c->indexBuf.get(), uint32_t(acc.count),
                   c->candidatesDev.get());
        SHB_CUDA(cudaMemcpyAsync(host, c->candidatesDev.get(), nOut * 12, cudaMemcpyDeviceToHost, st));
    }
    if(statsOut) {
        SHB_CUDA(cudaMemcpyAsync(statsOut, c->stats.get(), 3 * R * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    }
    SHB_CUDA(cudaEventRecord(totalTimer.b, st));
    SHB_CUDA(cudaStreamSynchronize(st));
    float totalMs = 0.f;
    SHB_CUDA(cudaEventElapsedTime(&totalMs, totalTimer.a, totalTimer.b));

    *candidatesOut = host;
    *candidateCountOut = nOut;
    if(result) {
        result->iterations = iteration;
        result->log2BucketCount = log2BucketCount;
        result->lowHashCount = lowHashCount;
        result->pairCount = pairCount;
        result->candidateCount = nOut;
        result->sweepMs = sweepMs;
        result->totalMs = totalMs;
        result->sweepLaunches = sweepLaunches;
        result->kernelLaunches = g_launchCount;
    }
}

} // namespace shb
