// computeAlignments on the GPU: host orchestration (src/AssemblerAlign.cpp:208-495 of chanzuckerberg/shasta).
#include "context.cuh"
#include "align_kernels.cuh"

#include <algorithm>
#include <cstring>
#include <limits>
#include <vector>

namespace shb {

extern thread_local uint64_t g_launchCount;

namespace {

template<class T> T readBack(const T* dev, cudaStream_t st)
{
    T v;
    SHB_CUDA(cudaMemcpyAsync(&v, dev, sizeof(T), cudaMemcpyDeviceToHost, st));
    SHB_CUDA(cudaStreamSynchronize(st));
    return v;
}

struct Events {
    cudaEvent_t a = nullptr, b = nullptr;
    Events() { cudaEventCreate(&a); cudaEventCreate(&b); }
    ~Events() { cudaEventDestroy(a); cudaEventDestroy(b); }
};

// Band-width classes: one launch per class so that shared memory is sized for the class.
const uint32_t kClassLimits[] = {64, 128, 256, 512, 1024, 2048, 4096, 8192, 16384};
constexpr int kClassCount = 9;

uint32_t warpsForClass(uint32_t wMax)
{
    // 3 * (wMax + 1) ints per warp; keep a block under ~200 KB of shared memory.
    const uint64_t perWarp = 3ull * (wMax + 1) * 4;
    uint32_t w = uint32_t(std::min<uint64_t>(kDpMaxWarpsPerBlock, (200ull * 1024) / perWarp));
    return w ? w : 1;
}

} // namespace

// Downsampled marker CSR for method 3 (cached in the context per (k, downsamplingFactor)).
struct DownsampledMarkers {
    DeviceBuffer<uint64_t> dsToc;
    DeviceBuffer<uint32_t> dsKmer, dsOrdinal;
    uint64_t total = 0;
    uint32_t maxRow = 0;
    uint32_t k = 0;
    double factor = -1.;
    const uint32_t* forKmerIds = nullptr;
};

static DownsampledMarkers& downsampled(shb_context* c)
{
    if(!c->alignCache) c->alignCache = new DownsampledMarkers();
    return *static_cast<DownsampledMarkers*>(c->alignCache);
}
void destroyAlignCache(shb_context* c)
{
    if(c->alignCache) { delete static_cast<DownsampledMarkers*>(c->alignCache); c->alignCache = nullptr; }
}

static void buildDownsampled(shb_context* c, uint32_t k, double factor)
{
    DownsampledMarkers& ds = downsampled(c);
    if(ds.k == k && ds.factor == factor && ds.forKmerIds == c->kmerIds && ds.dsToc.get()) return;
    cudaStream_t st = c->stream;
    const uint64_t M = c->localMarkerCount;
    const uint32_t rows = uint32_t(2 * c->readCountTotal);
    // src/AssemblerAlign3.cpp:71-72
    const uint32_t hashThreshold = uint32_t(factor * double(std::numeric_limits<uint32_t>::max()));
    ds.dsToc.reserve(uint64_t(rows) + 1);
    const uint32_t chunk = 1u << 27;
    c->flagsBuf.reserve(std::min<uint64_t>(chunk, M) + 1);
    c->indexBuf.reserve(std::min<uint64_t>(chunk, M) + 1);
    c->scanWs.reserve(scanWorkspaceElements(chunk));
    c->scalars.reserve(64);
    uint32_t* totalDev = reinterpret_cast<uint32_t*>(c->scalars.get() + 32);
    // Pass 1 sizes the output exactly; pass 2 compacts.
    std::vector<uint64_t> chunkBase;
    uint64_t total = 0;
    for(int pass = 0; pass < 2; pass++) {
        if(pass == 1) { ds.dsKmer.reserve(total + 1); ds.dsOrdinal.reserve(total + 1); }
        uint64_t running = 0;
        size_t ci = 0;
        for(uint64_t begin = 0; begin < M || (M == 0 && begin == 0); begin += chunk, ci++) {
            const uint32_t n = uint32_t(std::min<uint64_t>(chunk, M - begin));
            if(n == 0) break;
            SHB_LAUNCH(downsampleFlagsKernel, ceilDiv(n, 256), 256, 0, st, c->kmerIds, begin, n, k, hashThreshold, c->flagsBuf.get());
            exclusiveScan<uint32_t>(c->flagsBuf.get(), c->indexBuf.get(), n, totalDev, c->scanWs.get(), st);
            const uint32_t t = readBack<uint32_t>(totalDev, st);
            if(pass == 1) {
                SHB_LAUNCH(downsampleCompactKernel, ceilDiv(n, 256), 256, 0, st, c->kmerIds, begin, n,
                           (const uint32_t*)c->flagsBuf.get(), (const uint32_t*)c->indexBuf.get(),
                           (const uint64_t*)c->toc.get(), rows, running, ds.dsKmer.get(), ds.dsOrdinal.get());
                SHB_LAUNCH(downsampleTocKernel, ceilDiv(uint64_t(rows) + 1, 256), 256, 0, st, (const uint64_t*)c->toc.get(), rows,
                           begin, n, (const uint32_t*)c->indexBuf.get(), running, running + t, ds.dsToc.get());
            }
            running += t;
        }
        total = running;
    }
    if(M == 0) SHB_CUDA(cudaMemsetAsync(ds.dsToc.get(), 0, (uint64_t(rows) + 1) * 8, st));
    // Longest downsampled row (bounds the stage-1 band classes).
    std::vector<uint64_t> hostToc(uint64_t(rows) + 1);
    SHB_CUDA(cudaMemcpyAsync(hostToc.data(), ds.dsToc.get(), (uint64_t(rows) + 1) * 8, cudaMemcpyDeviceToHost, st));
    SHB_CUDA(cudaStreamSynchronize(st));
    uint32_t maxRow = 0;
    for(uint32_t r = 0; r < rows; r++) maxRow = std::max<uint32_t>(maxRow, uint32_t(hostToc[r+1] - hostToc[r]));
    ds.total = total; ds.maxRow = maxRow; ds.k = k; ds.factor = factor; ds.forKmerIds = c->kmerIds;
}


void computeAlignments(shb_context* c, const void* candidatesHost, uint64_t n, const shb_align_options& o,
                       void** alignmentDataOut, uint64_t* alignmentCountOut,
                       uint64_t** compressedTocOut, uint8_t** compressedDataOut, shb_align_result* result)
{
    SHB_REQUIRE(c->haveMarkers, SHB_ERR_STATE, "Markers are not accessible.");
    SHB_REQUIRE(c->readBegin == 0 && c->readEnd == c->readCountTotal, SHB_ERR_STATE,
                "computeAlignments needs the markers of all reads on this GPU.");
    SHB_REQUIRE(o.alignMethod == 3, SHB_ERR_INVALID, "Only Align.alignMethod 3 is implemented in this build.");
    SHB_REQUIRE(o.gapScore <= 0, SHB_ERR_INVALID, "Align.gapScore must not be positive.");
    SHB_REQUIRE(o.k >= 1 && o.k <= 16, SHB_ERR_INVALID, "Invalid k.");
    SHB_REQUIRE(n == 0 || candidatesHost != nullptr, SHB_ERR_INVALID, "Null candidates.");
    SHB_CUDA(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    g_launchCount = 0;
    Events totalEv, dpEv;
    SHB_CUDA(cudaEventRecord(totalEv.a, st));
    double dpMs = 0.;

    // Validate candidates on the host (src/AssemblerAlign.cpp:378 asserts readIds[0] < readIds[1]).
    const uint32_t* cand = static_cast<const uint32_t*>(candidatesHost);
    for(uint64_t i = 0; i < n; i++) {
        SHB_REQUIRE(cand[3*i] < cand[3*i+1] && cand[3*i+1] < c->readCountTotal, SHB_ERR_INVALID, "Invalid alignment candidate.");
    }

    buildDownsampled(c, o.k, o.downsamplingFactor);
    DownsampledMarkers& ds = downsampled(c);

    const DpScores scores{o.matchScore, o.mismatchScore, o.gapScore};
    FilterOptions fo;
    fo.minAlignedMarkerCount = uint64_t(o.minAlignedMarkerCount); fo.maxSkip = uint64_t(o.maxSkip);
    fo.maxDrift = uint64_t(o.maxDrift); fo.maxTrim = uint64_t(o.maxTrim);
    fo.minAlignedFraction = o.minAlignedFraction; fo.suppressContainments = o.suppressContainments ? 1u : 0u;

    const uint32_t batchMax = 32768;
    DeviceBuffer<uint32_t> dCand, counts, infoWords, keep, keepIndex, cbytes32, records;
    DeviceBuffer<DpJob> jobs1, jobs2;
    DeviceBuffer<unsigned long long> tw, twOff, outCnt, outOff, cbytes, cbytesOff, scanWs64, ctoc;
    DeviceBuffer<uint32_t> trace;
    DeviceBuffer<uint2> ordinals;
    DeviceBuffer<uint8_t> cdata;
    dCand.reserve(3ull * batchMax); counts.reserve(batchMax); infoWords.reserve(13ull * batchMax);
    keep.reserve(batchMax); keepIndex.reserve(batchMax); cbytes32.reserve(batchMax);
    jobs1.reserve(batchMax); jobs2.reserve(batchMax);
    tw.reserve(batchMax); twOff.reserve(batchMax); outCnt.reserve(batchMax); outOff.reserve(batchMax);
    cbytes.reserve(batchMax); cbytesOff.reserve(batchMax);
    scanWs64.reserve(scanWorkspaceElements(batchMax));
    c->scanWs.reserve(scanWorkspaceElements(batchMax));
    c->scalars.reserve(64);
    unsigned long long* total64 = c->scalars.get() + 48;
    uint32_t* total32 = reinterpret_cast<uint32_t*>(c->scalars.get() + 32);

    std::vector<uint32_t> hostRecords;
    std::vector<uint64_t> hostToc;
    std::vector<uint8_t> hostData;
    hostToc.push_back(0);
    uint64_t skipped = 0, dpCells = 0;
    const uint32_t maxStage1Width = 2 * ds.maxRow + 2;
    SHB_REQUIRE(maxStage1Width <= 16384, SHB_ERR_INVALID, "Downsampled reads are too long for the stage-1 kernel (limit 8191 markers).");
    const uint32_t maxStage2Width = uint32_t(std::max(0, o.maxBand)) + 2;
    SHB_REQUIRE(maxStage2Width <= 16384, SHB_ERR_INVALID, "Align.maxBand too large for this implementation (limit 16382).");

    for(uint64_t begin = 0; begin < n; begin += batchMax) {
        const uint32_t nb = uint32_t(std::min<uint64_t>(batchMax, n - begin));
        SHB_CUDA(cudaMemcpyAsync(dCand.get(), cand + 3 * begin, 12ull * nb, cudaMemcpyHostToDevice, st));
        SHB_LAUNCH(method3SetupKernel, ceilDiv(nb, 256), 256, 0, st, (const uint32_t*)dCand.get(), nb,
                   (const uint64_t*)c->toc.get(), (const uint64_t*)ds.dsToc.get(), jobs1.get(), jobs2.get(), tw.get(), outCnt.get());
        // Stage 1 trace scratch.
        exclusiveScan<unsigned long long>(tw.get(), twOff.get(), nb, total64, scanWs64.get(), st);
        const unsigned long long traceWords1 = readBack<unsigned long long>(total64, st);
        trace.reserve(traceWords1 + 1);
        SHB_LAUNCH(setTraceOffsetsKernel, ceilDiv(nb, 256), 256, 0, st, jobs1.get(), nb, (const unsigned long long*)twOff.get(),
                   (const unsigned long long*)nullptr);
        Method3Args g1;
        g1.candidates = dCand.get(); g1.candidateBegin = begin; g1.n = nb;
        g1.toc = c->toc.get(); g1.dsToc = ds.dsToc.get(); g1.dsKmer = ds.dsKmer.get(); g1.dsOrdinal = ds.dsOrdinal.get();
        g1.scores = scores; g1.bandExtend = o.bandExtend; g1.maxBand = o.maxBand;
        SHB_CUDA(cudaEventRecord(dpEv.a, st));
        {
            uint32_t wMin = 0;
            for(int k = 0; k < kClassCount; k++) {
                const uint32_t wMax = kClassLimits[k];
                const uint32_t warps = warpsForClass(wMax);
                const size_t smem = size_t(warps) * 3 * (wMax + 1) * 4;
                g1.wMin = wMin; g1.wMax = wMax;
                SHB_CUDA(cudaFuncSetAttribute(method3Stage1Kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
                SHB_LAUNCH(method3Stage1Kernel, ceilDiv(nb, warps), warps * 32, smem, st, g1, jobs1.get(), trace.get(), jobs2.get());
                wMin = wMax;
                if(wMax >= maxStage1Width) break;
            }
        }
        // Stage 2 scratch: trace words and ordinal slots.
        SHB_LAUNCH(stage2TraceWordsKernel, ceilDiv(nb, 256), 256, 0, st, (const DpJob*)jobs2.get(), nb, tw.get());
        exclusiveScan<unsigned long long>(tw.get(), twOff.get(), nb, total64, scanWs64.get(), st);
        const unsigned long long traceWords2 = readBack<unsigned long long>(total64, st);
        exclusiveScan<unsigned long long>(outCnt.get(), outOff.get(), nb, total64, scanWs64.get(), st);
        const unsigned long long ordinalSlots = readBack<unsigned long long>(total64, st);
        trace.reserve(traceWords2 + 1);
        ordinals.reserve(ordinalSlots + 1);
        dpCells += 16ull * (traceWords1 + traceWords2);
        SHB_LAUNCH(setTraceOffsetsKernel, ceilDiv(nb, 256), 256, 0, st, jobs2.get(), nb, (const unsigned long long*)twOff.get(),
                   (const unsigned long long*)outOff.get());
        SHB_CUDA(cudaMemsetAsync(counts.get(), 0, 4ull * nb, st));
        BandedArgs g2;
        g2.n = nb; g2.kmerIds = c->kmerIds; g2.scores = scores;
        {
            uint32_t wMin = 0;
            for(int k = 0; k < kClassCount; k++) {
                const uint32_t wMax = kClassLimits[k];
                const uint32_t warps = warpsForClass(wMax);
                const size_t smem = size_t(warps) * 3 * (wMax + 1) * 4;
                g2.wMin = wMin; g2.wMax = wMax;
                SHB_CUDA(cudaFuncSetAttribute(bandedAlignKernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
                SHB_LAUNCH(bandedAlignKernel, ceilDiv(nb, warps), warps * 32, smem, st, g2, (const DpJob*)jobs2.get(), trace.get(),
                           ordinals.get(), counts.get());
                wMin = wMax;
                if(wMax >= maxStage2Width) break;
            }
        }
        SHB_CUDA(cudaEventRecord(dpEv.b, st));
        // Epilogue: info, filters, compressed sizes; then compaction.
        SHB_LAUNCH(alignmentInfoKernel, ceilDiv(nb, 128), 128, 0, st, nb, (const DpJob*)jobs2.get(), (const uint2*)ordinals.get(),
                   (const uint32_t*)counts.get(), fo, infoWords.get(), keep.get(), cbytes32.get());
        exclusiveScan<uint32_t>(keep.get(), keepIndex.get(), nb, total32, c->scanWs.get(), st);
        const uint32_t kept = readBack<uint32_t>(total32, st);
        SHB_LAUNCH(widenBytesKernel, ceilDiv(nb, 256), 256, 0, st, (const uint32_t*)cbytes32.get(), nb, cbytes.get());
        exclusiveScan<unsigned long long>(cbytes.get(), cbytesOff.get(), nb, total64, scanWs64.get(), st);
        const unsigned long long bytes = readBack<unsigned long long>(total64, st);
        {
            float ms = 0.f;
            SHB_CUDA(cudaEventElapsedTime(&ms, dpEv.a, dpEv.b));
            dpMs += ms;
        }
        // Count the candidates the reference would skip with a logged exception.
        {
            std::vector<DpJob> hj(nb);
            SHB_CUDA(cudaMemcpyAsync(hj.data(), jobs2.get(), sizeof(DpJob) * nb, cudaMemcpyDeviceToHost, st));
            SHB_CUDA(cudaStreamSynchronize(st));
            for(const DpJob& j : hj) if(j.state == kStateSkipped) skipped++;
        }
        if(kept) {
            records.reserve(16ull * kept); ctoc.reserve(kept); cdata.reserve(bytes + 16);
            SHB_LAUNCH(alignmentWriteKernel, ceilDiv(nb, 128), 128, 0, st, nb, (const uint32_t*)dCand.get(), (const DpJob*)jobs2.get(),
                       (const uint2*)ordinals.get(), (const uint32_t*)counts.get(), (const uint32_t*)infoWords.get(),
                       (const uint32_t*)keep.get(), (const uint32_t*)keepIndex.get(), (const unsigned long long*)cbytesOff.get(),
                       0ull, 0ull, records.get(), ctoc.get(), cdata.get());
            const size_t r0 = hostRecords.size(), b0 = hostData.size();
            hostRecords.resize(r0 + 16ull * kept);
            hostData.resize(b0 + bytes);
            std::vector<unsigned long long> tocBatch(kept);
            SHB_CUDA(cudaMemcpyAsync(hostRecords.data() + r0, records.get(), 64ull * kept, cudaMemcpyDeviceToHost, st));
            SHB_CUDA(cudaMemcpyAsync(tocBatch.data(), ctoc.get(), 8ull * kept, cudaMemcpyDeviceToHost, st));
            if(bytes) SHB_CUDA(cudaMemcpyAsync(hostData.data() + b0, cdata.get(), bytes, cudaMemcpyDeviceToHost, st));
            SHB_CUDA(cudaStreamSynchronize(st));
            for(uint32_t i = 0; i < kept; i++) {
                // toc entry = start of alignment i; the end is the next start (or the batch end).
                const uint64_t end = (i + 1 < kept) ? tocBatch[i + 1] : bytes;
                hostToc.push_back(b0 + end);
            }
        }
    }

    SHB_CUDA(cudaEventRecord(totalEv.b, st));
    SHB_CUDA(cudaStreamSynchronize(st));
    float totalMs = 0.f;
    SHB_CUDA(cudaEventElapsedTime(&totalMs, totalEv.a, totalEv.b));

    const uint64_t count = hostRecords.size() / 16;
    void* recOut = malloc(count ? 64 * count : 1);
    uint64_t* tocOut = (uint64_t*)malloc(8 * (count + 1));
    uint8_t* dataOut = (uint8_t*)malloc(hostData.size() ? hostData.size() : 1);
    SHB_REQUIRE(recOut && tocOut && dataOut, SHB_ERR_OOM, "Out of host memory for the alignments.");
    if(count) memcpy(recOut, hostRecords.data(), 64 * count);
    memcpy(tocOut, hostToc.data(), 8 * (count + 1));
    if(!hostData.empty()) memcpy(dataOut, hostData.data(), hostData.size());
    *alignmentDataOut = recOut; *alignmentCountOut = count; *compressedTocOut = tocOut; *compressedDataOut = dataOut;
    if(result) {
        result->candidateCount = n; result->alignmentCount = count; result->skippedCount = skipped;
        result->dpCells = dpCells; result->dpMs = dpMs; result->totalMs = totalMs; result->kernelLaunches = g_launchCount;
    }
}

} // namespace shb
