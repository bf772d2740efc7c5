// computeAlignments on the GPU: host orchestration (src/AssemblerAlign.cpp:208-495 of chanzuckerberg/shasta).
// Method 3: stage 1 (downsampled, unbanded) -> band -> stage 2 (banded, all markers) -> epilogue.
// Method 4: Align4 front end (cells/components) -> one banded DP per component -> best -> epilogue.
#include "context.cuh"
#include "align_kernels.cuh"
#include "hostpool.cuh"

#include <algorithm>
#include <chrono>
#include <cstring>
#include <sys/mman.h>
#include <thread>
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

// Band-width classes, one launch per class. Up to 1024 offsets the register-resident wavefront kernel runs with
// C = offsets/64 per sub-chunk (no shared memory); wider bands use the shared-memory scan kernel (C = 0).
struct DpClass { uint32_t wMax; int c; };
const DpClass kClasses[] = {{64, 1}, {128, 2}, {192, 3}, {256, 4}, {384, 6}, {512, 8}, {768, 12}, {1024, 16},
                            {2048, 0}, {4096, 0}, {8192, 0}, {16384, 0}};
constexpr int kClassCount = 12;
constexpr uint32_t kMaxBandWidth = 16384;

uint32_t warpsForClass(const DpClass& k)
{
    if(k.c > 0) return kDpMaxWarpsPerBlock;
    // scan kernel: 3 * (wMax + 1) ints per warp; keep a block under ~200 KB of shared memory.
    const uint64_t perWarp = 3ull * (k.wMax + 1) * 4;
    const uint32_t w = uint32_t(std::min<uint64_t>(kDpMaxWarpsPerBlock, (200ull * 1024) / perWarp));
    return w ? w : 1;
}
size_t smemForClass(const DpClass& k, uint32_t warps) { return k.c > 0 ? 0 : size_t(warps) * 3 * (k.wMax + 1) * 4; }

template<class... Args> void launchStage1(const DpClass& k, uint32_t blocks, uint32_t threads, size_t smem, cudaStream_t st, Args... args)
{
    switch(k.c) {
    case 1: SHB_LAUNCH(method3Stage1Kernel<1>, blocks, threads, smem, st, args...); break;
    case 2: SHB_LAUNCH(method3Stage1Kernel<2>, blocks, threads, smem, st, args...); break;
    case 3: SHB_LAUNCH(method3Stage1Kernel<3>, blocks, threads, smem, st, args...); break;
    case 4: SHB_LAUNCH(method3Stage1Kernel<4>, blocks, threads, smem, st, args...); break;
    case 6: SHB_LAUNCH(method3Stage1Kernel<6>, blocks, threads, smem, st, args...); break;
    case 8: SHB_LAUNCH(method3Stage1Kernel<8>, blocks, threads, smem, st, args...); break;
    case 12: SHB_LAUNCH(method3Stage1Kernel<12>, blocks, threads, smem, st, args...); break;
    case 16: SHB_LAUNCH(method3Stage1Kernel<16>, blocks, threads, smem, st, args...); break;
    default:
        SHB_CUDA(cudaFuncSetAttribute(method3Stage1Kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
        SHB_LAUNCH(method3Stage1Kernel<0>, blocks, threads, smem, st, args...); break;
    }
}
// Stage 1 of method 3 for downsampled reads of at most 32*R markers (rows per lane R).
const int kForwardRows[] = {2, 4, 6, 8, 12, 16};
constexpr int kForwardClassCount = 6;        // row limits 64 .. 512 = kClasses[0..5].wMax
template<class... Args> void launchStage1Forward(int rows, uint32_t blocks, uint32_t threads, cudaStream_t st, Args... args)
{
    switch(rows) {
    case 2: SHB_LAUNCH(method3Stage1ForwardKernel<2>, blocks, threads, 0, st, args...); break;
    case 4: SHB_LAUNCH(method3Stage1ForwardKernel<4>, blocks, threads, 0, st, args...); break;
    case 6: SHB_LAUNCH(method3Stage1ForwardKernel<6>, blocks, threads, 0, st, args...); break;
    case 8: SHB_LAUNCH(method3Stage1ForwardKernel<8>, blocks, threads, 0, st, args...); break;
    case 12: SHB_LAUNCH(method3Stage1ForwardKernel<12>, blocks, threads, 0, st, args...); break;
    default: SHB_LAUNCH(method3Stage1ForwardKernel<16>, blocks, threads, 0, st, args...); break;
    }
}
template<class... Args> void launchBanded(const DpClass& k, uint32_t blocks, uint32_t threads, size_t smem, cudaStream_t st, Args... args)
{
    switch(k.c) {
    case 1: SHB_LAUNCH(bandedAlignKernel<1>, blocks, threads, smem, st, args...); break;
    case 2: SHB_LAUNCH(bandedAlignKernel<2>, blocks, threads, smem, st, args...); break;
    case 3: SHB_LAUNCH(bandedAlignKernel<3>, blocks, threads, smem, st, args...); break;
    case 4: SHB_LAUNCH(bandedAlignKernel<4>, blocks, threads, smem, st, args...); break;
    case 6: SHB_LAUNCH(bandedAlignKernel<6>, blocks, threads, smem, st, args...); break;
    case 8: SHB_LAUNCH(bandedAlignKernel<8>, blocks, threads, smem, st, args...); break;
    case 12: SHB_LAUNCH(bandedAlignKernel<12>, blocks, threads, smem, st, args...); break;
    case 16: SHB_LAUNCH(bandedAlignKernel<16>, blocks, threads, smem, st, args...); break;
    default:
        SHB_CUDA(cudaFuncSetAttribute(bandedAlignKernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
        SHB_LAUNCH(bandedAlignKernel<0>, blocks, threads, smem, st, args...); break;
    }
}

// Buffers shared by both methods for one batch of candidates.
struct Batch {
    DeviceBuffer<uint32_t> cand, counts, infoWords, jobKeep, jobBytes, keep, keepIndex, bytes32, selected, records;
    DeviceBuffer<DpJob> jobs1, jobs;
    DeviceBuffer<unsigned long long> tw, twOff, outCnt, outOff, bytes64, bytesOff, scanWs64, ctoc;
    DeviceBuffer<uint32_t> trace;
    DeviceBuffer<uint2> ordinals, runs;
    DeviceBuffer<int2> endCells;
    DeviceBuffer<uint32_t> runCounts;
    DeviceBuffer<uint8_t> cdata;
    // method 4
    DeviceBuffer<unsigned long long> cellCnt, cellOff;
    DeviceBuffer<uint32_t> gridCounts, gridAux, gridList, componentCount, jobOffsets;
    DeviceBuffer<uint8_t> gridFlags;
    DeviceBuffer<int32_t> gridBands;
    // band-class ordering of the DP jobs
    DeviceBuffer<uint32_t> classLimits, orderValsA, orderValsB;
    DeviceBuffer<uint64_t> orderKeysA, orderKeysB;
    const uint32_t* order = nullptr;
};

// Derived per-marker data cached in the context (per marker generation).
struct AlignCache {
    // method 3: downsampled marker CSR
    DeviceBuffer<uint64_t> dsToc;
    DeviceBuffer<uint32_t> dsKmer, dsOrdinal;
    uint32_t dsMaxRow = 0, dsK = 0;
    double dsFactor = -1.;
    uint64_t dsGeneration = ~0ull;
    // method 4: markers sorted by k-mer id within each oriented read
    DeviceBuffer<uint32_t> sortedKmer, sortedOrdinal;
    uint64_t sortedGeneration = ~0ull;
    uint64_t lengthCheckGeneration = ~0ull;
    // per-batch scratch and the device-side result accumulation: kept across calls so that a steady-state call does
    // not allocate or free device memory
    Batch batch;
    DeviceBuffer<uint32_t> outRecords;
    DeviceBuffer<unsigned long long> outToc;
    DeviceBuffer<uint8_t> outData;
    // side streams: the band classes of one batch are independent launches (disjoint jobs and scratch), so the few
    // long jobs of the wide classes run beside the big narrow-band launch instead of after it
    static constexpr int kSideStreams = 3;
    cudaStream_t side[kSideStreams] = {nullptr, nullptr, nullptr};
    cudaEvent_t forkEv = nullptr, joinEv[kSideStreams] = {nullptr, nullptr, nullptr};
    // high-priority stream for the short latency-bound kernels (traceback, filter) that follow each DP chunk: their
    // blocks are scheduled ahead of the pending blocks of the other chunks' DP kernels
    static constexpr int kHiStreams = 4;
    cudaStream_t hiStream[kHiStreams] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t hiJoinEv[kHiStreams] = {nullptr, nullptr, nullptr, nullptr};
    std::vector<cudaEvent_t> unitEvents;
    ~AlignCache()
    {
        for(int i = 0; i < kHiStreams; i++) { if(hiStream[i]) cudaStreamDestroy(hiStream[i]); if(hiJoinEv[i]) cudaEventDestroy(hiJoinEv[i]); }
        for(cudaEvent_t e : unitEvents) cudaEventDestroy(e);
        for(int i = 0; i < kSideStreams; i++) { if(side[i]) cudaStreamDestroy(side[i]); if(joinEv[i]) cudaEventDestroy(joinEv[i]); }
        if(forkEv) cudaEventDestroy(forkEv);
    }
};

AlignCache& cache(shb_context* c)
{
    if(!c->alignCache) c->alignCache = new AlignCache();
    return *static_cast<AlignCache*>(c->alignCache);
}

// Runs launch(k, count, offset, stream) for every non-empty band class, cut into chunks of at most chunkMax jobs:
// the widest classes first, round-robin over the side streams and the main stream, so that independent launches
// (disjoint jobs and scratch) overlap: the few long jobs of the wide classes run beside the big narrow-band launch,
// and the latency-bound traceback of one chunk runs beside the issue-bound DP of the next. The main stream continues
// after all of them.
template<class F> void forEachClassConcurrently(shb_context* c, const std::vector<uint64_t>& classCounts, uint32_t chunkMax, F launch)
{
    AlignCache& ac = cache(c);
    cudaStream_t st = c->stream;
    if(!ac.forkEv) {
        SHB_CUDA(cudaEventCreateWithFlags(&ac.forkEv, cudaEventDisableTiming));
        for(int i = 0; i < AlignCache::kSideStreams; i++) {
            SHB_CUDA(cudaStreamCreateWithFlags(&ac.side[i], cudaStreamNonBlocking));
            SHB_CUDA(cudaEventCreateWithFlags(&ac.joinEv[i], cudaEventDisableTiming));
        }
    }
    struct Unit { int k; uint32_t count; uint64_t offset; };
    std::vector<Unit> units;
    const int classCount = int(classCounts.size());
    std::vector<uint64_t> offsets(classCounts.size(), 0);
    for(int k = 1; k < classCount; k++) offsets[k] = offsets[k-1] + classCounts[k-1];
    for(int k = classCount - 1; k >= 0; k--) {
        const uint64_t count = classCounts[k];
        if(!count) continue;
        const uint64_t chunks = (count + chunkMax - 1) / chunkMax, per = (count + chunks - 1) / chunks;
        for(uint64_t begin = 0; begin < count; begin += per) units.push_back({k, uint32_t(std::min(per, count - begin)), offsets[k] + begin});
    }
    if(units.empty()) return;
    constexpr int kStreams = AlignCache::kSideStreams + 1;      // the last one is the main stream
    bool used[AlignCache::kSideStreams] = {false, false, false};
    if(units.size() > 1) SHB_CUDA(cudaEventRecord(ac.forkEv, st));
    for(size_t u = 0; u < units.size(); u++) {
        // the last unit always goes to the main stream
        const int i = (u + 1 == units.size()) ? kStreams - 1 : int(u % kStreams);
        cudaStream_t s = st;
        if(i < AlignCache::kSideStreams) {
            s = ac.side[i];
            if(!used[i]) { SHB_CUDA(cudaStreamWaitEvent(s, ac.forkEv, 0)); used[i] = true; }
        }
        launch(units[u].k, units[u].count, units[u].offset, s);
    }
    for(int i = 0; i < AlignCache::kSideStreams; i++) {
        if(!used[i]) continue;
        SHB_CUDA(cudaEventRecord(ac.joinEv[i], ac.side[i]));
        SHB_CUDA(cudaStreamWaitEvent(st, ac.joinEv[i], 0));
    }
}

void buildDownsampled(shb_context* c, uint32_t k, double factor)
{
    AlignCache& ds = cache(c);
    if(ds.dsK == k && ds.dsFactor == factor && ds.dsGeneration == c->markerGeneration && ds.dsToc.get()) return;
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
    // scalars: 512 entries, allocated once at context creation
    uint32_t* totalDev = reinterpret_cast<uint32_t*>(c->scalars.get() + 32);
    // Pass 0 sizes the output exactly; pass 1 compacts.
    uint64_t total = 0;
    for(int pass = 0; pass < 2; pass++) {
        if(pass == 1) { ds.dsKmer.reserve(total + 1); ds.dsOrdinal.reserve(total + 1); }
        uint64_t running = 0;
        for(uint64_t begin = 0; begin < M; begin += chunk) {
            const uint32_t n = uint32_t(std::min<uint64_t>(chunk, M - begin));
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
    ds.dsMaxRow = maxRow; ds.dsK = k; ds.dsFactor = factor; ds.dsGeneration = c->markerGeneration;
}

// computeSortedMarkers (src/AssemblerAlign4.cpp:190-261): per oriented read, (kmerId, ordinal) sorted by kmerId.
void buildSortedMarkers(shb_context* c, uint32_t k)
{
    AlignCache& sc = cache(c);
    if(sc.sortedGeneration == c->markerGeneration && sc.sortedKmer.get()) return;
    cudaStream_t st = c->stream;
    const uint64_t M = c->localMarkerCount;
    const uint32_t rows = uint32_t(2 * c->readCountTotal);
    sc.sortedKmer.reserve(M + 1);
    sc.sortedOrdinal.reserve(M + 1);
    const std::vector<uint64_t>& toc = c->tocHost;
    DeviceBuffer<uint64_t> keysA, keysB;
    DeviceBuffer<uint32_t> valsA, valsB;
    const uint64_t chunkLimit = 1ull << 28;
    uint32_t rowBegin = 0;
    while(rowBegin < rows) {
        uint32_t rowEnd = rowBegin + 1;
        while(rowEnd < rows && toc[rowEnd + 1] - toc[rowBegin] <= chunkLimit) rowEnd++;
        const uint64_t markerBegin = toc[rowBegin];
        const uint64_t n64 = toc[rowEnd] - markerBegin;
        SHB_REQUIRE(n64 < (1ull << 32), SHB_ERR_INVALID, "An oriented read has more than 2^32-1 markers.");
        const uint32_t n = uint32_t(n64);
        if(n) {
            keysA.reserve(n); keysB.reserve(n); valsA.reserve(n); valsB.reserve(n);
            SHB_LAUNCH(sortedMarkerKeysKernel, ceilDiv(n, 256), 256, 0, st, c->kmerIds, (const uint64_t*)c->toc.get(),
                       rowBegin, rowEnd, markerBegin, n, keysA.get(), valsA.get());
            uint32_t rowBits = 1;
            while((1ull << rowBits) < uint64_t(rowEnd - rowBegin)) rowBits++;
            const int ranges[2][2] = {{0, int(2 * k)}, {32, 32 + int(rowBits)}};
            const bool inB = radixSort<true>(keysA.get(), keysB.get(), valsA.get(), valsB.get(), n, ranges, 2, c->sortWs, st);
            SHB_LAUNCH(sortedMarkerUnpackKernel, ceilDiv(n, 256), 256, 0, st, (const uint64_t*)(inB ? keysB.get() : keysA.get()), n,
                       sc.sortedKmer.get() + markerBegin);
            SHB_CUDA(cudaMemcpyAsync(sc.sortedOrdinal.get() + markerBegin, inB ? valsB.get() : valsA.get(), 4ull * n,
                                     cudaMemcpyDeviceToDevice, st));
        }
        rowBegin = rowEnd;
    }
    SHB_CUDA(cudaStreamSynchronize(st));
    sc.sortedGeneration = c->markerGeneration;
}

struct DpTotals { unsigned long long traceWords = 0; double ms = 0.; };

uint32_t envCount(const char* name, uint32_t dflt)
{
    const char* v = getenv(name);
    if(!v) return dflt;
    const long x = strtol(v, nullptr, 10);
    return x > 0 ? uint32_t(x) : dflt;
}

// SHB_TRACE only: wall time of the stage-2 sub-phases (each bracketed by stream synchronisation), summed per call.
double g_tracePhaseMs[4] = {0., 0., 0., 0.};

// Host-side phase timing, printed to stderr when SHB_TRACE is set (diagnostics only).
struct PhaseClock {
    bool on = getenv("SHB_TRACE") != nullptr;
    std::chrono::steady_clock::time_point t = std::chrono::steady_clock::now();
    double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    void lap(int phase, cudaStream_t st)
    {
        if(!on) return;
        cudaStreamSynchronize(st);
        const auto now = std::chrono::steady_clock::now();
        acc[phase] += std::chrono::duration<double, std::milli>(now - t).count();
        t = now;
    }
    void report(const char* const* names, int n) const
    {
        if(!on) return;
        fprintf(stderr, "[shb] computeAlignments phases (ms):");
        for(int i = 0; i < n; i++) fprintf(stderr, " %s=%.1f", names[i], acc[i]);
        fprintf(stderr, "\n");
    }
};

void parallelMemcpy(uint8_t* dst, const uint8_t* src, uint64_t n)
{
    constexpr int kThreads = 4;
    if(n < (4ull << 20)) { memcpy(dst, src, n); return; }
    std::thread workers[kThreads - 1];
    const uint64_t part = (n / kThreads + 4095) & ~4095ull;
    for(int t = 1; t < kThreads; t++) {
        const uint64_t off = std::min<uint64_t>(n, part * t), len = std::min<uint64_t>(part, n - off);
        workers[t - 1] = std::thread([=] { if(len) memcpy(dst + off, src + off, len); });
    }
    memcpy(dst, src, std::min<uint64_t>(part, n));
    for(int t = 1; t < kThreads; t++) workers[t - 1].join();
}

// Device -> pageable host copy through two pinned staging buffers: the DMA of chunk k overlaps the host memcpy of
// chunk k-1 (a plain cudaMemcpy into pageable memory serialises the two).
void copyToHostPipelined(shb_context* c, void* dstHost, const void* srcDevice, uint64_t bytes)
{
    if(bytes == 0) return;
    if(HostPool::instance().isPageLocked(dstHost)) {        // recycled, page-locked result block: direct DMA
        SHB_CUDA(cudaMemcpyAsync(dstHost, srcDevice, bytes, cudaMemcpyDeviceToHost, c->stream));
        return;
    }
    constexpr uint64_t kChunk = 32ull << 20;
    if(!c->pinnedStage[0]) {
        SHB_CUDA(cudaHostAlloc(&c->pinnedStage[0], kChunk, cudaHostAllocDefault));
        SHB_CUDA(cudaHostAlloc(&c->pinnedStage[1], kChunk, cudaHostAllocDefault));
        SHB_CUDA(cudaEventCreateWithFlags(&c->stageEvent[0], cudaEventDisableTiming));
        SHB_CUDA(cudaEventCreateWithFlags(&c->stageEvent[1], cudaEventDisableTiming));
    }
    cudaStream_t st = c->stream;
    const uint64_t chunks = (bytes + kChunk - 1) / kChunk;
    for(uint64_t k = 0; k <= chunks; k++) {
        if(k < chunks) {
            const uint64_t off = k * kChunk, n = std::min(kChunk, bytes - off);
            SHB_CUDA(cudaMemcpyAsync(c->pinnedStage[k & 1], static_cast<const uint8_t*>(srcDevice) + off, n, cudaMemcpyDeviceToHost, st));
            SHB_CUDA(cudaEventRecord(c->stageEvent[k & 1], st));
        }
        if(k > 0) {
            const uint64_t j = k - 1, off = j * kChunk, n = std::min(kChunk, bytes - off);
            SHB_CUDA(cudaEventSynchronize(c->stageEvent[j & 1]));
            parallelMemcpy(static_cast<uint8_t*>(dstHost) + off, static_cast<const uint8_t*>(c->pinnedStage[j & 1]), n);
        }
    }
}

// Groups the runnable jobs by band class (longest first inside a class). Returns per-class counts; b.order holds the
// job indices, class after class.
void buildClassOrder(shb_context* c, Batch& b, const DpJob* jobs, uint32_t nJobs, std::vector<uint64_t>& classCounts,
                     uint32_t forwardClasses = 0)
{
    cudaStream_t st = c->stream;
    classCounts.assign(kClassCount + forwardClasses, 0);
    if(nJobs == 0) return;
    if(!b.classLimits.get()) {
        b.classLimits.reserve(kClassCount);
        uint32_t limits[kClassCount];
        for(int k = 0; k < kClassCount; k++) limits[k] = kClasses[k].wMax;
        SHB_CUDA(cudaMemcpyAsync(b.classLimits.get(), limits, sizeof(limits), cudaMemcpyHostToDevice, st));
        SHB_CUDA(cudaStreamSynchronize(st));
    }
    b.orderKeysA.reserve(nJobs); b.orderKeysB.reserve(nJobs); b.orderValsA.reserve(nJobs); b.orderValsB.reserve(nJobs);
    SHB_LAUNCH(dpClassKeysKernel, ceilDiv(nJobs, 256), 256, 0, st, jobs, nJobs, (const uint32_t*)b.classLimits.get(), uint32_t(kClassCount),
               forwardClasses, b.orderKeysA.get(), b.orderValsA.get());
    const int ranges[1][2] = {{0, kDpLengthKeyBits + kDpClassKeyBits}};
    const bool inB = radixSort<true>(b.orderKeysA.get(), b.orderKeysB.get(), b.orderValsA.get(), b.orderValsB.get(), nJobs, ranges, 1, c->sortWs, st);
    b.order = inB ? b.orderValsB.get() : b.orderValsA.get();
    // scalars: 512 entries, allocated once at context creation
    unsigned long long* dCounts = c->scalars.get() + 64;
    SHB_CUDA(cudaMemsetAsync(dCounts, 0, 256 * sizeof(unsigned long long), st));
    SHB_LAUNCH(digitCountKernel, ceilDiv(nJobs, 256), 256, 0, st, (const uint64_t*)(inB ? b.orderKeysB.get() : b.orderKeysA.get()), nJobs, kDpLengthKeyBits, kDpClassNone, dCounts);
    unsigned long long h[256];
    SHB_CUDA(cudaMemcpyAsync(h, dCounts, sizeof(h), cudaMemcpyDeviceToHost, st));
    SHB_CUDA(cudaStreamSynchronize(st));
    for(size_t k = 0; k < classCounts.size(); k++) classCounts[k] = h[k];
}

// Scratch offsets + the banded DP + traceback for nJobs jobs whose lo/hi/state are set.
void runBandedJobs(shb_context* c, Batch& b, uint32_t nJobs, const uint32_t* sequences, DpScores scores,
                   Events& ev, DpTotals& totals)
{
    cudaStream_t st = c->stream;
    static const bool traceOn = getenv("SHB_TRACE") != nullptr;
    auto lap = [&](int k, std::chrono::steady_clock::time_point& t0) {
        if(!traceOn) return;
        cudaStreamSynchronize(st);
        const auto t1 = std::chrono::steady_clock::now();
        g_tracePhaseMs[k] += std::chrono::duration<double, std::milli>(t1 - t0).count();
        t0 = t1;
    };
    auto t0 = std::chrono::steady_clock::now();
    unsigned long long* total64 = c->scalars.get() + 48;
    b.scanWs64.reserve(scanWorkspaceElements(nJobs));
    b.twOff.reserve(nJobs); b.outOff.reserve(nJobs); b.counts.reserve(nJobs); b.endCells.reserve(nJobs); b.runCounts.reserve(nJobs);
    exclusiveScan<unsigned long long>(b.tw.get(), b.twOff.get(), nJobs, total64, b.scanWs64.get(), st);
    const unsigned long long traceWords = readBack<unsigned long long>(total64, st);
    exclusiveScan<unsigned long long>(b.outCnt.get(), b.outOff.get(), nJobs, total64, b.scanWs64.get(), st);
    const unsigned long long ordinalSlots = readBack<unsigned long long>(total64, st);
    b.trace.reserve(traceWords + 1);
    b.ordinals.reserve(ordinalSlots + 1);
    b.runs.reserve(ordinalSlots + 1);
    totals.traceWords += traceWords;
    SHB_LAUNCH(setTraceOffsetsKernel, ceilDiv(nJobs, 256), 256, 0, st, b.jobs.get(), nJobs, (const unsigned long long*)b.twOff.get(),
               (const unsigned long long*)b.outOff.get());
    SHB_CUDA(cudaMemsetAsync(b.counts.get(), 0, 4ull * nJobs, st));
    std::vector<uint64_t> classCounts;
    buildClassOrder(c, b, b.jobs.get(), nJobs, classCounts);
    BandedArgs g;
    g.kmerIds = sequences; g.scores = scores;
    lap(3, t0);
    SHB_CUDA(cudaEventRecord(ev.a, st));
    // Per chunk: DP (warp per job) on its stream, then traceback (thread per job) and equal-k-mer filter (warp per job)
    // on the high-priority stream.
    AlignCache& ac = cache(c);
    if(!ac.hiStream[0]) {
        int least = 0, greatest = 0;
        SHB_CUDA(cudaDeviceGetStreamPriorityRange(&least, &greatest));
        for(int i = 0; i < AlignCache::kHiStreams; i++) {
            SHB_CUDA(cudaStreamCreateWithPriority(&ac.hiStream[i], cudaStreamNonBlocking, greatest));
            SHB_CUDA(cudaEventCreateWithFlags(&ac.hiJoinEv[i], cudaEventDisableTiming));
        }
    }
    size_t unit = 0;
    forEachClassConcurrently(c, classCounts, envCount("SHB_ALIGN_CHUNK", 32768), [&](int k, uint32_t count, uint64_t offset, cudaStream_t s) {
        const uint32_t warps = warpsForClass(kClasses[k]);
        const size_t smem = smemForClass(kClasses[k], warps);
        BandedArgs gk = g;
        gk.n = count; gk.order = b.order + offset; gk.wMax = kClasses[k].wMax;
        launchBanded(kClasses[k], ceilDiv(count, warps), warps * 32, smem, s, gk, (const DpJob*)b.jobs.get(), b.trace.get(), b.endCells.get());
        if(unit == ac.unitEvents.size()) {
            cudaEvent_t e = nullptr;
            SHB_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            ac.unitEvents.push_back(e);
        }
        SHB_CUDA(cudaEventRecord(ac.unitEvents[unit], s));
        cudaStream_t hs = ac.hiStream[unit % AlignCache::kHiStreams];
        SHB_CUDA(cudaStreamWaitEvent(hs, ac.unitEvents[unit], 0));
        unit++;
        SHB_LAUNCH(tracebackKernel, ceilDiv(count, 128), 128, 0, hs, count, gk.order, (const DpJob*)b.jobs.get(),
                   (const int2*)b.endCells.get(), (const uint32_t*)b.trace.get(), b.runs.get(), b.runCounts.get());
        SHB_LAUNCH(filterStepsKernel, ceilDiv(count, 4), 128, 0, hs, count, gk.order, (const DpJob*)b.jobs.get(), sequences,
                   (const uint2*)b.runs.get(), (const uint32_t*)b.runCounts.get(), b.ordinals.get(), b.counts.get());
    });
    for(size_t i = 0; i < std::min<size_t>(unit, AlignCache::kHiStreams); i++) {
        SHB_CUDA(cudaEventRecord(ac.hiJoinEv[i], ac.hiStream[i]));
        SHB_CUDA(cudaStreamWaitEvent(st, ac.hiJoinEv[i], 0));
    }
    lap(2, t0);
    SHB_CUDA(cudaEventRecord(ev.b, st));
}

} // namespace

void destroyAlignCache(shb_context* c)
{
    if(c->alignCache) { delete static_cast<AlignCache*>(c->alignCache); c->alignCache = nullptr; }
}


void computeAlignments(shb_context* c, const void* candidatesHost, uint64_t n, const shb_align_options& o,
                       void** alignmentDataOut, uint64_t* alignmentCountOut,
                       uint64_t** compressedTocOut, uint8_t** compressedDataOut, shb_align_result* result)
{
    SHB_REQUIRE(c->haveMarkers, SHB_ERR_STATE, "Markers are not accessible.");
    SHB_REQUIRE(c->readBegin == 0 && c->readEnd == c->readCountTotal, SHB_ERR_STATE,
                "computeAlignments needs the markers of all reads on this GPU.");
    SHB_REQUIRE(o.alignMethod == 3 || o.alignMethod == 4, SHB_ERR_INVALID,
                "Only Align.alignMethod 3 and 4 are implemented (0 and 1 are not on the hot path).");
    SHB_REQUIRE(o.gapScore <= 0, SHB_ERR_INVALID, "Align.gapScore must not be positive.");
    SHB_REQUIRE(o.k >= 1 && o.k <= 16, SHB_ERR_INVALID, "Invalid k.");
    SHB_REQUIRE(n == 0 || candidatesHost != nullptr, SHB_ERR_INVALID, "Null candidates.");
    SHB_CUDA(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    g_launchCount = 0;
    const auto wall0 = std::chrono::steady_clock::now();
    Events totalEv, dpEv1, dpEv2;
    SHB_CUDA(cudaEventRecord(totalEv.a, st));
    double dpMs = 0.;
    const bool method4 = (o.alignMethod == 4);

    // src/AssemblerAlign.cpp:378 asserts readIds[0] < readIds[1].
    const uint32_t* cand = static_cast<const uint32_t*>(candidatesHost);
    for(uint64_t i = 0; i < n; i++) {
        SHB_REQUIRE(cand[3*i] < cand[3*i+1] && cand[3*i+1] < c->readCountTotal, SHB_ERR_INVALID, "Invalid alignment candidate.");
    }

    AlignCache& ac = cache(c);
    uint32_t maxStage1Width = 0;
    if(method4) {
        SHB_REQUIRE(o.align4DeltaX >= 1 && o.align4DeltaY >= 1 && o.align4DeltaX < (1ull << 31) && o.align4DeltaY < (1ull << 31),
                    SHB_ERR_INVALID, "Invalid Align.align4.deltaX / deltaY.");
        buildSortedMarkers(c, o.k);
    } else {
        buildDownsampled(c, o.k, o.downsamplingFactor);
        maxStage1Width = 2 * ac.dsMaxRow + 2 + 64;
        SHB_REQUIRE(maxStage1Width <= kMaxBandWidth, SHB_ERR_INVALID,
                    "Downsampled reads are too long for the stage-1 kernel (limit 8191 downsampled markers).");
    }
    if(ac.lengthCheckGeneration != c->markerGeneration) {       // the traceback packs a run length above a 28-bit ordinal
        for(size_t r = 0; r + 1 < c->tocHost.size(); r++) {
            SHB_REQUIRE(c->tocHost[r + 1] - c->tocHost[r] < (1ull << kRunLengthShift), SHB_ERR_INVALID, "A read has 2^28 or more markers.");
        }
        ac.lengthCheckGeneration = c->markerGeneration;
    }
    const uint32_t maxStage2Width = uint32_t(std::max(0, o.maxBand)) + 2 + 64;     // padded to a multiple of 64
    SHB_REQUIRE(maxStage2Width <= kMaxBandWidth, SHB_ERR_INVALID, "Align.maxBand too large for this implementation (limit 16382).");

    // Method 3 uses the configured scores; Align4 hard-codes 6/-1/-1 (src/Align4.hpp:159-161: never overwritten).
    const DpScores scores = method4 ? DpScores{6, -1, -1} : DpScores{o.matchScore, o.mismatchScore, o.gapScore};
    FilterOptions fo;
    fo.minAlignedMarkerCount = uint64_t(o.minAlignedMarkerCount); fo.maxSkip = uint64_t(o.maxSkip);
    fo.maxDrift = uint64_t(o.maxDrift); fo.maxTrim = uint64_t(o.maxTrim);
    fo.minAlignedFraction = o.minAlignedFraction;
    fo.suppressContainments = (!method4 && o.suppressContainments) ? 1u : 0u;     // method 4 applies it after the selection

    // SHB_ALIGN_BATCH / SHB_ALIGN_CHUNK: test hooks that shrink the batch and chunk sizes so that small inputs exercise the
    // multi-batch, multi-chunk paths (tests/test_gpu_scale.py).
    const uint32_t batchMax = envCount("SHB_ALIGN_BATCH", method4 ? 32768 : 262144);
    const uint64_t cellBudget = 192ull << 20;      // method 4: cells of scratch per batch
    Batch& b = ac.batch;
    c->scanWs.reserve(scanWorkspaceElements(4ull * batchMax * 64));
    // scalars: 512 entries, allocated once at context creation
    unsigned long long* total64 = c->scalars.get() + 48;
    uint32_t* total32 = reinterpret_cast<uint32_t*>(c->scalars.get() + 32);

    // Kept alignments accumulate on the device and are copied to the host once at the end.
    DeviceBuffer<uint32_t>& outRecords = ac.outRecords;
    DeviceBuffer<unsigned long long>& outToc = ac.outToc;
    DeviceBuffer<uint8_t>& outData = ac.outData;
    uint64_t outCount = 0, outBytes = 0;
    unsigned long long* skippedDev = c->scalars.get() + 56;
    unsigned long long* forwardCellsDev = c->scalars.get() + 57;          // cells of the trace-free stage-1 jobs
    SHB_CUDA(cudaMemsetAsync(skippedDev, 0, 2 * sizeof(unsigned long long), st));
    uint64_t dpCells = 0;
    const std::vector<uint64_t>& toc = c->tocHost;

    PhaseClock phases;
    phases.lap(0, st);
    for(uint64_t begin = 0; begin < n; ) {
        // Batch size: bounded number of candidates and (method 4) of grid cells.
        uint32_t nb = 0;
        if(method4) {
            uint64_t cells = 0;
            while(begin + nb < n && nb < batchMax) {
                const uint64_t i = begin + nb;
                const uint64_t o0 = 2ull * cand[3*i], o1 = 2ull * cand[3*i+1] + ((cand[3*i+2] & 0xff) ? 0 : 1);
                const uint64_t nx = toc[o0+1] - toc[o0], ny = toc[o1+1] - toc[o1];
                uint64_t cc = 2;
                if(nx && ny) cc += ((nx + ny - 2) / o.align4DeltaX + 1) * ((nx + ny - 2) / o.align4DeltaY + 1);
                if(nb && cells + cc > cellBudget) break;
                cells += cc; nb++;
            }
        } else nb = uint32_t(std::min<uint64_t>(batchMax, n - begin));

        b.cand.reserve(3ull * nb);
        SHB_CUDA(cudaMemcpyAsync(b.cand.get(), cand + 3 * begin, 12ull * nb, cudaMemcpyHostToDevice, st));
        b.keep.reserve(nb); b.keepIndex.reserve(nb); b.bytes32.reserve(nb); b.bytes64.reserve(nb); b.bytesOff.reserve(nb);
        b.scanWs64.reserve(scanWorkspaceElements(nb));
        uint32_t nJobs = 0;
        const uint32_t* jobIndex = nullptr;
        DpTotals totals;
        bool stage1Timed = false;

        if(!method4) {
            // ---- method 3 ---------------------------------------------------------------------------
            nJobs = nb;
            b.jobs1.reserve(nb); b.jobs.reserve(nb); b.tw.reserve(nb); b.twOff.reserve(nb); b.outCnt.reserve(nb);
            SHB_LAUNCH(method3SetupKernel, ceilDiv(nb, 256), 256, 0, st, (const uint32_t*)b.cand.get(), nb,
                       (const uint64_t*)c->toc.get(), (const uint64_t*)ac.dsToc.get(), b.jobs1.get(), b.jobs.get(), b.tw.get(), b.outCnt.get(), forwardCellsDev);
            exclusiveScan<unsigned long long>(b.tw.get(), b.twOff.get(), nb, total64, b.scanWs64.get(), st);
            const unsigned long long traceWords1 = readBack<unsigned long long>(total64, st);
            b.trace.reserve(traceWords1 + 1);
            dpCells += 16ull * traceWords1;
            SHB_LAUNCH(setTraceOffsetsKernel, ceilDiv(nb, 256), 256, 0, st, b.jobs1.get(), nb, (const unsigned long long*)b.twOff.get(),
                       (const unsigned long long*)nullptr);
            // Ordinal slots (also the stage-1 path scratch): offsets must be in jobs[] before stage 1 runs.
            b.outOff.reserve(nb);
            exclusiveScan<unsigned long long>(b.outCnt.get(), b.outOff.get(), nb, total64, b.scanWs64.get(), st);
            const unsigned long long slots1 = readBack<unsigned long long>(total64, st);
            b.ordinals.reserve(slots1 + 1);
            SHB_LAUNCH(setOutOffsetsKernel, ceilDiv(nb, 256), 256, 0, st, b.jobs.get(), nb, (const unsigned long long*)b.outOff.get());
            Method3Args g1;
            g1.candidates = b.cand.get(); g1.candidateBegin = begin; g1.n = nb;
            g1.toc = c->toc.get(); g1.dsToc = ac.dsToc.get(); g1.dsKmer = ac.dsKmer.get(); g1.dsOrdinal = ac.dsOrdinal.get();
            g1.scores = scores; g1.bandExtend = o.bandExtend; g1.maxBand = o.maxBand;
            std::vector<uint64_t> classCounts1;
            buildClassOrder(c, b, b.jobs1.get(), nb, classCounts1, kForwardClassCount);
            SHB_CUDA(cudaEventRecord(dpEv1.a, st));
            forEachClassConcurrently(c, classCounts1, 0xffffffffu, [&](int k, uint32_t count, uint64_t offset, cudaStream_t s) {
                Method3Args gk = g1;
                gk.n = count; gk.order = b.order + offset;
                if(k < kForwardClassCount) {
                    gk.wMax = 0;
                    launchStage1Forward(kForwardRows[k], ceilDiv(count, kDpMaxWarpsPerBlock), kDpMaxWarpsPerBlock * 32, s, gk,
                                        (const DpJob*)b.jobs1.get(), b.jobs.get());
                    return;
                }
                const DpClass& cls = kClasses[k - kForwardClassCount];      // too long for the forward kernel: DP with a trace
                const uint32_t warps = warpsForClass(cls);
                const size_t smem = smemForClass(cls, warps);
                gk.wMax = cls.wMax;
                launchStage1(cls, ceilDiv(count, warps), warps * 32, smem, s, gk, b.jobs1.get(), b.trace.get(), b.jobs.get(), b.ordinals.get());
            });
            SHB_CUDA(cudaEventRecord(dpEv1.b, st));
            phases.lap(1, st);
            if(phases.on && begin == 0) {        // SHB_TRACE: band-width and active-column statistics of the first batch
                std::vector<DpJob> h(nb);
                SHB_CUDA(cudaMemcpy(h.data(), b.jobs.get(), sizeof(DpJob) * nb, cudaMemcpyDeviceToHost));
                uint64_t hist[12] = {0}, run = 0, cols = 0, fullCols = 0;
                for(const DpJob& j : h) {
                    if(j.state != kStateRun) continue;
                    run++;
                    const uint32_t w = uint32_t(j.hi - j.lo + 1);
                    hist[std::min<uint32_t>(11, (w + 7) / 8)]++;
                    cols += uint64_t(dpLastColumn(j.nx, j.ny, j.hi) - dpFirstColumn(j.lo)); fullCols += j.nx;
                }
                fprintf(stderr, "[shb] stage-2 jobs %llu of %u; band width histogram (bins of 8 offsets, last = wider):", (unsigned long long)run, nb);
                for(int k = 0; k < 12; k++) fprintf(stderr, " %llu", (unsigned long long)hist[k]);
                fprintf(stderr, "; active columns %.1f of %.1f per job\n", double(cols) / double(run ? run : 1), double(fullCols) / double(run ? run : 1));
            }
            SHB_LAUNCH(stage2TraceWordsKernel, ceilDiv(nb, 256), 256, 0, st, (const DpJob*)b.jobs.get(), nb, b.tw.get());
            runBandedJobs(c, b, nJobs, c->kmerIds, scores, dpEv2, totals);
            phases.lap(2, st);
            // Epilogue per job == per candidate.
            b.infoWords.reserve(13ull * nJobs);
            SHB_LAUNCH(alignmentInfoKernel, ceilDiv(nJobs, 4), 128, 0, st, nJobs, (const DpJob*)b.jobs.get(), (const uint2*)b.ordinals.get(),
                       (const uint32_t*)b.counts.get(), fo, b.infoWords.get(), b.keep.get(), b.bytes32.get(), skippedDev);
            stage1Timed = true;
        } else {
            // ---- method 4 ---------------------------------------------------------------------------
            b.cellCnt.reserve(nb); b.cellOff.reserve(nb); b.componentCount.reserve(nb); b.jobOffsets.reserve(nb); b.selected.reserve(nb);
            SHB_LAUNCH(align4CellCountKernel, ceilDiv(nb, 256), 256, 0, st, (const uint32_t*)b.cand.get(), nb, (const uint64_t*)c->toc.get(),
                       uint32_t(o.align4DeltaX), uint32_t(o.align4DeltaY), b.cellCnt.get());
            exclusiveScan<unsigned long long>(b.cellCnt.get(), b.cellOff.get(), nb, total64, b.scanWs64.get(), st);
            const unsigned long long cells = readBack<unsigned long long>(total64, st);
            b.gridCounts.reserve(cells + 1); b.gridAux.reserve(cells + 1); b.gridList.reserve(cells + 1);
            b.gridFlags.reserve(cells + 1); b.gridBands.reserve(cells + 1);
            Align4Args g;
            g.candidates = b.cand.get(); g.n = nb; g.toc = c->toc.get();
            g.sortedKmer = ac.sortedKmer.get(); g.sortedOrdinal = ac.sortedOrdinal.get();
            g.deltaX = uint32_t(o.align4DeltaX); g.deltaY = uint32_t(o.align4DeltaY);
            g.minEntryCountPerCell = o.align4MinEntryCountPerCell; g.maxDistanceFromBoundary = o.align4MaxDistanceFromBoundary;
            g.maxBand = int64_t(uint64_t(o.maxBand));
            g.cellOffsets = b.cellOff.get(); g.counts = b.gridCounts.get(); g.aux = b.gridAux.get(); g.list = b.gridList.get();
            g.flags = b.gridFlags.get(); g.bands = b.gridBands.get(); g.componentCount = b.componentCount.get();
            SHB_LAUNCH(align4FrontEndKernel, ceilDiv(nb, 4), 128, 0, st, g);
            exclusiveScan<uint32_t>(b.componentCount.get(), b.jobOffsets.get(), nb, total32, c->scanWs.get(), st);
            nJobs = readBack<uint32_t>(total32, st);
            if(nJobs) {
                b.jobs.reserve(nJobs); b.tw.reserve(nJobs); b.outCnt.reserve(nJobs);
                SHB_LAUNCH(align4MakeJobsKernel, ceilDiv(nb, 256), 256, 0, st, (const uint32_t*)b.cand.get(), nb, (const uint64_t*)c->toc.get(),
                           (const unsigned long long*)b.cellOff.get(), (const int32_t*)b.gridBands.get(),
                           (const uint32_t*)b.componentCount.get(), (const uint32_t*)b.jobOffsets.get(), b.jobs.get(), b.tw.get(), b.outCnt.get());
                runBandedJobs(c, b, nJobs, c->kmerIds, scores, dpEv2, totals);
                b.infoWords.reserve(13ull * nJobs); b.jobKeep.reserve(nJobs); b.jobBytes.reserve(nJobs);
                // Align4's own filters (src/Align4.cpp:944-985), identical thresholds, no containment test.
                SHB_LAUNCH(alignmentInfoKernel, ceilDiv(nJobs, 4), 128, 0, st, nJobs, (const DpJob*)b.jobs.get(), (const uint2*)b.ordinals.get(),
                           (const uint32_t*)b.counts.get(), fo, b.infoWords.get(), b.jobKeep.get(), b.jobBytes.get(),
                           (unsigned long long*)nullptr);
            } else {
                b.infoWords.reserve(13); b.jobKeep.reserve(1); b.jobBytes.reserve(1); b.jobs.reserve(1); b.counts.reserve(1); b.ordinals.reserve(1);
            }
            SHB_LAUNCH(align4SelectKernel, ceilDiv(nb, 256), 256, 0, st, nb, (const uint32_t*)b.jobOffsets.get(),
                       (const uint32_t*)b.componentCount.get(), (const uint32_t*)b.jobKeep.get(), (const uint32_t*)b.infoWords.get(),
                       (const uint32_t*)b.jobBytes.get(), o.suppressContainments ? 1u : 0u, uint32_t(o.maxTrim),
                       b.selected.get(), b.keep.get(), b.bytes32.get());
            jobIndex = b.selected.get();
        }
        dpCells += 16ull * totals.traceWords;

        // ---- compaction + output of the kept alignments ---------------------------------------------------
        phases.lap(3, st);
        exclusiveScan<uint32_t>(b.keep.get(), b.keepIndex.get(), nb, total32, c->scanWs.get(), st);
        const uint32_t kept = readBack<uint32_t>(total32, st);
        SHB_LAUNCH(widenBytesKernel, ceilDiv(nb, 256), 256, 0, st, (const uint32_t*)b.bytes32.get(), nb, b.bytes64.get());
        exclusiveScan<unsigned long long>(b.bytes64.get(), b.bytesOff.get(), nb, total64, b.scanWs64.get(), st);
        const unsigned long long bytes = readBack<unsigned long long>(total64, st);
        if(nJobs) {
            float ms = 0.f;
            SHB_CUDA(cudaEventElapsedTime(&ms, dpEv2.a, dpEv2.b));
            dpMs += ms;
        }
        if(stage1Timed) {
            float ms = 0.f;
            SHB_CUDA(cudaEventElapsedTime(&ms, dpEv1.a, dpEv1.b));
            dpMs += ms;
        }
        if(kept) {
            outRecords.reserve(16ull * (outCount + kept), true, st);
            outToc.reserve(outCount + kept + 1, true, st);
            outData.reserve(outBytes + bytes + 16, true, st);
            SHB_LAUNCH(alignmentWriteKernel, ceilDiv(nb, 4), 128, 0, st, nb, (const uint32_t*)b.cand.get(), (const DpJob*)b.jobs.get(),
                       (const uint2*)b.ordinals.get(), (const uint32_t*)b.counts.get(), (const uint32_t*)b.infoWords.get(), jobIndex,
                       (const uint32_t*)b.keep.get(), (const uint32_t*)b.keepIndex.get(), (const unsigned long long*)b.bytesOff.get(),
                       outCount, outBytes, outRecords.get(), outToc.get(), outData.get());
            outCount += kept;
            outBytes += bytes;
        }
        phases.lap(4, st);
        begin += nb;
    }

    const uint64_t count = outCount;
    SHB_CUDA(cudaStreamSynchronize(st));
    const auto copy0 = std::chrono::steady_clock::now();
    void* recOut = allocHostResult(64 * count);
    uint64_t* tocOut = (uint64_t*)allocHostResult(8 * (count + 1));
    uint8_t* dataOut = (uint8_t*)allocHostResult(outBytes);
    SHB_REQUIRE(recOut && tocOut && dataOut, SHB_ERR_OOM, "Out of host memory for the alignments.");
    if(count) {
        copyToHostPipelined(c, recOut, outRecords.get(), 64 * count);
        copyToHostPipelined(c, tocOut, outToc.get(), 8 * count);
        copyToHostPipelined(c, dataOut, outData.get(), outBytes);
        SHB_CUDA(cudaStreamSynchronize(st));
    }
    const unsigned long long skipped = readBack<unsigned long long>(skippedDev, st);
    dpCells += readBack<unsigned long long>(forwardCellsDev, st);
    phases.lap(5, st);
    {
        static const char* const names[] = {"prepare", "setup+stage1", "stage2", "epilogue", "compact+write", "copy_to_host"};
        phases.report(names, 6);
        if(phases.on) {
            fprintf(stderr, "[shb] stage-2 detail (ms): dp+traceback+filter=%.1f setup=%.1f\n", g_tracePhaseMs[2], g_tracePhaseMs[3]);
            for(double& v : g_tracePhaseMs) v = 0.;
        }
    }
    tocOut[count] = outBytes;
    if(count == 0) tocOut[0] = 0;
    SHB_CUDA(cudaEventRecord(totalEv.b, st));
    SHB_CUDA(cudaStreamSynchronize(st));
    float totalMs = 0.f;
    SHB_CUDA(cudaEventElapsedTime(&totalMs, totalEv.a, totalEv.b));
    *alignmentDataOut = recOut; *alignmentCountOut = count; *compressedTocOut = tocOut; *compressedDataOut = dataOut;
    if(result) {
        result->candidateCount = n; result->alignmentCount = count; result->skippedCount = skipped;
        result->dpCells = dpCells; result->dpMs = dpMs; result->totalMs = totalMs; result->kernelLaunches = g_launchCount;
        const auto wall1 = std::chrono::steady_clock::now();
        result->outputCopyMs = std::chrono::duration<double, std::milli>(wall1 - copy0).count();
        result->hostWallMs = std::chrono::duration<double, std::milli>(wall1 - wall0).count();
    }
}

} // namespace shb
