// Device-wide LSD radix sort (64-bit keys, optional 32-bit payload), hand written for sm_100a: no CUB / Thrust.
//
// One pass = ONE kernel that reads every item once and writes it once ("onesweep" organisation):
//   * an up-front kernel reads the keys once and builds the global digit histograms of ALL passes; their exclusive scans
//     give, per pass, where each digit's run starts in the output;
//   * the pass kernel cuts the input into tiles of 4096 items. A block takes the next tile (atomic ticket, so tiles are
//     started in order), ranks its items stably inside the tile (per-warp digit counters in shared memory, match_any
//     inside a warp round), publishes the tile's digit counts, and obtains the number of equal-digit items in all EARLIER
//     tiles by decoupled look-back over the tiles' published counts / inclusive prefixes (thread d follows digit d);
//   * the items are permuted into digit order in shared memory (before the look-back, which needs no item registers)
//     and then written out run by run, so that consecutive threads write consecutive addresses.
// Per pass and item: 12 (8) bytes read + 12 (8) bytes written with (without) payload, against three kernels and two reads
// per digit of the previous histogram / scan / scatter organisation.
#pragma once

#include "common.cuh"

namespace shb {

constexpr int kRadixBits = 8;
constexpr int kRadix = 1 << kRadixBits;
constexpr int kSortThreads = 256;                     // == kRadix: thread d owns digit d in the per-digit steps
constexpr int kSortItemsPerThread = 16;
constexpr int kSortTile = kSortThreads * kSortItemsPerThread;     // 4096 items per tile
constexpr int kSortMaxPasses = 8;
static_assert(kSortThreads == kRadix, "one thread per digit");

struct SortPasses {
    int count;
    int shift[kSortMaxPasses];
    uint32_t mask[kSortMaxPasses];
};

// hist[p * 256 + d] += number of keys whose digit of pass p is d.
static __global__ void __launch_bounds__(kSortThreads)
radixGlobalHistogramKernel(const uint64_t* __restrict__ keys, uint32_t n, SortPasses passes, unsigned long long* __restrict__ hist)
{
    __shared__ uint32_t counts[kSortMaxPasses][kRadix];
    for(int p = 0; p < passes.count; p++) counts[p][threadIdx.x] = 0;
    __syncthreads();
    const uint32_t stride = gridDim.x * blockDim.x;
    for(uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
        const uint64_t key = keys[i];
        for(int p = 0; p < passes.count; p++) atomicAdd(&counts[p][uint32_t(key >> passes.shift[p]) & passes.mask[p]], 1u);
    }
    __syncthreads();
    for(int p = 0; p < passes.count; p++) {
        const uint32_t c = counts[p][threadIdx.x];
        if(c) atomicAdd(&hist[p * kRadix + threadIdx.x], (unsigned long long)c);
    }
}

// In place: hist[p][d] -> number of keys with a smaller digit in pass p (one block, one warp-scan per pass).
static __global__ void __launch_bounds__(kSortThreads)
radixDigitStartsKernel(unsigned long long* __restrict__ hist, int passCount)
{
    __shared__ unsigned long long warpTotals[kSortThreads / 32];
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    for(int p = 0; p < passCount; p++) {
        const unsigned long long v = hist[p * kRadix + threadIdx.x];
        unsigned long long inc = v;
#pragma unroll
        for(int d = 1; d < 32; d <<= 1) {
            const unsigned long long t = __shfl_up_sync(0xffffffffu, inc, d);
            if(lane >= unsigned(d)) inc += t;
        }
        if(lane == 31) warpTotals[warp] = inc;
        __syncthreads();
        unsigned long long offset = 0;
        for(unsigned w = 0; w < warp; w++) offset += warpTotals[w];
        hist[p * kRadix + threadIdx.x] = offset + inc - v;
        __syncthreads();
    }
}

// Tile status words of the look-back: bits 63..56 = tag, bits 55..0 = value. For a pass with tag base T:
// T = "digit count of the tile", T + 1 = "inclusive prefix (this tile and all earlier ones)"; anything else = not yet
// written in this pass. The tag base changes with every pass, so the array is never cleared between passes.
constexpr int kStatusTagShift = 56;
constexpr unsigned long long kStatusValueMask = (1ull << kStatusTagShift) - 1ull;

// The pass kernel runs 512 threads x 8 items on the same 4096-item tile: half the registers per thread of a 256 x 16
// organisation, so two blocks (32 warps) stay resident per SM and hide the shared-memory / match latencies of the ranking.
__device__ __forceinline__ unsigned long long loadStatus(const unsigned long long* p)
{
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void storeStatus(unsigned long long* p, unsigned long long v)
{
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" :: "l"(p), "l"(v) : "memory");
}

constexpr int kOnesweepThreads = 512;
constexpr int kOnesweepItems = kSortTile / kOnesweepThreads;

template<bool HAS_VALUES> __global__ void __launch_bounds__(kOnesweepThreads, 2)
radixOnesweepKernel(const uint64_t* __restrict__ keysIn, uint64_t* __restrict__ keysOut,
                    const uint32_t* __restrict__ valsIn, uint32_t* __restrict__ valsOut,
                    uint32_t n, int shift, uint32_t digitMask,
                    const unsigned long long* __restrict__ digitStart,      // [256] of this pass
                    unsigned long long* __restrict__ status,                 // [numTiles * 256]
                    uint32_t* __restrict__ ticket, uint32_t tagBase)
{
    constexpr int kWarps = kOnesweepThreads / 32;
    constexpr int kDigitWarps = kRadix / 32;                // warps whose threads each own one digit
    __shared__ uint32_t warpHist[kWarps][kRadix];           // per warp: digit count, then start inside the tile's digit run
    __shared__ uint32_t localOffset[kRadix];                // start of each digit's run inside the sorted tile
    __shared__ unsigned long long digitBase[kRadix];        // where the tile's digit run starts in the output
    extern __shared__ uint64_t sortedKeys[];                // kSortTile keys, then (HAS_VALUES) kSortTile payloads
    uint32_t* sortedVals = reinterpret_cast<uint32_t*>(sortedKeys + kSortTile);
    __shared__ uint32_t tileShared;
    __shared__ uint32_t scanTotals[kDigitWarps];

    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    const bool ownsDigit = threadIdx.x < unsigned(kRadix);
    if(threadIdx.x == 0) tileShared = atomicAdd(ticket, 1u);
#pragma unroll
    for(int i = 0; i < kWarps * kRadix / kOnesweepThreads; i++) (&warpHist[0][0])[i * kOnesweepThreads + threadIdx.x] = 0;
    __syncthreads();
    const uint32_t tile = tileShared;
    const uint32_t tileBegin = tile * uint32_t(kSortTile);
    const uint32_t tileCount = min(uint32_t(kSortTile), n - tileBegin);

    // Load: warp w owns the contiguous items [w*256, (w+1)*256) of the tile, 8 rounds of 32 lanes (coalesced).
    // Rank: items with the same digit keep their input order (warp, round, lane).
    uint64_t key[kOnesweepItems];
    uint32_t val[kOnesweepItems];
    uint32_t slot[kOnesweepItems];                          // digit | rank among the warp's items with that digit << 9
#pragma unroll
    for(int r = 0; r < kOnesweepItems; r++) {
        const uint32_t local = warp * (32u * kOnesweepItems) + uint32_t(r) * 32u + lane;
        const bool valid = local < tileCount;
        key[r] = valid ? keysIn[tileBegin + local] : ~0ull;
        if(HAS_VALUES) val[r] = valid ? valsIn[tileBegin + local] : 0u;
    }
    // All the match operations first (independent, so they pipeline), then per round one shared-memory atomic by each
    // digit group's first lane (returns the group's base rank and reserves its items) and one shuffle to hand the base to
    // the group: the rounds form no dependent chain through shared memory, and a warp's atomics on one counter are
    // performed in program order, which keeps the ranking stable.
    uint32_t digit[kOnesweepItems];
    unsigned peers[kOnesweepItems];
#pragma unroll
    for(int r = 0; r < kOnesweepItems; r++) {
        const uint32_t local = warp * (32u * kOnesweepItems) + uint32_t(r) * 32u + lane;
        digit[r] = (local < tileCount) ? (uint32_t(key[r] >> shift) & digitMask) : uint32_t(kRadix);      // invalid slots: digit 256
        peers[r] = __match_any_sync(0xffffffffu, digit[r]);
    }
#pragma unroll
    for(int r = 0; r < kOnesweepItems; r++) {
        const uint32_t rankInRound = __popc(peers[r] & ((1u << lane) - 1u));
        const int leader = __ffs(int(peers[r])) - 1;
        uint32_t base = 0;
        if(rankInRound == 0 && digit[r] < uint32_t(kRadix)) base = atomicAdd(&warpHist[warp][digit[r]], uint32_t(__popc(peers[r])));
        base = __shfl_sync(0xffffffffu, base, leader);
        slot[r] = digit[r] | ((base + rankInRound) << 9);
    }
    __syncthreads();

    // Per digit (thread d): the tile's count, the start of every warp's items inside the digit's run.
    const unsigned long long tagCount = (unsigned long long)(tagBase) << kStatusTagShift;
    const unsigned long long tagPrefix = (unsigned long long)(tagBase + 1u) << kStatusTagShift;
    unsigned long long* myStatus = status + uint64_t(tile) * kRadix + threadIdx.x;
    uint32_t count = 0, inc = 0;
    if(ownsDigit) {
#pragma unroll
        for(int w = 0; w < kWarps; w++) {
            const uint32_t c = warpHist[w][threadIdx.x];
            warpHist[w][threadIdx.x] = count;
            count += c;
        }
        storeStatus(myStatus, tagCount | count);
        // exclusive scan of the counts over the digits -> localOffset
        inc = count;
#pragma unroll
        for(int d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
            if(lane >= unsigned(d)) inc += t;
        }
        if(lane == 31) scanTotals[warp] = inc;
    }
    __syncthreads();
    if(ownsDigit) {
        uint32_t offset = 0;
        for(unsigned w = 0; w < warp; w++) offset += scanTotals[w];
        localOffset[threadIdx.x] = offset + inc - count;
    }
    __syncthreads();

    // Permute into digit order in shared memory (this needs only tile-local offsets, so it is done BEFORE the look-back:
    // the keys leave the registers, and the other warps' shared-memory traffic overlaps the look-back's global loads).
#pragma unroll
    for(int r = 0; r < kOnesweepItems; r++) {
        const uint32_t d = slot[r] & 0x1ffu;
        if(d < uint32_t(kRadix)) {
            const uint32_t pos = localOffset[d] + warpHist[warp][d] + (slot[r] >> 9);
            sortedKeys[pos] = key[r];
            if(HAS_VALUES) sortedVals[pos] = val[r];
        }
    }

    // Look-back. In the steady state the previous tile has already published its inclusive prefix and ONE load ends the
    // walk; only when it has not (the first wave of tiles, which all reach this point together) does the walk continue,
    // and then over a window of earlier tiles per round trip: tile j of such a wave has to pass about j/2 tiles that have
    // published only their counts. A status word carries its own tag and value, so relaxed gpu-scope accesses are enough.
    if(ownsDigit) {
        unsigned long long earlier = 0;
        int64_t t = int64_t(tile) - 1;
        if(t >= 0) {
            const unsigned long long sw = loadStatus(status + uint64_t(t) * kRadix + threadIdx.x);
            const unsigned long long tag = sw & ~kStatusValueMask;
            if(tag == tagPrefix) { earlier = sw & kStatusValueMask; t = -1; }
            else if(tag == tagCount) { earlier = sw & kStatusValueMask; t--; }
        }
        constexpr int kLookback = 8;
        while(t >= 0) {
            unsigned long long window[kLookback];
#pragma unroll
            for(int k = 0; k < kLookback; k++) window[k] = (t - k >= 0) ? loadStatus(status + uint64_t(t - k) * kRadix + threadIdx.x) : tagPrefix;
            bool done = false;
#pragma unroll
            for(int k = 0; k < kLookback; k++) {
                const unsigned long long sw = window[k];
                const unsigned long long tag = sw & ~kStatusValueMask;
                const bool isPrefix = tag == tagPrefix, isCount = tag == tagCount;
                if(!done && (isPrefix || isCount)) { earlier += sw & kStatusValueMask; t = isPrefix ? -1 : t - 1; }
                done = done || !isCount;        // a prefix ends the walk; an unpublished tile is read again (it is running: tiles start in ticket order)
            }
        }
        storeStatus(myStatus, tagPrefix | (earlier + count));
        digitBase[threadIdx.x] = digitStart[threadIdx.x] + earlier;
    }
    __syncthreads();

    // Write the runs out: item i of the sorted tile goes to digitBase[d] + (i - localOffset[d]).
#pragma unroll
    for(int r = 0; r < kOnesweepItems; r++) {
        const uint32_t i = uint32_t(r) * kOnesweepThreads + threadIdx.x;
        if(i < tileCount) {
            const uint64_t k = sortedKeys[i];
            const uint32_t d = uint32_t(k >> shift) & digitMask;
            const unsigned long long dst = digitBase[d] + (i - localOffset[d]);
            keysOut[dst] = k;
            if(HAS_VALUES) valsOut[dst] = sortedVals[i];
        }
    }
}

struct SortWorkspace {
    DeviceBuffer<unsigned long long> hist;          // [kSortMaxPasses * 256] digit starts + [kSortMaxPasses] tickets (as uint32 pairs)
    DeviceBuffer<unsigned long long> status;        // look-back status words
    uint32_t nextTag = 2;                           // tags 2..253, two per pass; the status array is cleared when they wrap
};

// Sorts n (key[,value]) items on the bit ranges given (each range [begin,end) is processed in 8-bit passes, least
// significant range first). Stable. Buffers ping-pong between (keysA,valsA) and (keysB,valsB); returns true if the
// result ends up in the B buffers.
template<bool HAS_VALUES>
bool radixSort(uint64_t* keysA, uint64_t* keysB, uint32_t* valsA, uint32_t* valsB, uint64_t n,
               const int (*bitRanges)[2], int rangeCount, SortWorkspace& ws, cudaStream_t stream)
{
    SHB_REQUIRE(n < (1ull << 32), SHB_ERR_INVALID, "radixSort: more than 2^32-1 items in one sort.");
    if(n == 0) return false;
    SortPasses passes;
    passes.count = 0;
    for(int r = 0; r < rangeCount; r++) {
        for(int bit = bitRanges[r][0]; bit < bitRanges[r][1]; bit += kRadixBits) {
            SHB_REQUIRE(passes.count < kSortMaxPasses, SHB_ERR_INVALID, "radixSort: more than 64 key bits requested.");
            const int bits = (bitRanges[r][1] - bit < kRadixBits) ? (bitRanges[r][1] - bit) : kRadixBits;
            passes.shift[passes.count] = bit;
            passes.mask[passes.count] = (1u << bits) - 1u;
            passes.count++;
        }
    }
    if(passes.count == 0) return false;
    const uint32_t numTiles = ceilDiv(n, kSortTile);
    constexpr uint64_t kHistWords = uint64_t(kSortMaxPasses) * kRadix + kSortMaxPasses;
    ws.hist.reserve(kHistWords);
    const uint64_t statusWords = uint64_t(numTiles) * kRadix;
    if(ws.status.capacity() < statusWords) {
        ws.status.reserve(statusWords);
        SHB_CUDA(cudaMemsetAsync(ws.status.get(), 0, ws.status.capacity() * sizeof(unsigned long long), stream));
        ws.nextTag = 2;
    }
    if(ws.nextTag + 2u * uint32_t(passes.count) > 254u) {       // tags about to wrap: forget everything older
        SHB_CUDA(cudaMemsetAsync(ws.status.get(), 0, ws.status.capacity() * sizeof(unsigned long long), stream));
        ws.nextTag = 2;
    }
    SHB_CUDA(cudaMemsetAsync(ws.hist.get(), 0, kHistWords * sizeof(unsigned long long), stream));
    const uint32_t histBlocks = std::min<uint32_t>(ceilDiv(n, kSortThreads * 8), 148u * 8u);
    SHB_LAUNCH(radixGlobalHistogramKernel, histBlocks, kSortThreads, 0, stream, (const uint64_t*)keysA, uint32_t(n), passes, ws.hist.get());
    SHB_LAUNCH(radixDigitStartsKernel, 1, kSortThreads, 0, stream, ws.hist.get(), passes.count);
    uint32_t* tickets = reinterpret_cast<uint32_t*>(ws.hist.get() + uint64_t(kSortMaxPasses) * kRadix);
    constexpr int kDynamicBytes = kSortTile * (HAS_VALUES ? 12 : 8);            // + ~11 KB static: above the 48 KB default
    static bool attributeSet = false;           // per template instantiation (and per process: one device per process)
    if(!attributeSet) {
        SHB_CUDA(cudaFuncSetAttribute(radixOnesweepKernel<HAS_VALUES>, cudaFuncAttributeMaxDynamicSharedMemorySize, kDynamicBytes));
        SHB_CUDA(cudaFuncSetAttribute(radixOnesweepKernel<HAS_VALUES>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        attributeSet = true;
    }
    bool inB = false;
    for(int p = 0; p < passes.count; p++) {
        uint64_t* kin = inB ? keysB : keysA;
        uint64_t* kout = inB ? keysA : keysB;
        uint32_t* vin = inB ? valsB : valsA;
        uint32_t* vout = inB ? valsA : valsB;
        SHB_LAUNCH((radixOnesweepKernel<HAS_VALUES>), numTiles, kOnesweepThreads, kDynamicBytes, stream,
                   (const uint64_t*)kin, kout, (const uint32_t*)vin, vout, uint32_t(n), passes.shift[p], passes.mask[p],
                   (const unsigned long long*)(ws.hist.get() + uint64_t(p) * kRadix),
                   ws.status.get(), tickets + 2 * p, ws.nextTag);
        ws.nextTag += 2;
        inB = !inB;
    }
    return inB;
}

} // namespace shb
