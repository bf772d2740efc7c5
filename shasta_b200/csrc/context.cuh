// The per-GPU context behind the C ABI: device-resident markers and reusable workspaces.
#pragma once

#include "common.cuh"
#include "primitives.cuh"

#include <vector>

namespace shb {
constexpr int kMaxFusedIterations = 16;         // LowHash iterations hashed per pass over the k-mer ids (one slab each)
struct LowHashAccumulator {
    uint64_t count = 0;     // reduced (pairKey, count) items in the acc ping-pong buffers
    bool inB = false;       // which of the acc ping-pong buffers holds the data
    bool sorted = false;    // the reduced items are one sorted run with unique keys (nothing left to merge)
    uint64_t rawCount = 0;  // raw pair hits (one key per hit, any order) of the iterations since the last reduction, in pairsA
    uint64_t rawLimit = 1ull << 31;     // reduce the raw hits when one more iteration would exceed this many (SHB_LOWHASH_RAW_LIMIT)
};
// State of a staged LowHash0 run (shb_lowhash_begin ... shb_lowhash_emit).
struct LowHashState {
    bool active = false;
    shb_lowhash_params p{};
    uint64_t log2BucketCount = 0, bucketMask = 0, hashThreshold = 0, capacity = 0;
    uint32_t readBits = 1, slabGroup = 0;
    bool aggregateByRead = false;       // pair hits are counted per read in shared memory before they reach the accumulator
    LowHashAccumulator acc;
    uint64_t lowHashCount = 0, pairCount = 0, sweepLaunches = 0;
    uint64_t emittedCount = 0, candidateDigest = 0;     // of the last shb_lowhash_emit
    double sweepMs = 0.;
};
}

struct shb_context {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t copyStream[2] = {nullptr, nullptr};

    // ---- markers (a1, a4) -------------------------------------------------------------------
    bool haveMarkers = false;
    uint64_t markerGeneration = 0;      // bumped by every shb_set_markers*: invalidates derived caches
    uint64_t readCountTotal = 0;        // R of the whole assembly
    uint64_t readBegin = 0, readEnd = 0; // reads whose rows live on this GPU
    uint64_t totalMarkerCount = 0;      // over all reads (bucket-count rule)
    uint64_t localMarkerCount = 0;
    shb::DeviceBuffer<uint32_t> kmerIdsOwned;
    const uint32_t* kmerIds = nullptr;  // device, localMarkerCount entries (+ padding when owned)
    shb::DeviceBuffer<uint64_t> toc;    // device, relative, 2*(readEnd-readBegin)+1 entries
    shb::DeviceBuffer<uint8_t> readFlags; // device, readCountTotal entries
    std::vector<uint64_t> tocHost;      // host copy of the relative toc
    std::vector<uint8_t> readFlagsHost;

    // ---- shared workspaces --------------------------------------------------------------------
    shb::SortWorkspace sortWs;
    shb::DeviceBuffer<uint32_t> scanWs;
    shb::DeviceBuffer<unsigned long long> scalars;   // small device scratch for totals/counters

    // ---- LowHash buffers (see lowhash.cu) -----------------------------------------------------
    shb::DeviceBuffer<uint64_t> sweepKeys;  shb::DeviceBuffer<uint32_t> sweepVals;
    shb::DeviceBuffer<uint32_t> sweepTileFirstRead;     // per sweep tile: the oriented read that holds its first position
    uint64_t sweepTileGeneration = ~0ull;               // markerGeneration the table was built for
    shb::DeviceBuffer<uint64_t> entryKeysTmp; shb::DeviceBuffer<uint32_t> entryValsTmp;
    shb::DeviceBuffer<uint32_t> flagsBuf, indexBuf, segStartBuf, countsBuf;
    shb::DeviceBuffer<uint64_t> pairsA, pairsB;
    shb::DeviceBuffer<uint64_t> accKeysA, accKeysB;
    shb::DeviceBuffer<uint32_t> accValsA, accValsB;
    shb::DeviceBuffer<unsigned long long> stats;
    shb::DeviceBuffer<uint32_t> candidatesDev;

    shb::DeviceBuffer<uint64_t> partKeys; shb::DeviceBuffer<uint32_t> partVals;
    void* lowhashState = nullptr;
    void* pinnedStage[2] = {nullptr, nullptr};      // pinned staging for pipelined device->host result copies
    cudaEvent_t stageEvent[2] = {nullptr, nullptr};

    // ---- alignment cache (downsampled markers; see align.cu) ------------------------------------
    void* alignCache = nullptr;
    // ---- multi-GPU state (NCCL communicator, exchange buffers, gathered markers; see dist.cu) --------
    void* dist = nullptr;
};
