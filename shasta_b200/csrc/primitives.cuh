// Hand-written device-wide primitives used by the LowHash and alignment pipelines:
// exclusive scan, stream compaction helpers, and (radix_sort.cuh) the LSD radix sort.
// All are plain sm_100a CUDA (warp shuffles / match_any / shared-memory atomics); no CUB/Thrust.
#pragma once

#include "common.cuh"

namespace shb {

constexpr int kScanThreads = 256;
constexpr int kScanItemsPerThread = 16;
constexpr int kScanTile = kScanThreads * kScanItemsPerThread;     // 4096

// ---------------------------------------------------------------------------------------------
// Block-wide exclusive scan of one value per thread (256 threads). Returns the exclusive prefix;
// `total` receives the block total. `smem` must hold 8 T values. Ends with a __syncthreads so the
// scratch can be reused immediately.
template<class T> __device__ __forceinline__ T blockExclusiveScan256(T v, T& total, T* smem)
{
    const unsigned lane = threadIdx.x & 31u;
    const unsigned warp = threadIdx.x >> 5;
    T inc = v;
#pragma unroll
    for(int d = 1; d < 32; d <<= 1) {
        T t = __shfl_up_sync(0xffffffffu, inc, d);
        if(lane >= (unsigned)d) inc += t;
    }
    if(lane == 31) smem[warp] = inc;
    __syncthreads();
    T warpOffset = 0;
    T tot = 0;
#pragma unroll
    for(int w = 0; w < kScanThreads / 32; w++) {
        T s = smem[w];
        if((unsigned)w < warp) warpOffset += s;
        tot += s;
    }
    __syncthreads();
    total = tot;
    return warpOffset + inc - v;
}

template<class T> __global__ void __launch_bounds__(kScanThreads)
scanReduceKernel(const T* __restrict__ in, T* __restrict__ blockSums, uint64_t n)
{
    __shared__ T smem[kScanThreads / 32];
    const uint64_t base = uint64_t(blockIdx.x) * kScanTile;
    T sum = 0;
#pragma unroll
    for(int i = 0; i < kScanItemsPerThread; i++) {
        const uint64_t idx = base + uint64_t(i) * kScanThreads + threadIdx.x;
        if(idx < n) sum += in[idx];
    }
    T total;
    blockExclusiveScan256<T>(sum, total, smem);
    if(threadIdx.x == 0) blockSums[blockIdx.x] = total;
}

// out[i] = blockOffsets[block] + exclusive prefix within the block tile. in and out may alias.
// If totalOut != nullptr (single-block top level) it receives the grand total.
template<class T> __global__ void __launch_bounds__(kScanThreads)
scanDownsweepKernel(const T* in, T* out, const T* __restrict__ blockOffsets, uint64_t n, T* totalOut)
{
    __shared__ T smem[kScanThreads / 32];
    const uint64_t base = uint64_t(blockIdx.x) * kScanTile;
    T running = blockOffsets ? blockOffsets[blockIdx.x] : T(0);
#pragma unroll 1
    for(int i = 0; i < kScanItemsPerThread; i++) {
        const uint64_t idx = base + uint64_t(i) * kScanThreads + threadIdx.x;
        const T v = (idx < n) ? in[idx] : T(0);
        T total;
        const T ex = blockExclusiveScan256<T>(v, total, smem);
        if(idx < n) out[idx] = running + ex;
        running += total;
    }
    if(totalOut && threadIdx.x == 0) *totalOut = running;
}

// Workspace elements (of T) needed by exclusiveScan for n items.
inline uint64_t scanWorkspaceElements(uint64_t n)
{
    uint64_t total = 0;
    while(n > kScanTile) {
        n = (n + kScanTile - 1) / kScanTile;
        total += n;
    }
    return total + 1;
}

// Exclusive scan of n items. in/out may alias. totalOut (device pointer, may be null) receives the
// sum of all items. workspace must hold scanWorkspaceElements(n) items.
template<class T> void exclusiveScan(const T* in, T* out, uint64_t n, T* totalOut, T* workspace, cudaStream_t stream)
{
    if(n == 0) {
        if(totalOut) SHB_CUDA(cudaMemsetAsync(totalOut, 0, sizeof(T), stream));
        return;
    }
    if(n <= kScanTile) {
        SHB_LAUNCH((scanDownsweepKernel<T>), 1, kScanThreads, 0, stream, in, out, (const T*)nullptr, n, totalOut);
        return;
    }
    const uint64_t blocks = (n + kScanTile - 1) / kScanTile;
    T* sums = workspace;
    SHB_LAUNCH((scanReduceKernel<T>), (unsigned)blocks, kScanThreads, 0, stream, in, sums, n);
    exclusiveScan<T>(sums, sums, blocks, totalOut, workspace + blocks, stream);
    SHB_LAUNCH((scanDownsweepKernel<T>), (unsigned)blocks, kScanThreads, 0, stream, in, out, (const T*)sums, n, (T*)nullptr);
}

} // namespace shb

#include "radix_sort.cuh"      // radixSort<HAS_VALUES>, SortWorkspace

namespace shb {

// counts[digit] += 1 for every key (warp-aggregated atomics; the input is sorted by digit so runs are long).
static __global__ void digitCountKernel(const uint64_t* __restrict__ keys, uint32_t n, int shift, uint32_t digitMask,
                                        unsigned long long* __restrict__ counts)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = i < n;
    const uint32_t d = active ? (uint32_t(keys[i] >> shift) & digitMask) : 0xffffffffu;
    const unsigned peers = __match_any_sync(0xffffffffu, d);
    if(active && (__ffs(peers) - 1) == int(threadIdx.x & 31u)) atomicAdd(&counts[d], (unsigned long long)__popc(peers));
}

// ---------------------------------------------------------------------------------------------
// Run detection on a sorted key array: flags[i] = 1 where (keys[i] >> shift) differs from its
// predecessor (or i == 0).
static __global__ void headFlagsKernel(const uint64_t* __restrict__ keys, uint32_t n, int shift, uint32_t* __restrict__ flags)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) return;
    flags[i] = (i == 0 || (keys[i] >> shift) != (keys[i-1] >> shift)) ? 1u : 0u;
}

// Given flags and their exclusive scan (segIndex), write segStart[segIndex[i]] = i for heads and
// segStart[numSegments] = n (thread n-1 does the latter).
static __global__ void segmentStartsKernel(const uint32_t* __restrict__ flags, const uint32_t* __restrict__ segIndexExclusive,
                                    uint32_t n, uint32_t* __restrict__ segStart)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) return;
    if(flags[i]) segStart[segIndexExclusive[i]] = i;
    if(i == n - 1) segStart[segIndexExclusive[i] + flags[i]] = n;
}

} // namespace shb
