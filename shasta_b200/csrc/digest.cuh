// Order-independent digests of the path's outputs (include/shasta_b200.h: shb_digest_records / shb_digest_compressed),
// computed on the device next to the results so that every bench line can carry them: equal digests for 1, 2, 4, 8 GPUs
// and against the CPU path = parity at full scale without shipping gigabytes around.
#pragma once
#include "common.cuh"

namespace shb {

constexpr uint64_t kFnvOffset = 0xcbf29ce484222325ull, kFnvPrime = 0x100000001b3ull;

__host__ __device__ inline uint64_t fnvWord(uint64_t h, uint32_t w) { return (h ^ uint64_t(w)) * kFnvPrime; }
__host__ __device__ inline uint64_t fnvFinish(uint64_t h) { return h ^ (h >> 32); }

__device__ __forceinline__ void digestBlockAdd(uint64_t h, unsigned long long* __restrict__ out)
{
#pragma unroll
    for(int d = 16; d > 0; d >>= 1) h += __shfl_xor_sync(0xffffffffu, h, d);
    if((threadIdx.x & 31u) == 0 && h) atomicAdd(out, (unsigned long long)h);
}

// records: `words` uint32 words per record (word 2 is masked to its low byte when maskFlagWord: the bool of an
// OrientedReadPair, whose three padding bytes are zero in our output anyway).
static __global__ void digestRecordsKernel(const uint32_t* __restrict__ records, uint64_t n, uint32_t words,
                                           unsigned long long* __restrict__ out)
{
    const uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    uint64_t h = 0;
    if(i < n) {
        h = kFnvOffset;
        for(uint32_t k = 0; k < words; k++) h = fnvWord(h, records[i * words + k]);
        h = fnvFinish(h);
    }
    digestBlockAdd(h, out);
}

// One thread per alignment: pair (from the 16-word record) then the compressed bytes.
static __global__ void digestCompressedKernel(const uint32_t* __restrict__ records, uint64_t n,
                                              const unsigned long long* __restrict__ toc, uint64_t endOffset,
                                              const uint8_t* __restrict__ data, unsigned long long* __restrict__ out)
{
    const uint64_t i = uint64_t(blockIdx.x) * blockDim.x + threadIdx.x;
    uint64_t h = 0;
    if(i < n) {
        h = kFnvOffset;
        h = fnvWord(h, records[16 * i]); h = fnvWord(h, records[16 * i + 1]); h = fnvWord(h, records[16 * i + 2] & 0xffu);
        const uint64_t b = toc[i], e = (i + 1 < n) ? toc[i + 1] : endOffset;
        for(uint64_t p = b; p < e; p++) h = fnvWord(h, data[p]);
        h = fnvFinish(h);
    }
    digestBlockAdd(h, out);
}

} // namespace shb
