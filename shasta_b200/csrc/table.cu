// computeAlignmentTable on the GPU (src/AssemblerAlign.cpp:509-571): 4 (row, other, alignmentIndex) entries per
// alignment, stable radix sort by (row, other), row starts by binary search.
#include "context.cuh"

namespace shb {
namespace {

__global__ void alignmentTableKeysKernel(const uint32_t* __restrict__ records, uint32_t n, uint64_t* __restrict__ keys,
                                         uint32_t* __restrict__ vals)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) return;
    const uint32_t r0 = records[16ull * i], r1 = records[16ull * i + 1];
    const bool same = (records[16ull * i + 2] & 0xffu) != 0;
    const uint32_t o0 = 2 * r0, o1 = 2 * r1 + (same ? 0u : 1u);
    // Generated in alignment order; the sort is stable, so equal (row, other) keep increasing alignment index.
    keys[4ull * i + 0] = (uint64_t(o0) << 32) | o1;            vals[4ull * i + 0] = i;
    keys[4ull * i + 1] = (uint64_t(o1) << 32) | o0;            vals[4ull * i + 1] = i;
    keys[4ull * i + 2] = (uint64_t(o0 ^ 1u) << 32) | (o1 ^ 1u); vals[4ull * i + 2] = i;
    keys[4ull * i + 3] = (uint64_t(o1 ^ 1u) << 32) | (o0 ^ 1u); vals[4ull * i + 3] = i;
}

__global__ void alignmentTableTocKernel(const uint64_t* __restrict__ sortedKeys, uint32_t entries, uint32_t rows, uint32_t* __restrict__ toc)
{
    const uint32_t row = blockIdx.x * blockDim.x + threadIdx.x;
    if(row > rows) return;
    uint32_t lo = 0, hi = entries;          // first entry whose row is >= this row
    while(lo < hi) { const uint32_t mid = lo + ((hi - lo) >> 1); if(uint32_t(sortedKeys[mid] >> 32) < row) lo = mid + 1; else hi = mid; }
    toc[row] = lo;
}

} // namespace

void computeAlignmentTable(shb_context* c, const void* alignmentData, uint64_t n, uint64_t readCount,
                           uint32_t** tocOut, uint32_t** dataOut)
{
    SHB_REQUIRE(4 * n < (1ull << 32), SHB_ERR_INVALID, "Too many alignments for a uint32 alignment table.");
    SHB_REQUIRE(readCount < (1ull << 31), SHB_ERR_INVALID, "Too many reads.");
    SHB_CUDA(cudaSetDevice(c->device));
    cudaStream_t st = c->stream;
    const uint32_t rows = uint32_t(2 * readCount);
    uint32_t* toc = (uint32_t*)malloc(sizeof(uint32_t) * (uint64_t(rows) + 1));
    uint32_t* data = (uint32_t*)malloc(sizeof(uint32_t) * (4 * n + 1));
    SHB_REQUIRE(toc && data, SHB_ERR_OOM, "Out of host memory for the alignment table.");
    if(n == 0) {
        memset(toc, 0, sizeof(uint32_t) * (uint64_t(rows) + 1));
        *tocOut = toc; *dataOut = data;
        return;
    }
    const uint32_t* rec = static_cast<const uint32_t*>(alignmentData);
    for(uint64_t i = 0; i < n; i++) {
        SHB_REQUIRE(rec[16*i] < readCount && rec[16*i+1] < readCount, SHB_ERR_INVALID, "Alignment refers to a read that does not exist.");
    }
    DeviceBuffer<uint32_t> dRec, valsA, valsB, dToc;
    DeviceBuffer<uint64_t> keysA, keysB;
    const uint32_t entries = uint32_t(4 * n);
    dRec.reserve(16 * n); keysA.reserve(entries); keysB.reserve(entries); valsA.reserve(entries); valsB.reserve(entries);
    dToc.reserve(uint64_t(rows) + 1);
    SHB_CUDA(cudaMemcpyAsync(dRec.get(), rec, 64 * n, cudaMemcpyHostToDevice, st));
    SHB_LAUNCH(alignmentTableKeysKernel, ceilDiv(n, 256), 256, 0, st, (const uint32_t*)dRec.get(), uint32_t(n), keysA.get(), valsA.get());
    uint32_t rowBits = 1;
    while((1ull << rowBits) < uint64_t(rows)) rowBits++;
    const int ranges[2][2] = {{0, int(rowBits)}, {32, 32 + int(rowBits)}};
    const bool inB = radixSort<true>(keysA.get(), keysB.get(), valsA.get(), valsB.get(), entries, ranges, 2, c->sortWs, st);
    SHB_LAUNCH(alignmentTableTocKernel, ceilDiv(uint64_t(rows) + 1, 256), 256, 0, st, (const uint64_t*)(inB ? keysB.get() : keysA.get()),
               entries, rows, dToc.get());
    SHB_CUDA(cudaMemcpyAsync(toc, dToc.get(), sizeof(uint32_t) * (uint64_t(rows) + 1), cudaMemcpyDeviceToHost, st));
    SHB_CUDA(cudaMemcpyAsync(data, inB ? valsB.get() : valsA.get(), sizeof(uint32_t) * entries, cudaMemcpyDeviceToHost, st));
    SHB_CUDA(cudaStreamSynchronize(st));
    *tocOut = toc; *dataOut = data;
}

} // namespace shb
