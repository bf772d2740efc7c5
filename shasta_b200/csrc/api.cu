// C ABI (include/shasta_b200.h): argument checking, exception -> status translation, marker upload.
#include "context.cuh"
#include "lowhash_kernels.cuh"

#include <cstring>
#include <string>

namespace shb {

static thread_local std::string g_lastError;
void setLastError(const std::string& message) { g_lastError = message; }

void lowhash0(shb_context* c, const shb_lowhash_params& p, void** candidatesOut, uint64_t* candidateCountOut,
              uint64_t* statsOut, uint64_t* iterSummary, uint64_t maxIterSummary, shb_lowhash_result* result);

void computeAlignments(shb_context* c, const void* candidatesHost, uint64_t n, const shb_align_options& o,
                       void** alignmentDataOut, uint64_t* alignmentCountOut,
                       uint64_t** compressedTocOut, uint8_t** compressedDataOut, shb_align_result* result);
void destroyAlignCache(shb_context* c);

template<class F> shb_status guarded(F&& f)
{
    try {
        f();
        return SHB_OK;
    } catch(const Error& e) {
        setLastError(e.what());
        return e.status;
    } catch(const std::exception& e) {
        setLastError(e.what());
        return SHB_ERR_INVALID;
    }
}

static void setCommonMarkerState(shb_context* c, uint64_t readCountTotal, uint64_t readBegin, uint64_t readEnd,
                                 const uint64_t* toc, const uint8_t* readFlags, uint64_t totalMarkerCount)
{
    SHB_REQUIRE(readBegin <= readEnd && readEnd <= readCountTotal, SHB_ERR_INVALID, "Invalid read range.");
    SHB_REQUIRE(toc != nullptr && (readFlags != nullptr || readCountTotal == 0), SHB_ERR_INVALID, "Null marker arrays.");
    const uint64_t rows = 2 * (readEnd - readBegin);
    SHB_REQUIRE(toc[0] == 0, SHB_ERR_INVALID, "The marker toc must be relative (toc[0] == 0).");
    for(uint64_t i = 0; i < rows; i++) {
        SHB_REQUIRE(toc[i] <= toc[i+1], SHB_ERR_INVALID, "The marker toc is not monotonic.");
    }
    c->markerGeneration++;
    c->readCountTotal = readCountTotal;
    c->readBegin = readBegin;
    c->readEnd = readEnd;
    c->totalMarkerCount = totalMarkerCount;
    c->localMarkerCount = toc[rows];
    c->tocHost.assign(toc, toc + rows + 1);
    c->readFlagsHost.assign(readFlags, readFlags + readCountTotal);
    c->toc.reserve(rows + 1);
    c->readFlags.reserve(readCountTotal + 1);
    SHB_CUDA(cudaMemcpyAsync(c->toc.get(), toc, (rows + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, c->stream));
    if(readCountTotal) {
        SHB_CUDA(cudaMemcpyAsync(c->readFlags.get(), readFlags, readCountTotal, cudaMemcpyHostToDevice, c->stream));
    }
}

} // namespace shb

using namespace shb;

extern "C" {

const char* shb_last_error(void) { return g_lastError.c_str(); }

shb_status shb_context_create(int device, shb_context** ctx)
{
    return guarded([&] {
        SHB_REQUIRE(ctx != nullptr, SHB_ERR_INVALID, "Null context pointer.");
        int count = 0;
        SHB_CUDA(cudaGetDeviceCount(&count));
        SHB_REQUIRE(device >= 0 && device < count, SHB_ERR_CUDA, "No such CUDA device.");
        cudaDeviceProp prop;
        SHB_CUDA(cudaGetDeviceProperties(&prop, device));
        SHB_REQUIRE(prop.major == 10, SHB_ERR_CUDA,
                    std::string("shasta_b200 is built for sm_100a only; device is sm_") + std::to_string(prop.major) + std::to_string(prop.minor));
        SHB_CUDA(cudaSetDevice(device));
        shb_context* c = new shb_context();
        c->device = device;
        SHB_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        SHB_CUDA(cudaStreamCreateWithFlags(&c->copyStream[0], cudaStreamNonBlocking));
        SHB_CUDA(cudaStreamCreateWithFlags(&c->copyStream[1], cudaStreamNonBlocking));
        *ctx = c;
    });
}

void shb_context_destroy(shb_context* c)
{
    if(!c) return;
    cudaSetDevice(c->device);
    cudaDeviceSynchronize();
    destroyAlignCache(c);
    if(c->stream) cudaStreamDestroy(c->stream);
    for(int i = 0; i < 2; i++) if(c->copyStream[i]) cudaStreamDestroy(c->copyStream[i]);
    delete c;
}

void shb_free(void* p) { free(p); }

shb_status shb_set_markers(shb_context* c, uint64_t readCountTotal, uint64_t readBegin, uint64_t readEnd,
                           const uint64_t* toc, const uint8_t* markerData7, const uint8_t* readFlags,
                           uint64_t totalMarkerCount)
{
    return guarded([&] {
        SHB_REQUIRE(c != nullptr, SHB_ERR_INVALID, "Null context.");
        SHB_CUDA(cudaSetDevice(c->device));
        c->haveMarkers = false;
        setCommonMarkerState(c, readCountTotal, readBegin, readEnd, toc, readFlags, totalMarkerCount);
        const uint64_t M = c->localMarkerCount;
        SHB_REQUIRE(markerData7 != nullptr || M == 0, SHB_ERR_INVALID, "Null marker data.");
        c->kmerIdsOwned.reserve(M + 64);
        // Stream the 7-byte records through two staging buffers; only the uint32 SoA stays resident.
        const uint64_t chunkMarkers = 32ull << 20;                     // multiple of 1024
        const uint64_t chunkBytes = chunkMarkers * 7;
        DeviceBuffer<uint8_t> staging[2];
        for(uint64_t begin = 0, k = 0; begin < M; begin += chunkMarkers, k++) {
            const int b = int(k & 1);
            const uint64_t nMarkers = std::min(chunkMarkers, M - begin);
            const uint64_t nBytes = nMarkers * 7;
            if(!staging[b].get()) staging[b].reserve(std::min(chunkBytes, M * 7) + 16);
            cudaStream_t s = c->copyStream[b];
            SHB_CUDA(cudaMemcpyAsync(staging[b].get(), markerData7 + begin * 7, nBytes, cudaMemcpyHostToDevice, s));
            const uint64_t wordCount = (nBytes + 3) / 4;        // the last partial word is inside the +16 slack
            SHB_LAUNCH(extractKmerIdsKernel, ceilDiv(nMarkers, kExtractMarkersPerBlock), kExtractThreads, 0, s,
                       reinterpret_cast<const uint32_t*>(staging[b].get()), wordCount, nMarkers,
                       c->kmerIdsOwned.get() + begin);
        }
        SHB_CUDA(cudaStreamSynchronize(c->copyStream[0]));
        SHB_CUDA(cudaStreamSynchronize(c->copyStream[1]));
        SHB_CUDA(cudaStreamSynchronize(c->stream));
        c->kmerIds = c->kmerIdsOwned.get();
        c->haveMarkers = true;
    });
}

shb_status shb_set_markers_device(shb_context* c, uint64_t readCountTotal, uint64_t readBegin, uint64_t readEnd,
                                  const uint64_t* tocHost, const uint32_t* kmerIdsDevice,
                                  const uint8_t* readFlagsHost, uint64_t totalMarkerCount)
{
    return guarded([&] {
        SHB_REQUIRE(c != nullptr, SHB_ERR_INVALID, "Null context.");
        SHB_CUDA(cudaSetDevice(c->device));
        c->haveMarkers = false;
        setCommonMarkerState(c, readCountTotal, readBegin, readEnd, tocHost, readFlagsHost, totalMarkerCount);
        SHB_REQUIRE(kmerIdsDevice != nullptr || c->localMarkerCount == 0, SHB_ERR_INVALID, "Null k-mer id array.");
        SHB_CUDA(cudaStreamSynchronize(c->stream));
        c->kmerIdsOwned.release();
        c->kmerIds = kmerIdsDevice;
        c->haveMarkers = true;
    });
}

shb_status shb_lowhash0(shb_context* c, const shb_lowhash_params* params, void** candidates, uint64_t* candidateCount,
                        uint64_t* stats, uint64_t* iterSummary, uint64_t maxIterSummary, shb_lowhash_result* result)
{
    return guarded([&] {
        SHB_REQUIRE(c && params && candidates && candidateCount, SHB_ERR_INVALID, "Null argument.");
        lowhash0(c, *params, candidates, candidateCount, stats, iterSummary, maxIterSummary, result);
    });
}

shb_status shb_find_alignment_candidates_lowhash0(
    shb_context* c, uint64_t readCount, const uint64_t* toc, const uint8_t* markerData7, const uint8_t* readFlags,
    const shb_lowhash_params* params, void** candidates, uint64_t* candidateCount, uint64_t* stats,
    shb_lowhash_result* result)
{
    shb_status s = shb_set_markers(c, readCount, 0, readCount, toc, markerData7, readFlags, toc ? toc[2 * readCount] : 0);
    if(s != SHB_OK) return s;
    return shb_lowhash0(c, params, candidates, candidateCount, stats, nullptr, 0, result);
}

shb_status shb_compute_alignments(shb_context* c, const void* candidates, uint64_t candidateCount,
                                  const shb_align_options* options, void** alignmentData, uint64_t* alignmentCount,
                                  uint64_t** compressedToc, uint8_t** compressedData, shb_align_result* result)
{
    return guarded([&] {
        SHB_REQUIRE(c && options && alignmentData && alignmentCount && compressedToc && compressedData, SHB_ERR_INVALID, "Null argument.");
        computeAlignments(c, candidates, candidateCount, *options, alignmentData, alignmentCount, compressedToc, compressedData, result);
    });
}

} // extern "C"
