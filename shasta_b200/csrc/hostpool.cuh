// Host buffers handed to the caller (candidates, AlignmentData, compressed alignments, alignment table).
//
// Small buffers are plain malloc. Large ones (>= 8 MiB) are 2 MiB aligned with transparent huge pages requested and
// are RECYCLED: shb_free puts them on a free list instead of returning them to the OS, so that a caller that runs the
// path repeatedly (the steady state bench.py measures) neither page-faults a fresh gigabyte per call nor munmaps
// one. A block that is reused is page-locked once (cudaHostRegister) and from then on filled by direct DMA instead
// of through the pinned staging buffers. A one-shot caller never pays for page-locking.
// shb_trim_host_cache() returns the cached blocks to the OS.
#pragma once
#include <cstdint>
#include <cstdlib>
#include <mutex>
#include <unordered_map>
#include <vector>
#include <sys/mman.h>
#include <cuda_runtime.h>

namespace shb {

class HostPool {
public:
    static HostPool& instance()
    {
        static HostPool* pool = new HostPool();      // never destroyed: blocks may outlive the CUDA runtime at exit
        return *pool;
    }

    // Returns nullptr when out of memory.
    void* allocate(uint64_t bytes)
    {
        if(bytes < kLargeBytes) return malloc(bytes ? bytes : 1);
        std::lock_guard<std::mutex> lock(mutex_);
        // smallest cached block that fits without wasting more than 4x
        int best = -1;
        for(size_t k = 0; k < free_.size(); k++) {
            if(free_[k].capacity >= bytes && free_[k].capacity / 4 <= bytes && (best < 0 || free_[k].capacity < free_[size_t(best)].capacity)) best = int(k);
        }
        Block blk;
        if(best >= 0) {
            blk = free_[size_t(best)];
            free_.erase(free_.begin() + best);
            if(!blk.registered && !blk.registerFailed) {
                if(cudaHostRegister(blk.p, blk.capacity, cudaHostRegisterPortable) == cudaSuccess) blk.registered = true;
                else { cudaGetLastError(); blk.registerFailed = true; }
            }
        } else {
            const uint64_t slack = bytes + bytes / 8;                           // call-to-call size jitter still fits
            blk.capacity = (slack + kAlign - 1) & ~(kAlign - 1);
            void* p = nullptr;
            if(posix_memalign(&p, kAlign, blk.capacity) != 0) return nullptr;
            madvise(p, blk.capacity, MADV_HUGEPAGE);
            blk.p = p;
        }
        live_[blk.p] = blk;
        return blk.p;
    }

    void release(void* p)
    {
        if(!p) return;
        std::lock_guard<std::mutex> lock(mutex_);
        auto it = live_.find(p);
        if(it == live_.end()) { free(p); return; }
        free_.push_back(it->second);
        live_.erase(it);
        while(free_.size() > kMaxCached) { destroy(free_.front()); free_.erase(free_.begin()); }    // oldest first
    }

    // True when device -> host copies into p can be plain asynchronous DMA.
    bool isPageLocked(const void* p)
    {
        std::lock_guard<std::mutex> lock(mutex_);
        auto it = live_.find(const_cast<void*>(p));
        return it != live_.end() && it->second.registered;
    }

    void trim()
    {
        std::lock_guard<std::mutex> lock(mutex_);
        for(Block& b : free_) destroy(b);
        free_.clear();
    }

private:
    struct Block { void* p = nullptr; uint64_t capacity = 0; bool registered = false, registerFailed = false; };
    static constexpr uint64_t kLargeBytes = 8ull << 20, kAlign = 2ull << 20;
    static constexpr size_t kMaxCached = 8;
    static void destroy(Block& b)
    {
        if(b.registered) { if(cudaHostUnregister(b.p) != cudaSuccess) cudaGetLastError(); }
        free(b.p);
    }
    std::mutex mutex_;
    std::vector<Block> free_;
    std::unordered_map<void*, Block> live_;
};

inline void* allocHostResult(uint64_t bytes) { return HostPool::instance().allocate(bytes); }

// Owns a host result block until it is handed to the caller: an exception on the way out releases it.
struct HostResult {
    void* p = nullptr;
    explicit HostResult(void* q = nullptr) : p(q) {}
    HostResult(const HostResult&) = delete;
    HostResult& operator=(const HostResult&) = delete;
    ~HostResult() { if(p) HostPool::instance().release(p); }
    void* take() { void* q = p; p = nullptr; return q; }
    void reset(void* q) { if(p) HostPool::instance().release(p); p = q; }
};

} // namespace shb
