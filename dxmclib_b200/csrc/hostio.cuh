// hostio.cuh — host<->device plumbing of the C ABI: a per-process pool of large device allocations and chunked,
// multi-threaded copies between pageable caller arrays and the device through pinned staging buffers.
//
// Transport::operator() (reference transport.hpp:135-202) takes plain host arrays and returns plain host arrays; at
// 512x512x400 that is 0.63 GB in and 1.26 GB out per call. A cudaMemcpy from pageable memory moves that at a few
// GB/s through the driver's single staging buffer, and cudaMalloc / cudaFree of the wave buffers (tens of GB) cost
// up to a second per call. Neither is transport work, so both are taken off the caller's critical path here.
#pragma once

#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <thread>
#include <unordered_map>
#include <vector>

namespace dxmcb200 {
namespace hostio {

    // ---- pool of large device allocations --------------------------------------------------------------------
    // Blocks of at least kPoolMinBytes are parked on release instead of being returned to the driver and handed out
    // again to the next request of a similar size on the same device (the accumulators, wave buffers and staging
    // arrays of consecutive Transport calls have the same sizes). dxmcb200_trim_pool() gives everything back.
    constexpr size_t kPoolMinBytes = size_t { 16 } << 20;

    // How many bytes may stay parked per process after their context is gone. Default 0: a library user gets all device
    // memory back when Transport::operator() returns (a GUI host shares the GPU with other code). Callers that run many
    // Transport calls back to back opt in with DXMCB200_POOL_GB=<n> in the environment or dxmcb200_set_pool_limit():
    // at 1e10 histories a call then saves about 0.5-1 s of cudaMalloc / cudaFree of 40+ GB of wave buffers.
    inline std::atomic<size_t>& poolLimit()
    {
        static std::atomic<size_t> limit { [] {
            const char* env = std::getenv("DXMCB200_POOL_GB");
            const double gb = env ? std::atof(env) : 0.0;
            return gb > 0 ? static_cast<size_t>(gb * 1073741824.0) : size_t { 0 };
        }() };
        return limit;
    }

    struct Pool {
        struct Block {
            int device;
            size_t bytes;
            void* ptr;
        };
        std::mutex mutex;
        std::vector<Block> parked;
        std::unordered_map<void*, Block> live; // blocks handed out by alloc(), with their true size
        size_t parkedBytes = 0;

        static Pool& instance()
        {
            static Pool* p = new Pool; // never destroyed: the CUDA runtime may be gone before static destructors run
            return *p;
        }

        cudaError_t alloc(int device, size_t bytes, void** out)
        {
            *out = nullptr;
            if (bytes == 0)
                return cudaSuccess;
            if (bytes >= kPoolMinBytes) {
                std::lock_guard<std::mutex> lock(mutex);
                int best = -1;
                for (int i = 0; i < static_cast<int>(parked.size()); ++i) {
                    const Block& b = parked[i];
                    if (b.device == device && b.bytes >= bytes && b.bytes <= bytes + bytes / 4
                        && (best < 0 || b.bytes < parked[best].bytes))
                        best = i;
                }
                if (best >= 0) {
                    const Block b = parked[best];
                    parked.erase(parked.begin() + best);
                    parkedBytes -= b.bytes;
                    live[b.ptr] = b;
                    *out = b.ptr;
                    return cudaSuccess;
                }
            }
            cudaError_t e = cudaMalloc(out, bytes);
            if (e != cudaSuccess) { // out of memory: return the parked blocks to the driver and try once more
                cudaGetLastError();
                trim(device);
                e = cudaMalloc(out, bytes);
            }
            if (e == cudaSuccess && bytes >= kPoolMinBytes) {
                std::lock_guard<std::mutex> lock(mutex);
                live[*out] = Block { device, bytes, *out };
            }
            return e;
        }

        // the caller guarantees that no work using the block is still in flight
        void release(void* ptr)
        {
            if (!ptr)
                return;
            {
                std::lock_guard<std::mutex> lock(mutex);
                const auto it = live.find(ptr);
                if (it != live.end()) {
                    const Block b = it->second;
                    live.erase(it);
                    if (parkedBytes + b.bytes <= poolLimit().load()) {
                        parked.push_back(b);
                        parkedBytes += b.bytes;
                        return;
                    }
                }
            }
            cudaFree(ptr);
        }

        void trim(int device) // device < 0: all devices
        {
            std::vector<Block> victims;
            {
                std::lock_guard<std::mutex> lock(mutex);
                for (size_t i = 0; i < parked.size();) {
                    if (device < 0 || parked[i].device == device) {
                        victims.push_back(parked[i]);
                        parkedBytes -= parked[i].bytes;
                        parked.erase(parked.begin() + static_cast<std::ptrdiff_t>(i));
                    } else {
                        ++i;
                    }
                }
            }
            int current = 0;
            cudaGetDevice(&current);
            for (const Block& b : victims) {
                cudaSetDevice(b.device);
                cudaFree(b.ptr);
            }
            cudaSetDevice(current);
        }
    };

    template <typename T>
    inline cudaError_t poolAlloc(int device, T** out, size_t bytes)
    {
        void* p = nullptr;
        const cudaError_t e = Pool::instance().alloc(device, bytes, &p);
        *out = static_cast<T*>(p);
        return e;
    }
    inline void poolFree(void* p) { Pool::instance().release(p); }

    // ---- chunked copies through pinned staging ----------------------------------------------------------------
    // kWorkers host threads, each with its own stream and two pinned slots: while a slot is on the wire the thread
    // fills (upload) or drains (download) the other one with memcpy, so PCIe and the host memory system both stay
    // busy. The pinned slots live for the life of the process (one set per device).
    constexpr int kWorkers = 6;
    constexpr size_t kSlotBytes = size_t { 4 } << 20;

    struct Segment { // one contiguous piece of a caller array and its device counterpart
        char* host;
        char* device;
        size_t bytes;
    };

    struct Staging {
        std::mutex mutex; // one chunked copy at a time per device
        char* slot[kWorkers][2] = {};
        cudaStream_t stream[kWorkers] = {};
        cudaEvent_t done[kWorkers][2] = {};
        bool ready = false;

        static Staging& forDevice(int device)
        {
            static Staging* all = new Staging[64];
            return all[std::clamp(device, 0, 63)];
        }

        cudaError_t prepare()
        {
            if (ready)
                return cudaSuccess;
            for (int w = 0; w < kWorkers; ++w) {
                cudaError_t e = cudaStreamCreateWithFlags(&stream[w], cudaStreamNonBlocking);
                for (int b = 0; b < 2 && e == cudaSuccess; ++b) {
                    e = cudaHostAlloc(reinterpret_cast<void**>(&slot[w][b]), kSlotBytes, cudaHostAllocDefault);
                    if (e == cudaSuccess)
                        e = cudaEventCreateWithFlags(&done[w][b], cudaEventDisableTiming);
                }
                if (e != cudaSuccess)
                    return e;
            }
            ready = true;
            return cudaSuccess;
        }
    };

    // toDevice: host -> device, else device -> host. Blocks until every byte has arrived.
    inline cudaError_t copyChunked(int device, const std::vector<Segment>& segments, bool toDevice)
    {
        Staging& st = Staging::forDevice(device);
        std::lock_guard<std::mutex> lock(st.mutex);
        cudaError_t e = cudaSetDevice(device);
        if (e == cudaSuccess)
            e = st.prepare();
        if (e != cudaSuccess)
            return e;
        struct Chunk {
            char* host;
            char* device;
            size_t bytes;
        };
        std::vector<Chunk> chunks;
        for (const Segment& s : segments)
            for (size_t off = 0; off < s.bytes; off += kSlotBytes)
                chunks.push_back({ s.host + off, s.device + off, std::min(kSlotBytes, s.bytes - off) });
        std::atomic<size_t> next { 0 };
        std::atomic<int> failed { static_cast<int>(cudaSuccess) };
        auto work = [&](int w) {
            cudaSetDevice(device);
            const Chunk* pending[2] = { nullptr, nullptr }; // download: chunk whose bytes are on their way into slot b
            int b = 0;
            for (;;) {
                const size_t k = next.fetch_add(1);
                const Chunk* c = k < chunks.size() ? &chunks[k] : nullptr;
                cudaError_t err = cudaSuccess;
                if (toDevice) {
                    if (!c)
                        break;
                    err = cudaEventSynchronize(st.done[w][b]); // the slot's previous transfer has left it
                    std::memcpy(st.slot[w][b], c->host, c->bytes);
                    if (err == cudaSuccess)
                        err = cudaMemcpyAsync(c->device, st.slot[w][b], c->bytes, cudaMemcpyHostToDevice, st.stream[w]);
                    if (err == cudaSuccess)
                        err = cudaEventRecord(st.done[w][b], st.stream[w]);
                } else {
                    if (c) {
                        err = cudaMemcpyAsync(st.slot[w][b], c->device, c->bytes, cudaMemcpyDeviceToHost, st.stream[w]);
                        if (err == cudaSuccess)
                            err = cudaEventRecord(st.done[w][b], st.stream[w]);
                        pending[b] = c;
                    }
                    const int other = b ^ 1; // drain the slot requested one round earlier while this one is on the wire
                    if (pending[other]) {
                        const cudaError_t e2 = cudaEventSynchronize(st.done[w][other]);
                        if (e2 == cudaSuccess)
                            std::memcpy(pending[other]->host, st.slot[w][other], pending[other]->bytes);
                        else if (err == cudaSuccess)
                            err = e2;
                        pending[other] = nullptr;
                    }
                    if (!c) {
                        if (pending[b]) {
                            const cudaError_t e2 = cudaEventSynchronize(st.done[w][b]);
                            if (e2 == cudaSuccess)
                                std::memcpy(pending[b]->host, st.slot[w][b], pending[b]->bytes);
                            else if (err == cudaSuccess)
                                err = e2;
                        }
                        if (err != cudaSuccess)
                            failed.store(static_cast<int>(err));
                        break;
                    }
                }
                if (err != cudaSuccess) {
                    failed.store(static_cast<int>(err));
                    break;
                }
                b ^= 1;
            }
            const cudaError_t tail = cudaStreamSynchronize(st.stream[w]);
            if (tail != cudaSuccess)
                failed.store(static_cast<int>(tail));
        };
        const int nWorkers = static_cast<int>(std::min<size_t>(kWorkers, std::max<size_t>(chunks.size(), 1)));
        std::vector<std::thread> pool;
        for (int w = 1; w < nWorkers; ++w)
            pool.emplace_back(work, w);
        work(0);
        for (auto& t : pool)
            t.join();
        return static_cast<cudaError_t>(failed.load());
    }

} // namespace hostio
} // namespace dxmcb200
