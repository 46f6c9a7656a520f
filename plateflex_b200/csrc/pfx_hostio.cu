// pfx_hostio.cu -- copies between the device and PAGEABLE host memory (plain numpy arrays, what the Python
// layer hands to the "_host" entry points).  The driver stages such copies through one internal pinned buffer
// with a single host thread, and a freshly allocated destination is page-faulted by that same thread: 4.8 GB/s
// measured for the outputs of a transform.  Here the staging is explicit -- two pinned chunks, the DMA of chunk
// n+1 overlapping the host-side copy of chunk n -- and the host side of each chunk is split over a small team
// of threads, so first-touch faults of the destination are taken in parallel.  Pinned (or registered) host
// memory takes the plain cudaMemcpyAsync path.
#include <algorithm>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#include "pfx_common.cuh"

namespace pfx {

namespace {

constexpr size_t CHUNK = 32u << 20;      // bytes per pinned staging chunk
constexpr size_t STAGE_MIN = 4u << 20;   // smaller copies are left to the driver

struct Staging {
    std::mutex mu;
    char *chunk[2] = {nullptr, nullptr};
    cudaEvent_t ev[2] = {nullptr, nullptr};
    bool ok = false, tried = false;
    bool init()
    {
        if (tried) return ok;
        tried = true;
        for (int b = 0; b < 2; b++) {
            if (cudaHostAlloc((void **)&chunk[b], CHUNK, cudaHostAllocDefault) != cudaSuccess ||
                cudaEventCreateWithFlags(&ev[b], cudaEventDisableTiming) != cudaSuccess) {
                cudaGetLastError();
                return ok = false;
            }
        }
        return ok = true;
    }
};
// One Staging per device ordinal: its events belong to the device that is current when they are created, and
// a copy may only record them on streams of that device.  Copies to different devices do not serialise on
// each other; the state lives until the process exits (pinned memory is freed with the context).
std::mutex g_stage_mu;
std::map<int, std::unique_ptr<Staging>> g_stages;

Staging &stage_of_current_device()
{
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) cudaGetLastError();
    std::lock_guard<std::mutex> lock(g_stage_mu);
    std::unique_ptr<Staging> &p = g_stages[dev];
    if (!p) p.reset(new Staging());
    return *p;
}

int team_size()
{
    const unsigned hc = std::thread::hardware_concurrency();
    return (int)std::max(1u, std::min(8u, hc ? hc / 2 : 1u));
}

// memcpy split over the team (the calling thread takes the first slice)
void team_memcpy(char *dst, const char *src, size_t bytes)
{
    const int nt = bytes >= (8u << 20) ? team_size() : 1;
    if (nt == 1) {
        memcpy(dst, src, bytes);
        return;
    }
    const size_t per = ((bytes + nt - 1) / nt + 4095) & ~(size_t)4095;
    std::vector<std::thread> th;
    for (int t = 1; t < nt; t++) {
        const size_t off = (size_t)t * per;
        if (off >= bytes) break;
        th.emplace_back([=] { memcpy(dst + off, src + off, std::min(per, bytes - off)); });
    }
    memcpy(dst, src, std::min(per, bytes));
    for (auto &t : th) t.join();
}

}  // namespace

bool host_is_pageable(const void *p)
{
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return true;
    }
    return at.type == cudaMemoryTypeUnregistered;
}

// Device -> pageable host, synchronous: returns when dst holds the data.  Work already enqueued on other
// streams keeps running meanwhile.
int staged_d2h(void *dst, const void *src_dev, size_t bytes, cudaStream_t st)
{
    if (bytes < STAGE_MIN || !host_is_pageable(dst)) {
        PFX_CUDA(cudaMemcpyAsync(dst, src_dev, bytes, cudaMemcpyDeviceToHost, st));
        PFX_CUDA(cudaStreamSynchronize(st));
        return PFX_OK;
    }
    Staging &g_stage = stage_of_current_device();  // the caller holds a DeviceGuard of the stream's device
    std::lock_guard<std::mutex> lock(g_stage.mu);
    if (!g_stage.init()) {
        PFX_CUDA(cudaMemcpyAsync(dst, src_dev, bytes, cudaMemcpyDeviceToHost, st));
        PFX_CUDA(cudaStreamSynchronize(st));
        return PFX_OK;
    }
    const size_t n = (bytes + CHUNK - 1) / CHUNK;
    auto issue = [&](size_t c) -> cudaError_t {
        const size_t off = c * CHUNK, len = std::min(CHUNK, bytes - off);
        cudaError_t e = cudaMemcpyAsync(g_stage.chunk[c & 1], (const char *)src_dev + off, len, cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaEventRecord(g_stage.ev[c & 1], st);
        return e;
    };
    PFX_CUDA(issue(0));
    for (size_t c = 0; c < n; c++) {
        if (c + 1 < n) PFX_CUDA(issue(c + 1));  // chunk (c+1)&1 was drained in the previous trip
        PFX_CUDA(cudaEventSynchronize(g_stage.ev[c & 1]));
        const size_t off = c * CHUNK, len = std::min(CHUNK, bytes - off);
        team_memcpy((char *)dst + off, g_stage.chunk[c & 1], len);
    }
    return PFX_OK;
}

// Pageable host -> device, synchronous with respect to the source (it may be reused on return); the data is
// on the device once the stream has passed the last chunk.
int staged_h2d(void *dst_dev, const void *src, size_t bytes, cudaStream_t st)
{
    if (bytes < STAGE_MIN || !host_is_pageable(src)) {
        PFX_CUDA(cudaMemcpyAsync(dst_dev, src, bytes, cudaMemcpyHostToDevice, st));
        return PFX_OK;
    }
    Staging &g_stage = stage_of_current_device();
    std::lock_guard<std::mutex> lock(g_stage.mu);
    if (!g_stage.init()) {
        PFX_CUDA(cudaMemcpyAsync(dst_dev, src, bytes, cudaMemcpyHostToDevice, st));
        return PFX_OK;
    }
    const size_t n = (bytes + CHUNK - 1) / CHUNK;
    for (size_t c = 0; c < n; c++) {
        const size_t off = c * CHUNK, len = std::min(CHUNK, bytes - off);
        if (c >= 2) PFX_CUDA(cudaEventSynchronize(g_stage.ev[c & 1]));  // the DMA that last read this chunk is done
        team_memcpy(g_stage.chunk[c & 1], (const char *)src + off, len);
        PFX_CUDA(cudaMemcpyAsync((char *)dst_dev + off, g_stage.chunk[c & 1], len, cudaMemcpyHostToDevice, st));
        PFX_CUDA(cudaEventRecord(g_stage.ev[c & 1], st));
    }
    // the chunks are shared: leave them drained
    PFX_CUDA(cudaEventSynchronize(g_stage.ev[0]));
    if (n > 1) PFX_CUDA(cudaEventSynchronize(g_stage.ev[1]));
    return PFX_OK;
}

}  // namespace pfx
