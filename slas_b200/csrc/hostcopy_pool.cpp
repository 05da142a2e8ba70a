// hostcopy_pool.cpp -- the copy-thread pool behind the pageable-operand path (hostcopy.cu): fork-join copies between
// the caller's pageable memory and the library's pinned bounce ring.  Plain host C++ (g++): the copy loop uses
// non-temporal AVX2 stores where the CPU has them -- the destination is read next by a DMA engine (copy-in) or by
// nobody soon (copy-out), so a cached store would only add a read-for-ownership of every destination line: a third
// more DRAM traffic on a path that is host-memory-bandwidth bound.
#include <immintrin.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <vector>

namespace slas {

struct HostCopyJob { void *dst; const void *src; size_t bytes; };  // == the declaration in common.cuh
long tunable(const char *name);

namespace {

__attribute__((target("avx2"))) void stream_copy_avx2(char *dst, const char *src, size_t bytes) {
    // head: up to the first 32-byte boundary of the destination
    size_t head = (32 - ((uintptr_t)dst & 31)) & 31;
    if (head > bytes) head = bytes;
    memcpy(dst, src, head);
    dst += head; src += head; bytes -= head;
    size_t i = 0;
    for (; i + 128 <= bytes; i += 128) {
        const __m256i v0 = _mm256_loadu_si256((const __m256i *)(src + i));
        const __m256i v1 = _mm256_loadu_si256((const __m256i *)(src + i + 32));
        const __m256i v2 = _mm256_loadu_si256((const __m256i *)(src + i + 64));
        const __m256i v3 = _mm256_loadu_si256((const __m256i *)(src + i + 96));
        _mm256_stream_si256((__m256i *)(dst + i), v0);
        _mm256_stream_si256((__m256i *)(dst + i + 32), v1);
        _mm256_stream_si256((__m256i *)(dst + i + 64), v2);
        _mm256_stream_si256((__m256i *)(dst + i + 96), v3);
    }
    _mm_sfence();
    memcpy(dst + i, src + i, bytes - i);
}

void copy_piece(char *dst, const char *src, size_t bytes) {
    static const bool avx2 = __builtin_cpu_supports("avx2");
    if (avx2 && bytes >= 4096 && tunable("host_copy_nt") != 1) stream_copy_avx2(dst, src, bytes);
    else memcpy(dst, src, bytes);
}

// ---- copy-thread pool: fork-join over fixed-size pieces ------------------------------------------------------------
struct CopyPool {
    struct Piece { char *dst; const char *src; size_t bytes; };
    std::vector<std::thread> threads;
    std::mutex mu;              // pool state
    std::mutex call_mu;         // one parallel copy at a time (it uses every thread anyway)
    std::condition_variable cv_work, cv_done;
    std::vector<Piece> pieces;
    std::atomic<size_t> next{0};
    size_t generation = 0;
    int active = 0;
    bool quit = false;

    void worker() {
        size_t seen = 0;
        for (;;) {
            {
                std::unique_lock<std::mutex> lk(mu);
                cv_work.wait(lk, [&] { return quit || generation != seen; });
                if (quit) return;
                seen = generation;
            }
            drain();
            {
                std::lock_guard<std::mutex> lk(mu);
                if (--active == 0) cv_done.notify_all();
            }
        }
    }
    void drain() {
        for (;;) {
            const size_t i = next.fetch_add(1, std::memory_order_relaxed);
            if (i >= pieces.size()) return;
            copy_piece(pieces[i].dst, pieces[i].src, pieces[i].bytes);
        }
    }
    void start(int n) {
        for (int i = 0; i < n; ++i) threads.emplace_back([this] { worker(); });
    }
    ~CopyPool() {
        {
            std::lock_guard<std::mutex> lk(mu);
            quit = true;
            cv_work.notify_all();
        }
        for (auto &t : threads)
            if (t.joinable()) t.join();
    }
} g_pool;
std::once_flag g_pool_once;

int pool_threads() {
    long knob = tunable("host_copy_threads");
    if (knob <= 0) {
        const char *env = getenv("SLAS_CUDA_HOST_COPY_THREADS");  // run-time knob in the style of SLAS_CUDA_DEVICES
        if (env) knob = atol(env);
    }
    if (knob > 0) return (int)(knob > 64 ? 64 : knob);
    // measured on the 16-core B200 host (profiles/r2e_tune_pageable.log): 4 threads 92 ms, 8 threads 66 ms, 12 threads
    // 58.7 ms, 16 threads 60 ms per 3 GiB step -- at 12 the host DRAM is the limit (9 GiB cross it per step)
    const unsigned hw = std::thread::hardware_concurrency();
    int n = hw ? (int)(hw * 3 / 4) : 4;
    if (n < 2) n = 2;
    if (n > 12) n = 12;
    return n;
}

}  // namespace

void parallel_memcpy(const HostCopyJob *jobs, int njobs) {
    size_t total = 0;
    for (int j = 0; j < njobs; ++j) total += jobs[j].bytes;
    if (total < ((size_t)512 << 10)) {  // not worth waking anybody
        for (int j = 0; j < njobs; ++j)
            if (jobs[j].bytes) memcpy(jobs[j].dst, jobs[j].src, jobs[j].bytes);
        return;
    }
    std::call_once(g_pool_once, [] { g_pool.start(pool_threads() - 1); });  // the caller is the last copier
    std::lock_guard<std::mutex> call_lk(g_pool.call_mu);
    constexpr size_t kPiece = (size_t)256 << 10;  // a 4 MiB operand still keeps every copy thread busy
    {
        std::lock_guard<std::mutex> lk(g_pool.mu);
        g_pool.pieces.clear();
        for (int j = 0; j < njobs; ++j)
            for (size_t off = 0; off < jobs[j].bytes; off += kPiece)
                g_pool.pieces.push_back({(char *)jobs[j].dst + off, (const char *)jobs[j].src + off,
                                         jobs[j].bytes - off < kPiece ? jobs[j].bytes - off : kPiece});
        g_pool.next.store(0, std::memory_order_relaxed);
        g_pool.active = (int)g_pool.threads.size();
        ++g_pool.generation;
        g_pool.cv_work.notify_all();
    }
    g_pool.drain();
    std::unique_lock<std::mutex> lk(g_pool.mu);
    g_pool.cv_done.wait(lk, [&] { return g_pool.active == 0; });
}

}  // namespace slas
