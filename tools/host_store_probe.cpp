// Host side of draw storage: how fast can n_threads copy a 328 MB slice into FRESH pageable memory, and what do
// madvise(MADV_POPULATE_WRITE) / MADV_HUGEPAGE change?   g++ -O2 -pthread -o build/host_store_probe tools/host_store_probe.cpp
#include <sys/mman.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>
#ifndef MADV_POPULATE_WRITE
#define MADV_POPULATE_WRITE 23
#endif
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
template <class F> static double par(int T, size_t bytes, size_t chunk, F f) {
    const double t0 = now();
    std::vector<std::thread> th;
    const size_t nch = (bytes + chunk - 1) / chunk;
    for (int t = 0; t < T; ++t) th.emplace_back([&, t] { for (size_t c = t; c < nch; c += T) { size_t o = c * chunk; f(o, std::min(chunk, bytes - o)); } });
    for (auto& x : th) x.join();
    return now() - t0;
}
int main() {
    const size_t slice = (size_t)4096 * 10000 * 8, chunk = (size_t)32 << 20;
    char* src = (char*)mmap(nullptr, slice, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    memset(src, 1, slice);
    printf("hardware threads %u\n", std::thread::hardware_concurrency());
    for (int mode = 0; mode < 5; ++mode) {
        for (int T : {4, 8, 12, 16}) {
            char* dst = (char*)mmap(nullptr, slice + (2 << 20), PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
            dst = (char*)(((uintptr_t)dst + (2 << 20) - 1) & ~(uintptr_t)((2 << 20) - 1));
            double tp = 0.0, tc = 0.0;
            const char* name = "";
            if (mode == 0) { name = "memcpy into fresh 4K pages"; }
            if (mode == 1) { name = "MADV_POPULATE_WRITE (same threads) then memcpy"; tp = par(T, slice, chunk, [&](size_t o, size_t l) { madvise(dst + o, l, MADV_POPULATE_WRITE); }); }
            if (mode == 2) { name = "MADV_HUGEPAGE then memcpy"; madvise(dst, slice, MADV_HUGEPAGE); }
            if (mode == 3) { name = "MADV_HUGEPAGE + POPULATE_WRITE then memcpy"; madvise(dst, slice, MADV_HUGEPAGE); tp = par(T, slice, chunk, [&](size_t o, size_t l) { madvise(dst + o, l, MADV_POPULATE_WRITE); }); }
            if (mode == 4) { name = "memcpy into warm pages (second pass)"; par(T, slice, chunk, [&](size_t o, size_t l) { memcpy(dst + o, src + o, l); }); }
            tc = par(T, slice, chunk, [&](size_t o, size_t l) { memcpy(dst + o, src + o, l); });
            printf("%-48s T=%2d  populate %6.1f ms  copy %6.1f ms  total %6.1f ms  (%.1f GB/s)\n", name, T, tp * 1e3, tc * 1e3, (tp + tc) * 1e3, slice / (tp + tc) * 1e-9);
            munmap((void*)((uintptr_t)dst & ~(uintptr_t)4095), slice);
        }
    }
    return 0;
}
