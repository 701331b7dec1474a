// Host-only check of jpeg2png_b200/csrc/copy_pool.h (built and run by tests/test_copy_pool.py).
#include "../jpeg2png_b200/csrc/copy_pool.h"

#include <stdint.h>
#include <stdio.h>

#include <thread>
#include <vector>

static int failures = 0;
#define CHECK(cond)                                                     \
    do {                                                                \
        if (!(cond)) { printf("FAILED %s:%d %s\n", __FILE__, __LINE__, #cond); failures++; } \
    } while (0)

static void fill(std::vector<unsigned char> &v, uint32_t seed) {
    uint32_t x = seed * 2654435761u + 1;
    for (auto &b : v) { x = x * 1664525u + 1013904223u; b = (unsigned char)(x >> 24); }
}

int main() {
    j2p::CopyPool &pool = j2p::CopyPool::instance();
    CHECK(pool.threads() >= 1);
    const size_t sizes[] = {0, 1, 4095, 4096, 4097, (2u << 20) - 1, (2u << 20), (2u << 20) + 1, (8u << 20) - 3, 5000000, (8u << 20)};
    for (size_t n : sizes) {
        std::vector<unsigned char> src(n + 64), dst(n + 64, 0xAB);
        fill(src, (uint32_t)n);
        pool.copy(dst.data() + 7, src.data() + 3, n);                  // unaligned on purpose
        CHECK(memcmp(dst.data() + 7, src.data() + 3, n) == 0);
        for (size_t i = 0; i < 7; i++) CHECK(dst[i] == 0xAB);           // nothing written outside
        for (size_t i = n + 7; i < dst.size(); i++) CHECK(dst[i] == 0xAB);
    }
    // concurrent callers (compute() is re-entrant): two threads, interleaved jobs
    std::vector<std::thread> callers;
    for (int t = 0; t < 3; t++)
        callers.emplace_back([&, t] {
            for (int rep = 0; rep < 6; rep++) {
                const size_t n = (3u << 20) + 12345 * (size_t)(t + 1) + (size_t)rep;
                std::vector<unsigned char> src(n), dst(n, 0);
                fill(src, (uint32_t)(t * 100 + rep));
                pool.copy(dst.data(), src.data(), n);
                if (memcmp(dst.data(), src.data(), n) != 0) { printf("FAILED concurrent copy t=%d rep=%d\n", t, rep); failures++; }
            }
        });
    for (auto &c : callers) c.join();
    // touch(): must not change contents it did not have to, must not write outside
    {
        const size_t n = (5u << 20) + 77;
        std::vector<unsigned char> buf(n + 16, 0x5A);
        pool.touch(buf.data() + 8, n);
        for (size_t i = 0; i < 8; i++) CHECK(buf[i] == 0x5A);
        for (size_t i = n + 8; i < buf.size(); i++) CHECK(buf[i] == 0x5A);
    }
    printf(failures ? "copy_pool: %d failures\n" : "copy_pool ok (%d)\n", failures);
    return failures ? 1 : 0;
}
