// Stress of csrc/helper_pool.hpp on the CPU: every index of every run() is executed exactly once, whatever the number of
// helpers, the limit set between runs, or the number of pieces; runs follow each other without a pause.
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "helper_pool.hpp"

using acb200::HelperPool;

int main()
{
    unsigned seed = 12345;
    auto rnd = [&seed](unsigned n) { seed = seed * 1664525u + 1013904223u; return (seed >> 8) % n; };
    long runs = 0, pieces = 0;
    for (int helpers : {0, 1, 3, 11}) {
        HelperPool pool(helpers);
        if (pool.wanted() != helpers) { printf("wanted() = %d, expected %d\n", pool.wanted(), helpers); return 1; }
        for (int round = 0; round < 3000; ++round) {
            if (round % 50 == 0) {
                const int limit = (int)rnd((unsigned)helpers + 2);        // also beyond the number of helpers
                pool.set_limit(limit);
                if (pool.helpers() != std::min(limit, helpers)) { printf("helpers() = %d\n", pool.helpers()); return 1; }
            }
            const int n = (int)rnd(70);                                    // 0 .. 69 pieces
            std::vector<std::atomic<int>> hit(n ? n : 1);
            for (auto &h : hit) h.store(0);
            std::atomic<long> sum{0};
            pool.run(n, [&](int i) {
                hit[i].fetch_add(1);
                long local = 0;
                for (int k = 0; k < (i % 7) * 200; ++k) local += k;        // pieces of uneven length
                sum.fetch_add(local + 1);
            });
            for (int i = 0; i < n; ++i)
                if (hit[i].load() != 1) { printf("helpers %d round %d: index %d ran %d times\n", helpers, round, i, hit[i].load()); return 1; }
            if (n == 0 && sum.load() != 0) { printf("an empty run executed something\n"); return 1; }
            ++runs; pieces += n;
        }
    }
    printf("ok %ld runs %ld pieces\n", runs, pieces);
    return 0;
}
