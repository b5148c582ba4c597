// csrc/staging_copy.hpp against memcpy: every destination / source alignment, lengths around the thresholds of the
// non-temporal path (256 bytes, 64-byte blocks, the 16-byte head), guard bytes on both sides untouched.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "staging_copy.hpp"

int main()
{
    const size_t cap = 1 << 16;
    std::vector<char> src(cap + 256), dst(cap + 256), ref(cap + 256);
    unsigned seed = 99;
    for (auto &c : src) { seed = seed * 1664525u + 1013904223u; c = (char)(seed >> 24); }
    const size_t lens[] = {0, 1, 15, 16, 17, 63, 64, 65, 255, 256, 257, 270, 271, 272, 319, 320, 321, 1023, 4096, 8191, 8192, 8193, 65536 - 64};
    long cases = 0;
    for (int nt = 0; nt < 2; ++nt)
        for (size_t doff = 0; doff < 64; ++doff)
            for (size_t soff = 0; soff < 17; soff += (soff < 3 ? 1 : 7))
                for (size_t n : lens) {
                    if (doff + n + 64 > dst.size() || soff + n > src.size()) continue;
                    memset(dst.data(), 0x5a, dst.size());
                    memset(ref.data(), 0x5a, ref.size());
                    acb200::copy_to_staging(dst.data() + doff, src.data() + soff, n, nt != 0);
                    acb200::staging_fence();
                    memcpy(ref.data() + doff, src.data() + soff, n);
                    if (memcmp(dst.data(), ref.data(), dst.size()) != 0) {
                        printf("mismatch: nt %d dst offset %zu src offset %zu length %zu\n", nt, doff, soff, n);
                        return 1;
                    }
                    ++cases;
                }
    printf("ok %ld cases\n", cases);
    return 0;
}
