// blob.cpp — save / load of a finalized automaton.
//
// The reference cannot serialise a trie: every PHP request that needs the same dictionary pays
// ahocorasick_init + finalize again (16 s for 100k signatures in the reference, 1.6 s here).  The flat
// description ac_trie_finalize() produces — breadth-first order, failure links, trie edges, output lists,
// prefilter bitmaps, accepted patterns — is position independent, so it is written as is; loading replays
// only the device part of finalize (table expansion + uploads).
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "acb200.h"
#include "automaton.hpp"

namespace acb200 {

static const char BLOB_MAGIC[8] = {'A', 'C', 'B', '2', '0', '0', 'v', '2'};     // v2: level-2 bitmap indexed by filter_l2_index

struct Writer {
    FILE *f; bool ok = true;
    void raw(const void *p, size_t n) { if (ok && n && fwrite(p, 1, n, f) != n) ok = false; }
    template <typename T> void pod(const T &v) { raw(&v, sizeof(T)); }
    template <typename T> void vec(const std::vector<T> &v) { const uint64_t n = v.size(); pod(n); raw(v.data(), n * sizeof(T)); }
    void str(const char *p, size_t n) { const uint64_t m = n; pod(m); raw(p, n); }
};

struct Reader {
    FILE *f; bool ok = true;
    void raw(void *p, size_t n) { if (ok && n && fread(p, 1, n, f) != n) ok = false; }
    template <typename T> void pod(T &v) { raw(&v, sizeof(T)); }
    template <typename T> void vec(std::vector<T> &v, uint64_t limit = (1ull << 34)) {
        uint64_t n = 0; pod(n);
        if (!ok || n * sizeof(T) > limit) { ok = false; return; }
        v.resize((size_t)n); raw(v.data(), (size_t)n * sizeof(T));
    }
    void str(std::string &s) {
        uint64_t n = 0; pod(n);
        if (!ok || n > (1ull << 20)) { ok = false; return; }
        s.resize((size_t)n); raw(&s[0], (size_t)n);
    }
};

bool save_flat(const FlatAutomaton &fl, const char *path, std::string &err)
{
    if (fl.bfs_order.empty()) { err = "automaton holds no expansion data (not finalized?)"; return false; }
    FILE *f = fopen(path, "wb");
    if (!f) { err = std::string("cannot open ") + path + " for writing"; return false; }
    Writer w{f};
    w.raw(BLOB_MAGIC, 8);
    w.pod(fl.n_states); w.pod(fl.n_rows); w.pod(fl.n_classes); w.pod(fl.final_bound); w.pod(fl.root);
    w.pod(fl.max_pattern_len); w.pod(fl.n_used_bytes); w.raw(fl.cls_map, 256);
    const uint32_t range_map = fl.range_map ? 1u : 0u;
    w.pod(range_map); w.pod(fl.range_lo);
    w.vec(fl.bfs_order); w.vec(fl.level_off); w.vec(fl.fail);
    w.vec(fl.edge_src); w.vec(fl.edge_dst); w.vec(fl.edge_cls); w.vec(fl.level_edge_off);
    w.vec(fl.out_off); w.vec(fl.out_idx);
    w.pod(fl.min_pattern_len); w.pod(fl.filter_w); w.pod(fl.l1_bits); w.vec(fl.l1);
    w.pod(fl.l2_log2); w.vec(fl.l2); w.pod(fl.n_grams); w.pod(fl.l1_fill);
    const uint64_t np = fl.accepted.size();
    w.pod(np);
    for (const AC_PATTERN_t &p : fl.accepted) {
        w.str(p.ptext.astring, p.ptext.length);
        const int32_t type = (int32_t)p.id.type;
        w.pod(type);
        if (p.id.type == AC_PATTID_TYPE_STRING) w.str(p.id.u.stringy ? p.id.u.stringy : "", p.id.u.stringy ? strlen(p.id.u.stringy) : 0);
        else { const int64_t num = (int64_t)p.id.u.number; w.pod(num); }
        const uint64_t aux = (uint64_t)(uintptr_t)p.aux;     // opaque to the library: written back verbatim
        w.pod(aux);
    }
    const bool ok = w.ok && fclose(f) == 0;
    if (!ok) err = std::string("write to ") + path + " failed";
    return ok;
}

bool load_flat(FlatAutomaton &fl, std::deque<std::string> &arena, const char *path, std::string &err)
{
    FILE *f = fopen(path, "rb");
    if (!f) { err = std::string("cannot open ") + path; return false; }
    Reader r{f};
    char magic[8];
    r.raw(magic, 8);
    if (!r.ok || memcmp(magic, BLOB_MAGIC, 8) != 0) { fclose(f); err = "not an acb200 automaton blob (or another version)"; return false; }
    r.pod(fl.n_states); r.pod(fl.n_rows); r.pod(fl.n_classes); r.pod(fl.final_bound); r.pod(fl.root);
    r.pod(fl.max_pattern_len); r.pod(fl.n_used_bytes); r.raw(fl.cls_map, 256);
    uint32_t range_map = 0;
    r.pod(range_map); r.pod(fl.range_lo);
    fl.range_map = range_map != 0;
    r.vec(fl.bfs_order); r.vec(fl.level_off); r.vec(fl.fail);
    r.vec(fl.edge_src); r.vec(fl.edge_dst); r.vec(fl.edge_cls); r.vec(fl.level_edge_off);
    r.vec(fl.out_off); r.vec(fl.out_idx);
    r.pod(fl.min_pattern_len); r.pod(fl.filter_w); r.pod(fl.l1_bits); r.vec(fl.l1);
    r.pod(fl.l2_log2); r.vec(fl.l2); r.pod(fl.n_grams); r.pod(fl.l1_fill);
    uint64_t np = 0;
    r.pod(np);
    if (!r.ok || np > (1ull << 28)) { fclose(f); err = "truncated or corrupt blob"; return false; }
    fl.accepted.resize((size_t)np);
    for (uint64_t i = 0; i < np && r.ok; ++i) {
        AC_PATTERN_t &p = fl.accepted[(size_t)i];
        memset(&p, 0, sizeof(p));
        std::string bytes;
        r.str(bytes);
        arena.emplace_back(std::move(bytes));
        p.ptext.astring = arena.back().data();
        p.ptext.length = arena.back().size();
        int32_t type = 0;
        r.pod(type);
        p.id.type = (enum ac_pattid_type)type;
        if (type == AC_PATTID_TYPE_STRING) {
            std::string id;
            r.str(id);
            arena.emplace_back(std::move(id));
            p.id.u.stringy = arena.back().c_str();
        } else {
            int64_t num = 0;
            r.pod(num);
            p.id.u.number = (long)num;
        }
        uint64_t aux = 0;
        r.pod(aux);
        p.aux = (void *)(uintptr_t)aux;
    }
    fclose(f);
    // Structural checks before anything is indexed: a corrupt or truncated-then-padded file must fail to load, not
    // corrupt host or device memory or silently drop matches.
    auto ascending = [](const auto &v) {
        for (size_t i = 1; i < v.size(); ++i) if (v[i] < v[i - 1]) return false;
        return true;
    };
    bool sane = r.ok && fl.n_rows == fl.n_states + 1 && fl.bfs_order.size() == fl.n_states && fl.fail.size() == (size_t)fl.n_rows &&
                fl.edge_src.size() == fl.edge_dst.size() && fl.edge_src.size() == fl.edge_cls.size() &&
                fl.level_off.size() == fl.level_edge_off.size() && fl.level_off.size() >= 2 &&
                fl.out_off.size() == (size_t)fl.final_bound && fl.root == fl.final_bound && fl.root >= 1 && fl.root < fl.n_rows &&
                fl.n_classes >= 1 && fl.n_classes <= 256 && fl.n_used_bytes <= 256 &&
                (uint64_t)fl.n_rows * fl.n_classes < (1ull << 32);
    // breadth-first levels partition the states and the edges
    sane = sane && fl.level_off.front() == 0 && fl.level_off.back() == fl.n_states && ascending(fl.level_off) &&
           fl.level_edge_off.front() == 0 && fl.level_edge_off.back() == fl.edge_src.size() && ascending(fl.level_edge_off);
    // output lists: non-decreasing offsets that end at the list's length, every final state reports something
    sane = sane && !fl.out_off.empty() && fl.out_off.front() == 0 && fl.out_off.back() == fl.out_idx.size() && ascending(fl.out_off);
    for (size_t i = 0; sane && i + 1 < fl.out_off.size(); ++i) sane = fl.out_off[i + 1] > fl.out_off[i];
    for (size_t i = 0; sane && i < fl.out_idx.size(); ++i) sane = fl.out_idx[i] < np;
    for (size_t i = 0; sane && i < fl.bfs_order.size(); ++i) sane = fl.bfs_order[i] >= 1 && fl.bfs_order[i] < fl.n_rows;
    for (size_t i = 0; sane && i < fl.fail.size(); ++i) sane = fl.fail[i] < fl.n_rows;
    for (size_t i = 0; sane && i < fl.edge_src.size(); ++i)
        sane = fl.edge_src[i] >= 1 && fl.edge_src[i] < fl.n_rows && fl.edge_dst[i] >= 1 && fl.edge_dst[i] < fl.n_rows &&
               fl.edge_cls[i] < fl.n_classes;
    for (int b = 0; sane && b < 256; ++b) sane = fl.cls_map[b] < fl.n_classes;
    // pattern lengths are recomputed, never trusted: a wrong Lmax shrinks the halo and drops matches, a wrong
    // minimum breaks the prefilter's ">= 2W bytes" premise
    uint32_t lmax = 0, lmin = 0;
    for (size_t i = 0; sane && i < fl.accepted.size(); ++i) {
        const size_t L = fl.accepted[i].ptext.length;
        sane = L >= 1 && L <= AC_PATTRN_MAX_LENGTH;
        lmax = std::max<uint32_t>(lmax, (uint32_t)L);
        lmin = lmin ? std::min<uint32_t>(lmin, (uint32_t)L) : (uint32_t)L;
    }
    sane = sane && fl.max_pattern_len == lmax && fl.min_pattern_len == lmin;
    // prefilter tables: the word size the kernels are built for, the bitmap size they stage, a level 2 of the size its
    // hash addresses
    sane = sane && (fl.filter_w == 0 || fl.filter_w == 4 || fl.filter_w == 8);
    if (sane && fl.filter_w) {
        sane = lmin >= 2 * fl.filter_w && fl.l1_bits == FILTER_L1_BITS && fl.l1.size() == FILTER_L1_BITS / 32 &&
               (fl.l2_log2 == 0 ? fl.l2.empty() : (fl.l2_log2 >= 23 && fl.l2_log2 <= 30 && fl.l2.size() == ((size_t)1 << (fl.l2_log2 - 5))));
    } else if (sane) {
        sane = fl.l1.empty() && fl.l2.empty() && fl.l2_log2 == 0;
    }
    if (!sane) { err = "truncated or corrupt blob"; return false; }
    fl.out_pat.resize(fl.out_idx.size());
    for (size_t i = 0; i < fl.out_idx.size(); ++i) fl.out_pat[i] = fl.accepted[fl.out_idx[i]];
    return true;
}

} // namespace acb200
