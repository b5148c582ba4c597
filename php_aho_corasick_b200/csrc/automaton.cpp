// automaton.cpp — see automaton.hpp.
#include "automaton.hpp"
#include "filter_hash.hpp"

#include <algorithm>
#include <cstring>
#include <cstdlib>

namespace acb200 {

// ----------------------------------------------------------------- EdgeMap --

static inline uint64_t mix64(uint64_t x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL;
    x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL;
    x ^= x >> 33;
    return x;
}

EdgeMap::EdgeMap() : keys_(1024, 0), vals_(1024, 0), used_(0), mask_(1023) {}

uint32_t EdgeMap::find(uint32_t node, uint8_t byte) const {
    const uint64_t key = (((uint64_t)node << 8) | byte) + 1;
    size_t i = mix64(key) & mask_;
    while (true) {
        const uint64_t k = keys_[i];
        if (k == key) return vals_[i];
        if (k == 0) return NONE;
        i = (i + 1) & mask_;
    }
}

void EdgeMap::insert(uint32_t node, uint8_t byte, uint32_t child) {
    if ((used_ + 1) * 10 > keys_.size() * 6) grow();
    const uint64_t key = (((uint64_t)node << 8) | byte) + 1;
    size_t i = mix64(key) & mask_;
    while (keys_[i] != 0) i = (i + 1) & mask_;
    keys_[i] = key;
    vals_[i] = child;
    ++used_;
}

void EdgeMap::grow() {
    std::vector<uint64_t> ok; ok.swap(keys_);
    std::vector<uint32_t> ov; ov.swap(vals_);
    keys_.assign(ok.size() * 2, 0);
    vals_.assign(ok.size() * 2, 0);
    mask_ = keys_.size() - 1;
    for (size_t j = 0; j < ok.size(); ++j) {
        if (!ok[j]) continue;
        size_t i = mix64(ok[j]) & mask_;
        while (keys_[i] != 0) i = (i + 1) & mask_;
        keys_[i] = ok[j];
        vals_[i] = ov[j];
    }
}

// ---------------------------------------------------------------- HostTrie --

HostTrie::HostTrie() {
    parent_.push_back(0);
    in_byte_.push_back(0);
    depth_.push_back(0);
    own_.push_back(-1);
}

const char *HostTrie::keep_bytes(const char *p, size_t n) {
    arena_.emplace_back(p ? p : "", p ? n : 0);
    return arena_.back().data();
}

// Acceptance rules of the reference's ac_trie_add (src/multifast/ahocorasick.c:91-131):
// empty -> ZERO_PATTERN, longer than AC_PATTRN_MAX_LENGTH -> LONG_PATTERN, the
// path is created first and only then is an already-accepting node reported as
// DUPLICATE_PATTERN (first pattern wins).  The closed-trie check lives in the
// C-ABI wrapper, which owns the finalized flag.
AC_STATUS_t HostTrie::add(const AC_PATTERN_t *patt, int copy) {
    const size_t len = patt->ptext.length;
    if (len == 0) return ACERR_ZERO_PATTERN;
    if (len > AC_PATTRN_MAX_LENGTH) return ACERR_LONG_PATTERN;

    const uint8_t *bytes = (const uint8_t *)patt->ptext.astring;
    uint32_t n = 0;
    for (size_t i = 0; i < len; ++i) {
        uint32_t next = edges_.find(n, bytes[i]);
        if (next == EdgeMap::NONE) {
            next = (uint32_t)parent_.size();
            parent_.push_back(n);
            in_byte_.push_back(bytes[i]);
            depth_.push_back((uint16_t)(depth_[n] + 1));
            own_.push_back(-1);
            edges_.insert(n, bytes[i], next);
        }
        n = next;
    }
    if (own_[n] >= 0) return ACERR_DUPLICATE_PATTERN;

    AC_PATTERN_t rec = *patt;
    if (copy) {
        // Binary-safe deep copy (the reference's pooled copy is strncpy-based,
        // src/multifast/mpool.c:174, but is never read on the match path).
        rec.ptext.astring = keep_bytes(patt->ptext.astring, len);
        rec.rtext.astring = patt->rtext.length ? keep_bytes(patt->rtext.astring, patt->rtext.length) : NULL;
        if (patt->id.type == AC_PATTID_TYPE_STRING && patt->id.u.stringy)
            rec.id.u.stringy = keep_bytes(patt->id.u.stringy, strlen(patt->id.u.stringy));
    }
    own_[n] = (int32_t)patterns_.size();
    patterns_.push_back(rec);
    return ACERR_SUCCESS;
}

void HostTrie::release_build_memory() {
    std::vector<uint32_t>().swap(parent_);
    std::vector<uint8_t>().swap(in_byte_);
    std::vector<uint16_t>().swap(depth_);
    std::vector<int32_t>().swap(own_);
    edges_ = EdgeMap();
}

// Breadth-first construction of the automaton the reference builds with its
// depth-first passes (src/multifast/ahocorasick.c:143-155):
//  * fail(v)  = deepest trie node that is a proper suffix of path(v), else root
//               (definition at ahocorasick.c:344-368);
//  * final(v) = v or any node on its failure chain accepts a pattern
//               (node.c:432-436);
//  * matched(v) = own pattern, then the patterns along the failure chain in
//               chain order = strictly decreasing length (node.c:424-437; the
//               by-text de-duplication of node.c:150-175 never fires because two
//               suffixes of one string with equal length are equal).
void HostTrie::flatten(FlatAutomaton &flat) {
    const uint32_t N = (uint32_t)parent_.size();
    const uint32_t NONE = EdgeMap::NONE;

    // children lists (counting sort by parent), each sorted by byte
    std::vector<uint32_t> child_off(N + 1, 0);
    for (uint32_t v = 1; v < N; ++v) child_off[parent_[v] + 1]++;
    for (uint32_t v = 0; v < N; ++v) child_off[v + 1] += child_off[v];
    std::vector<uint32_t> children(N ? N - 1 : 0);
    {
        std::vector<uint32_t> cur(child_off.begin(), child_off.end() - 1);
        for (uint32_t v = 1; v < N; ++v) children[cur[parent_[v]]++] = v;
    }
    for (uint32_t v = 0; v < N; ++v) {
        auto b = children.begin() + child_off[v], e = children.begin() + child_off[v + 1];
        if (e - b > 1)
            std::sort(b, e, [&](uint32_t x, uint32_t y) { return in_byte_[x] < in_byte_[y]; });
    }

    // BFS order (old ids), level offsets
    std::vector<uint32_t> order; order.reserve(N);
    std::vector<uint32_t> level_off;
    order.push_back(0);
    level_off.push_back(0);
    size_t head = 0;
    while (head < order.size()) {
        const size_t level_end = order.size();
        for (; head < level_end; ++head) {
            const uint32_t v = order[head];
            for (uint32_t k = child_off[v]; k < child_off[v + 1]; ++k) order.push_back(children[k]);
        }
        level_off.push_back((uint32_t)level_end);   // end of this level = start of the next
    }

    // failure + dictionary-suffix links in BFS order
    std::vector<uint32_t> fail(N, 0), dlink(N, NONE);
    std::vector<uint8_t> is_final(N, 0);
    for (uint32_t idx = 1; idx < N; ++idx) {
        const uint32_t v = order[idx];
        const uint32_t p = parent_[v];
        const uint8_t c = in_byte_[v];
        uint32_t f = 0;
        if (p != 0) {
            uint32_t g = fail[p];
            while (true) {
                const uint32_t t = edges_.find(g, c);
                if (t != NONE) { f = t; break; }
                if (g == 0) { f = 0; break; }
                g = fail[g];
            }
        }
        fail[v] = f;
        dlink[v] = (own_[f] >= 0) ? f : dlink[f];
        is_final[v] = (own_[v] >= 0 || dlink[v] != NONE) ? 1 : 0;
    }

    // State numbering (ids are what the device reports in events):
    //   0                      reserved "left the hot set" marker, never a real state
    //   [1, final_bound)       final states, DEEPEST first, so the shallow finals sit next to
    //   [final_bound, N+1)     the non-final states, breadth-first (root = final_bound).
    // The shared-memory window of the scan kernel is one contiguous id range around
    // final_bound: the shallowest finals below it and the shallowest non-finals above it.
    uint32_t n_final = 0;
    for (uint32_t v = 0; v < N; ++v) n_final += is_final[v];
    const uint32_t final_bound = 1 + n_final;
    std::vector<uint32_t> newid(N);
    {
        uint32_t a = final_bound, b = final_bound - 1;
        for (uint32_t idx = 0; idx < N; ++idx) {
            const uint32_t v = order[idx];
            newid[v] = is_final[v] ? b-- : a++;
        }
    }

    flat.n_states = N;
    flat.n_rows = N + 1;
    flat.final_bound = final_bound;
    flat.root = final_bound;

    // byte classes
    bool used[256] = {false};
    for (uint32_t v = 1; v < N; ++v) used[in_byte_[v]] = true;
    uint32_t n_used = 0; int lo = -1, hi = -1;
    for (int b = 0; b < 256; ++b) if (used[b]) { if (lo < 0) lo = b; hi = b; ++n_used; }
    flat.n_used_bytes = n_used;
    flat.n_classes = n_used + (n_used < 256 ? 1 : 0);
    flat.range_map = (n_used == 0) || ((uint32_t)(hi - lo + 1) == n_used);
    flat.range_lo = n_used ? (uint32_t)lo : 0;
    {
        uint32_t r = 0;
        for (int b = 0; b < 256; ++b) flat.cls_map[b] = used[b] ? (uint8_t)(r++) : (uint8_t)n_used;
        // n_used == 256 never takes the else branch; n_used < 256 keeps the value <= 255
    }

    // device expansion inputs
    flat.bfs_order.resize(N);
    flat.fail.assign((size_t)N + 1, flat.root);
    for (uint32_t idx = 0; idx < N; ++idx) flat.bfs_order[idx] = newid[order[idx]];
    for (uint32_t v = 0; v < N; ++v) flat.fail[newid[v]] = newid[fail[v]];
    flat.level_off = level_off;
    const size_t n_levels = level_off.size() - 1;
    flat.edge_src.clear(); flat.edge_dst.clear(); flat.edge_cls.clear();
    flat.edge_src.reserve(N); flat.edge_dst.reserve(N); flat.edge_cls.reserve(N);
    flat.level_edge_off.assign(n_levels + 1, 0);
    for (size_t d = 0; d < n_levels; ++d) {
        flat.level_edge_off[d] = (uint32_t)flat.edge_src.size();
        for (uint32_t idx = level_off[d]; idx < level_off[d + 1]; ++idx) {
            const uint32_t v = order[idx];
            for (uint32_t k = child_off[v]; k < child_off[v + 1]; ++k) {
                const uint32_t w = children[k];
                flat.edge_src.push_back(newid[v]);
                flat.edge_dst.push_back(newid[w]);
                flat.edge_cls.push_back(flat.cls_map[in_byte_[w]]);
            }
        }
    }
    flat.level_edge_off[n_levels] = (uint32_t)flat.edge_src.size();

    // output lists, indexed by (state - 1)
    flat.max_pattern_len = 0;
    for (const AC_PATTERN_t &p : patterns_)
        flat.max_pattern_len = std::max<uint32_t>(flat.max_pattern_len, (uint32_t)p.ptext.length);
    std::vector<uint32_t> old_of_final(n_final);
    for (uint32_t v = 0; v < N; ++v) if (is_final[v]) old_of_final[newid[v] - 1] = v;
    flat.out_off.assign((size_t)n_final + 1, 0);
    for (uint32_t i = 0; i < n_final; ++i) {
        uint64_t cnt = 0;
        uint32_t v = old_of_final[i];
        if (own_[v] >= 0) ++cnt;
        for (uint32_t d = dlink[v]; d != NONE; d = dlink[d]) ++cnt;
        flat.out_off[i + 1] = flat.out_off[i] + cnt;
    }
    flat.out_pat.resize(flat.out_off[n_final]);
    flat.out_idx.resize(flat.out_off[n_final]);
    flat.accepted = patterns_;
    for (uint32_t i = 0; i < n_final; ++i) {
        uint64_t o = flat.out_off[i];
        uint32_t v = old_of_final[i];
        if (own_[v] >= 0) { flat.out_idx[o] = (uint32_t)own_[v]; flat.out_pat[o++] = patterns_[own_[v]]; }
        for (uint32_t d = dlink[v]; d != NONE; d = dlink[d]) { flat.out_idx[o] = (uint32_t)own_[d]; flat.out_pat[o++] = patterns_[own_[d]]; }
    }

    build_filter(flat);
    build_gram_table(flat);
}

// Gram prefilter tables (see FlatAutomaton).  Every ACCEPTED pattern contributes the W words that can be
// "the last aligned word before the end offset" of one of its occurrences.
void HostTrie::build_filter(FlatAutomaton &flat) const {
    flat.filter_w = 0; flat.l1_bits = 0; flat.l1.clear(); flat.l2_log2 = 0; flat.l2.clear();
    flat.n_grams = 0; flat.l1_fill = 0.0;
    uint32_t min_len = 0;
    for (const AC_PATTERN_t &p : patterns_)
        min_len = min_len ? std::min<uint32_t>(min_len, (uint32_t)p.ptext.length) : (uint32_t)p.ptext.length;
    flat.min_pattern_len = min_len;
    const uint32_t W = (min_len >= 16) ? 8u : (min_len >= 8) ? 4u : 0u;
    if (!W) return;

    // the word at distance r from the pattern's end plus the byte after it (r >= 1: still inside the pattern)
    auto gram = [&](const AC_PATTERN_t &p, uint32_t r, uint32_t &lo, uint32_t &hi, uint32_t &nb) {
        const uint8_t *b = (const uint8_t *)p.ptext.astring + (p.ptext.length - W - r);
        lo = hi = 0;
        memcpy(&lo, b, 4);
        if (W == 8) memcpy(&hi, b + 4, 4);
        nb = b[W];
    };
    flat.filter_w = W;
    flat.l1_bits = FILTER_L1_BITS;
    flat.l1.assign(FILTER_L1_BITS / 32, 0);
    for (const AC_PATTERN_t &p : patterns_)
        for (uint32_t r = 1; r <= W; ++r) {
            uint32_t lo, hi, nb;
            gram(p, r, lo, hi, nb);
            for (uint32_t next : {nb, FILTER_NEXT_UNKNOWN}) {
                const uint32_t t = filter_mix1(lo, hi, next);
                flat.l1[filter_l1_word(t, next == FILTER_NEXT_UNKNOWN)] |= (1u << filter_bit1(t)) | (1u << filter_bit2(t));
            }
        }
    flat.n_grams = (uint64_t)patterns_.size() * W * 2;
    uint64_t set = 0;
    for (uint32_t w : flat.l1) set += (uint64_t)__builtin_popcount(w);
    flat.l1_fill = (double)set / (double)flat.l1_bits;

    // Level 1 tests two bits of one word: a random haystack word passes with probability ~fill^2.
    // Level 2 (global memory) is only worth its latency when that is still not selective.
    double l2_min_fill = 0.10;
    if (const char *e = getenv("ACB200_L2_MIN_FILL")) l2_min_fill = atof(e);
    if (flat.l1_fill > l2_min_fill) {
        uint32_t lg = 23;                                   // at least 1 MiB: ~256 bits per gram, capped at 128 MiB
        while (lg < 30 && (1ull << lg) < flat.n_grams * 256) ++lg;
        flat.l2_log2 = lg;
        flat.l2.assign((size_t)1 << (lg - 5), 0);
        for (const AC_PATTERN_t &p : patterns_)
            for (uint32_t r = 1; r <= W; ++r) {
                uint32_t lo, hi, nb;
                gram(p, r, lo, hi, nb);
                for (uint32_t next : {nb, FILTER_NEXT_UNKNOWN}) {
                    const uint32_t i = filter_l2_index(filter_mix1(lo, hi, next), lg);
                    flat.l2[i >> 5] |= 1u << (i & 31);
                }
            }
    }
}

// Exact gram table (gram_table.hpp), derived from the flat description alone so that a loaded blob gets it too.
void build_gram_table(FlatAutomaton &flat) {
    flat.gt_log2 = 0; flat.gt_slots.clear(); flat.gt_pat.clear(); flat.gt_keys = 0; flat.gt_walk_keys = 0;
    const uint32_t W = flat.filter_w;
    const size_t np = flat.accepted.size();
    if ((W != 4 && W != 8) || np == 0 || np * W > (1ull << 27)) return;
    // (a loaded blob is only structurally validated: make sure of what is indexed below)
    for (const AC_PATTERN_t &p : flat.accepted)
        if (p.ptext.length < 2 * W || p.ptext.length > AC_PATTRN_MAX_LENGTH) return;
    if (flat.level_off.empty() || flat.level_off.back() > flat.bfs_order.size() || flat.fail.size() < flat.n_rows ||
        flat.out_off.size() < flat.final_bound) return;
    for (size_t d = 0; d + 1 < flat.level_off.size(); ++d)
        if (flat.level_off[d] > flat.level_off[d + 1]) return;

    // depth of every state from the breadth-first order; states some other state fails to
    std::vector<uint32_t> depth(flat.n_rows, 0);
    for (size_t d = 0; d + 1 < flat.level_off.size(); ++d)
        for (uint32_t i = flat.level_off[d]; i < flat.level_off[d + 1]; ++i) depth[flat.bfs_order[i]] = (uint32_t)d;
    std::vector<uint8_t> fail_target(flat.n_rows, 0);
    for (uint32_t s : flat.bfs_order)
        if (s != flat.root) fail_target[flat.fail[s]] = 1;

    // the state of each pattern's own node: the final state whose longest output is the pattern at full depth
    std::vector<uint32_t> pat_state(np, 0);
    for (uint32_t s = 1; s < flat.final_bound; ++s) {
        const uint64_t o = flat.out_off[s - 1];
        if (o == flat.out_off[s]) continue;
        const uint32_t pi = flat.out_idx[o];
        if (flat.accepted[pi].ptext.length == depth[s]) pat_state[pi] = s;
    }

    // pattern store (indexed with 32 bits)
    {
        uint64_t words = 0;
        for (const AC_PATTERN_t &p : flat.accepted) words += (p.ptext.length + W - 1) / W * W / 4 + 2;
        if (words >= 0x7fffffffull) return;
    }
    std::vector<uint32_t> pat_ref(np, 0);
    for (size_t pi = 0; pi < np; ++pi) {
        const size_t len = flat.accepted[pi].ptext.length;
        const size_t padded = (len + W - 1) / W * W;
        const size_t at = flat.gt_pat.size();
        flat.gt_pat.resize(at + padded / 4 + 2, 0);               // bytes, state id, one word of padding (8-byte records)
        memcpy((uint8_t *)(flat.gt_pat.data() + at) + (padded - len), flat.accepted[pi].ptext.astring, len);
        pat_ref[pi] = (uint32_t)(at + padded / 4);
        flat.gt_pat[pat_ref[pi]] = pat_state[pi];
    }

    // Load factor <= 1/4 (<= 1/2 for tables beyond 64 MiB): the probe loop of a warp runs as long as its unluckiest
    // lane's, and every step is a dependent trip to L2 — ncu's source view of ac_walk_kernel had a third of its stall
    // samples on that loop at load 1/2 (6.5 steps per warp; absent keys, the Bloom false positives, probe longest).
    uint32_t lg = 10;
    while ((1ull << lg) < 4ull * np * W) ++lg;
    if (((size_t)sizeof(GramSlot) << lg) > ((size_t)64 << 20) && (1ull << (lg - 1)) >= 2ull * np * W) --lg;
    flat.gt_log2 = lg;
    flat.gt_slots.assign((size_t)1 << lg, GramSlot{0, 0, 0, 0, {0, 0, 0, 0}});
    const uint32_t mask = (1u << lg) - 1u;
    for (size_t pi = 0; pi < np; ++pi) {
        const AC_PATTERN_t &p = flat.accepted[pi];
        const uint32_t len = (uint32_t)p.ptext.length;
        // a pattern without a node of its own (cannot happen) or whose node is a failure target: walk
        const bool walk = pat_state[pi] == 0 || fail_target[pat_state[pi]];
        for (uint32_t r = 1; r <= W; ++r) {
            const uint8_t *b = (const uint8_t *)p.ptext.astring + (len - W - r);
            uint32_t lo = 0, hi = 0;
            memcpy(&lo, b, 4);
            if (W == 8) memcpy(&hi, b + 4, 4);
            const uint32_t nb = b[W];
            uint32_t i = gram_home(lo, hi, nb, lg);
            while (true) {
                GramSlot &s = flat.gt_slots[i];
                if (!(s.meta & GRAM_USED)) {
                    const bool inl = len <= 16;
                    s = GramSlot{lo, hi, gram_meta(len, r, nb) | (walk ? GRAM_WALK : 0u) | (inl ? GRAM_INLINE : 0u),
                                 inl ? pat_state[pi] : pat_ref[pi], {0, 0, 0, 0}};
                    const uint32_t nt = std::min<uint32_t>(len, 16u);
                    memcpy((uint8_t *)s.tail + (16 - nt), (const uint8_t *)p.ptext.astring + (len - nt), nt);
                    ++flat.gt_keys;
                    if (walk) ++flat.gt_walk_keys;
                    break;
                }
                if (s.key_lo == lo && s.key_hi == hi && gram_meta_next(s.meta) == nb) {
                    // a second (pattern, r) under one key: more than one candidate
                    if (!(s.meta & GRAM_WALK)) { s.meta |= GRAM_WALK; ++flat.gt_walk_keys; }
                    break;
                }
                i = (i + 1u) & mask;
            }
        }
    }
}

} // namespace acb200
