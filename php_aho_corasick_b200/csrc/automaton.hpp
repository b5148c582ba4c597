// automaton.hpp — host side of ahocorasick_finalize(): byte trie with the
// reference's acceptance rules, breadth-first failure/output links, and the
// flat description the device expands into the dense transition table.
//
// Semantics follow (not the code of) the reference:
//   acceptance rules        src/multifast/ahocorasick.c:91-131
//   failure link definition src/multifast/ahocorasick.c:344-368
//   output lists + finality src/multifast/node.c:424-441
#pragma once

#include <cstdint>
#include <cstddef>
#include <string>
#include <vector>
#include <deque>

#include "acb200.h"
#include "gram_table.hpp"

namespace acb200 {

// Open-addressing map (state, byte) -> child used while patterns are added.
class EdgeMap {
public:
    EdgeMap();
    uint32_t find(uint32_t node, uint8_t byte) const;   // NONE if absent
    void insert(uint32_t node, uint8_t byte, uint32_t child);
    size_t size() const { return used_; }
    static constexpr uint32_t NONE = 0xffffffffu;
private:
    void grow();
    std::vector<uint64_t> keys_;   // ((node << 8) | byte) + 1 ; 0 = empty slot
    std::vector<uint32_t> vals_;
    size_t used_ = 0;
    size_t mask_ = 0;
};

// Flat automaton in device state numbering (see HostTrie::flatten): id 0 is a
// reserved marker, ids [1, final_bound) are the pattern-reporting ("final")
// states, ids [final_bound, n_rows) the others, root == final_bound.
struct FlatAutomaton {
    uint32_t n_states = 1;        // real states incl. root
    uint32_t n_rows = 2;          // table rows = n_states + 1 (row 0 unused)
    uint32_t n_classes = 1;       // table columns
    uint32_t final_bound = 1;     // states < final_bound (and != 0) report patterns
    uint32_t root = 1;
    uint32_t max_pattern_len = 0; // Lmax over accepted patterns
    uint32_t n_used_bytes = 0;
    uint8_t  cls_map[256];        // byte -> column; unused bytes share the last column
    bool     range_map = true;    // used bytes are one contiguous range [range_lo, range_lo+n_used)
    uint32_t range_lo = 0;

    // breadth-first description consumed by the device expansion kernels
    std::vector<uint32_t> bfs_order;       // state ids level by level (root first)
    std::vector<uint32_t> level_off;       // level d = bfs_order[level_off[d] .. level_off[d+1])
    std::vector<uint32_t> fail;            // failure state per state id
    std::vector<uint32_t> edge_src;        // trie edges grouped by depth of src (same grouping as level_off)
    std::vector<uint32_t> edge_dst;
    std::vector<uint16_t> edge_cls;
    std::vector<uint32_t> level_edge_off;  // edges leaving level d = [level_edge_off[d], level_edge_off[d+1])

    // output lists for final states (index: state - 1), longest pattern first
    std::vector<uint64_t> out_off;
    std::vector<AC_PATTERN_t> out_pat;
    std::vector<uint32_t> out_idx;         // same order as out_pat: index of the pattern in acceptance order
    std::vector<AC_PATTERN_t> accepted;    // accepted patterns in acceptance order (what out_idx indexes)

    // Gram prefilter (filter_kernels.cuh).  With every accepted pattern at least 2W bytes long
    // (W = 8 or 4), a match that ends at stream offset p contains the aligned W-byte word k with
    // W(k+1) < p <= W(k+2) entirely inside its last 2W bytes.  `l1` is a bitmap over the hashes of
    // the W such words of every pattern (pattern[L-W-r, L-r), r = 1..W): a haystack word whose bit
    // is clear cannot be that word of any match, so only the W end offsets after a flagged word have
    // to be verified with the automaton.  `l2` is a second, larger bitmap over an independent hash,
    // built only when the first level is too full to be selective.
    uint32_t min_pattern_len = 0;
    uint32_t filter_w = 0;             // 0: dictionary not eligible (some pattern shorter than 8 bytes)
    uint32_t l1_bits = 0;
    std::vector<uint32_t> l1;          // l1_bits / 32 words
    uint32_t l2_log2 = 0;              // level 2 holds 2^l2_log2 bits; 0: not used
    std::vector<uint32_t> l2;
    uint64_t n_grams = 0;
    double l1_fill = 0.0;              // fraction of level-1 bits set

    // Exact gram table (gram_table.hpp): decides most flagged words with one comparison instead of a walk.
    // Derived from the fields above by build_gram_table(); not part of the blob.
    uint32_t gt_log2 = 0;              // 2^gt_log2 slots; 0: no table
    std::vector<GramSlot> gt_slots;
    std::vector<uint32_t> gt_pat;      // per pattern: its bytes, zero-padded in FRONT to a multiple of W, then its state id
    uint64_t gt_keys = 0, gt_walk_keys = 0;
};

// Fills the gt_* fields from the rest of `flat` (needs filter_w, accepted, out lists, fail, bfs order).
void build_gram_table(FlatAutomaton &flat);

// blob.cpp: position-independent dump of a finalized automaton
bool save_flat(const FlatAutomaton &flat, const char *path, std::string &err);
bool load_flat(FlatAutomaton &flat, std::deque<std::string> &arena, const char *path, std::string &err);

class HostTrie {
public:
    HostTrie();
    AC_STATUS_t add(const AC_PATTERN_t *patt, int copy);
    // Computes links and fills `flat`. Idempotent guard is the caller's job.
    void flatten(FlatAutomaton &flat);

    size_t n_nodes() const { return parent_.size(); }
    size_t n_patterns() const { return patterns_.size(); }
    void release_build_memory();

private:
    const char *keep_bytes(const char *p, size_t n);
    void build_filter(FlatAutomaton &flat) const;

    std::vector<uint32_t> parent_;
    std::vector<uint8_t>  in_byte_;
    std::vector<uint16_t> depth_;
    std::vector<int32_t>  own_;          // index into patterns_ or -1
    EdgeMap edges_;
    std::vector<AC_PATTERN_t> patterns_; // accepted patterns, acceptance order
    std::deque<std::string> arena_;      // owned copies of pattern bytes / string ids
};

} // namespace acb200
