/*
 * ac_oracle.c — CPU ORACLE.  TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C restatement of the algorithm on the hot path of
 * ph4r05/php_aho_corasick (the bundled MultiFast 2.0 automaton), written from
 * the reference's behaviour, exporting the same five entry points with the
 * value types of include/acb200.h.  Only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load it — as the checker
 * or the reported CPU baseline, never as a product path.  libacb200.so does
 * not link, load or call anything in this directory.
 *
 * Parity pin: tests/test_oracle.py checks this file against the golden vectors
 * transcribed from the reference's own tests/test1..6.phpt (tests/golden/) and,
 * when /root/reference is present, event-for-event against the reference's own
 * C sources compiled in place into oracle/_ref/ (see oracle/Makefile).
 *
 * Each function cites the reference file:line it follows (paths relative to
 * the reference tree).
 */
#include <stdlib.h>
#include <string.h>

#include "acb200.h"

typedef struct onode onode_t;

typedef struct oedge {
    signed char alpha;          /* src/multifast/node.h:63-67 — AC_ALPHABET_t is (signed) char */
    onode_t *next;
} oedge_t;

struct onode {
    int final;                  /* src/multifast/node.h:41 */
    size_t depth;               /* :42 */
    onode_t *failure;           /* :43 */
    oedge_t *out;               /* :45-47 */
    size_t out_n, out_cap;
    AC_PATTERN_t *matched;      /* :49-51 — own pattern, then failure-chain patterns */
    size_t matched_n, matched_cap;
    int has_own;                /* matched[0] is this node's own pattern */
};

struct ac_trie {
    onode_t *root;
    size_t patterns_count;
    int open;                   /* src/multifast/ahocorasick.h:45 trie_open */
    onode_t *last_node;         /* :51 */
    size_t base_position;       /* :52-57 */
    /* bookkeeping for release */
    onode_t **all; size_t all_n, all_cap;
    char **blobs; size_t blobs_n, blobs_cap;
};

static onode_t *onode_new(AC_TRIE_t *t)
{
    onode_t *n = (onode_t *)calloc(1, sizeof(onode_t));
    if (t->all_n == t->all_cap) {
        t->all_cap = t->all_cap ? t->all_cap * 2 : 256;
        t->all = (onode_t **)realloc(t->all, t->all_cap * sizeof(onode_t *));
    }
    t->all[t->all_n++] = n;
    return n;
}

static char *keep_blob(AC_TRIE_t *t, const char *p, size_t n)
{
    char *c = (char *)malloc(n + 1);
    if (n) memcpy(c, p, n);
    c[n] = 0;
    if (t->blobs_n == t->blobs_cap) {
        t->blobs_cap = t->blobs_cap ? t->blobs_cap * 2 : 256;
        t->blobs = (char **)realloc(t->blobs, t->blobs_cap * sizeof(char *));
    }
    t->blobs[t->blobs_n++] = c;
    return c;
}

/* linear edge lookup used while the trie is open — src/multifast/node.c:98-108 */
static onode_t *find_next_linear(const onode_t *n, signed char a)
{
    size_t i;
    for (i = 0; i < n->out_n; i++)
        if (n->out[i].alpha == a) return n->out[i].next;
    return NULL;
}

/* binary search over edges sorted by signed char — src/multifast/node.c:119-140 */
static onode_t *find_next_bs(const onode_t *n, signed char a)
{
    size_t lo = 0, hi = n->out_n;
    while (lo < hi) {
        size_t mid = (lo + hi) >> 1;
        signed char m = n->out[mid].alpha;
        if (m == a) return n->out[mid].next;
        if (m < a) lo = mid + 1; else hi = mid;
    }
    return NULL;
}

static void push_matched(onode_t *n, const AC_PATTERN_t *p)
{
    if (n->matched_n == n->matched_cap) {
        n->matched_cap = n->matched_cap ? n->matched_cap * 2 : 1;
        n->matched = (AC_PATTERN_t *)realloc(n->matched, n->matched_cap * sizeof(AC_PATTERN_t));
    }
    n->matched[n->matched_n++] = *p;
}

/* src/multifast/ahocorasick.c:59-77 */
AC_TRIE_t *ac_trie_create(void)
{
    AC_TRIE_t *t = (AC_TRIE_t *)calloc(1, sizeof(AC_TRIE_t));
    t->root = onode_new(t);
    t->open = 1;
    t->last_node = t->root;
    t->base_position = 0;
    return t;
}

/* src/multifast/ahocorasick.c:91-131 — statuses in the reference's order:
 * closed, empty, too long; the path is created BEFORE the duplicate check. */
AC_STATUS_t ac_trie_add(AC_TRIE_t *t, AC_PATTERN_t *patt, int copy)
{
    size_t i;
    onode_t *n = t->root, *next;
    AC_PATTERN_t rec;

    if (!t->open) return ACERR_TRIE_CLOSED;
    if (!patt->ptext.length) return ACERR_ZERO_PATTERN;
    if (patt->ptext.length > AC_PATTRN_MAX_LENGTH) return ACERR_LONG_PATTERN;

    for (i = 0; i < patt->ptext.length; i++) {
        signed char a = (signed char)patt->ptext.astring[i];
        next = find_next_linear(n, a);
        if (!next) {
            next = onode_new(t);
            next->depth = n->depth + 1;
            if (n->out_n == n->out_cap) {
                n->out_cap = n->out_cap ? n->out_cap * 2 : 2;
                n->out = (oedge_t *)realloc(n->out, n->out_cap * sizeof(oedge_t));
            }
            n->out[n->out_n].alpha = a;
            n->out[n->out_n].next = next;
            n->out_n++;
        }
        n = next;
    }
    if (n->final) return ACERR_DUPLICATE_PATTERN;

    n->final = 1;
    rec = *patt;
    if (copy) {   /* src/multifast/node.c:238-261 (deep copy; made binary safe here) */
        rec.ptext.astring = keep_blob(t, patt->ptext.astring, patt->ptext.length);
        rec.rtext.astring = patt->rtext.length ? keep_blob(t, patt->rtext.astring, patt->rtext.length) : NULL;
        if (patt->id.type == AC_PATTID_TYPE_STRING && patt->id.u.stringy)
            rec.id.u.stringy = keep_blob(t, patt->id.u.stringy, strlen(patt->id.u.stringy));
    }
    push_matched(n, &rec);      /* src/multifast/node.c:205-229 */
    n->has_own = 1;
    t->patterns_count++;
    return ACERR_SUCCESS;
}

/* src/multifast/ahocorasick.c:344-368 — the failure node is the deepest node
 * whose path equals a proper suffix of this node's path: try suffixes from
 * the longest (start 1) to the shortest, walking each from the root. */
static void set_failure(AC_TRIE_t *t, onode_t *node, const signed char *prefix)
{
    size_t i, j;
    onode_t *n;
    if (node == t->root) return;
    for (i = 1; i < node->depth; i++) {
        n = t->root;
        for (j = i; j < node->depth && n; j++) n = find_next_linear(n, prefix[j]);
        if (n) { node->failure = n; break; }
    }
    if (!node->failure) node->failure = t->root;
}

/* src/multifast/ahocorasick.c:381-396 — depth-first, prefix[] carries the path */
static void traverse_setfailure(AC_TRIE_t *t, onode_t *node, signed char *prefix)
{
    size_t i;
    set_failure(t, node, prefix);
    for (i = 0; i < node->out_n; i++) {
        prefix[node->depth] = node->out[i].alpha;
        traverse_setfailure(t, node->out[i].next, prefix);
    }
}

static int edge_cmp(const void *l, const void *r)
{
    /* src/multifast/node.c:305-315 — signed char order */
    return (int)((const oedge_t *)l)->alpha - (int)((const oedge_t *)r)->alpha;
}

/* src/multifast/node.c:424-441.  The reference appends, for every node n on
 * the failure chain, all of n->matched[] that this node does not hold yet
 * (equal length + equal bytes, node.c:150-175).  A chain node's list is its
 * own pattern followed by patterns of nodes further down the same chain, so
 * the de-duplicated result is exactly: own pattern, then the OWN pattern of
 * each chain node in chain order (strictly decreasing length).  Stated that
 * way here, which also avoids the reference's quartic blow-up on nested
 * patterns.  `final` is inherited from any accepting chain node. */
static void collect_matches(onode_t *nod)
{
    onode_t *n = nod;
    while ((n = n->failure)) {
        if (n->has_own) push_matched(nod, &n->matched[0]);
        if (n->final) nod->final = 1;
    }
    if (nod->out_n > 1) qsort(nod->out, nod->out_n, sizeof(oedge_t), edge_cmp);   /* node.c:322-326 */
}

/* src/multifast/ahocorasick.c:408-422 with top_down = 1.  A node's `final`
 * may be read by descendants' chains before or after it inherited finality;
 * the reference has the same order dependence and the same result, because a
 * chain is walked to the root and every accepting node on it is seen. */
static void traverse_collect(onode_t *node)
{
    size_t i;
    collect_matches(node);
    for (i = 0; i < node->out_n; i++) traverse_collect(node->out[i].next);
}

/* src/multifast/ahocorasick.c:143-155 */
void ac_trie_finalize(AC_TRIE_t *t)
{
    signed char prefix[AC_PATTRN_MAX_LENGTH];
    if (!t->open) return;
    traverse_setfailure(t, t->root, prefix);
    traverse_collect(t->root);
    t->open = 0;
}

/* src/multifast/ahocorasick.c:175-241 — the hot loop, restated:
 * goto if an edge exists, else follow the failure link without consuming the
 * byte, else (at the root) consume it; report only after a goto transition
 * into a final node. */
int ac_trie_search(AC_TRIE_t *t, AC_TEXT_t *text, int keep, AC_MATCH_CALBACK_f callback, void *user)
{
    size_t position = 0;
    onode_t *current, *next;
    AC_MATCH_t match;

    if (t->open) return -1;
    if (!keep) {                 /* ac_trie_reset, src/multifast/ahocorasick.c:330-335 */
        t->last_node = t->root;
        t->base_position = 0;
    }
    current = t->last_node;

    while (position < text->length) {
        next = find_next_bs(current, (signed char)text->astring[position]);
        if (!next) {
            if (current->failure) current = current->failure;
            else position++;
        } else {
            current = next;
            position++;
        }
        if (current->final && next) {
            match.position = position + t->base_position;
            match.size = current->matched_n;
            match.patterns = current->matched;
            if (callback(&match, user)) return 1;
        }
    }
    t->last_node = current;
    t->base_position += position;
    return 0;
}

/* src/multifast/ahocorasick.c:288-296 */
void ac_trie_release(AC_TRIE_t *t)
{
    size_t i;
    for (i = 0; i < t->all_n; i++) {
        free(t->all[i]->out);
        free(t->all[i]->matched);
        free(t->all[i]);
    }
    for (i = 0; i < t->blobs_n; i++) free(t->blobs[i]);
    free(t->all);
    free(t->blobs);
    free(t);
}
