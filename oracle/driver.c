/*
 * driver.c — TEST INFRASTRUCTURE.  One driver, three backends.
 *
 * Written once against the five-call MultiFast API and compiled
 *   (a) with the reference's own sources from /root/reference   -> oracle/_ref/libref_driver.so
 *   (b) with oracle/ac_oracle.c (the CPU restatement)            -> oracle/liboracle_driver.so
 *   (c) against include/acb200.h and linked to libacb200.so      -> oracle/libgpu_driver.so
 * so parity tests run byte-identical driver code over the reference, the
 * oracle and the CUDA product.  (c) is also the proof that a caller written
 * for the reference's C API compiles and links against the new library
 * unchanged.
 *
 * It restates the four things the reference's Zend glue does around the API:
 *   1. patterns of one init/add_patterns call are added in REVERSE array order,
 *      add status ignored          (src/php_ahocorasick.c:410-421, 457-486)
 *   2. every search uses keep = 0  (src/php_ahocorasick.c:745)
 *   3. per event, patterns j = 0..size are reported in order with
 *      pos = event position        (src/php_ahocorasick.c:555-584)
 *   4. findAll = false returns 1 from the callback after the first event
 *                                  (src/php_ahocorasick.c:588)
 */
#ifdef DRV_USE_REFERENCE_HEADERS
#include "ahocorasick.h"
#else
#include "acb200.h"
#endif

#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

typedef struct drv {
    AC_TRIE_t *trie;
    size_t n_added;     /* ordinals handed out so far (accepted or not) */
    int last_rc;
} drv_t;

typedef struct sink {
    uint64_t *hit_pos; uint32_t *hit_pat; uint32_t *hit_len; size_t cap;
    uint64_t n_hits, n_events, hash;
    int first_only;
} sink_t;

static inline uint64_t fold(uint64_t h, uint64_t v)
{
    h ^= v + 0x9e3779b97f4a7c15ULL + (h << 6) + (h >> 2);
    return h * 0xff51afd7ed558ccdULL;
}

static int on_match(AC_MATCH_t *m, void *param)
{
    sink_t *s = (sink_t *)param;
    size_t j;
    s->n_events++;
    /* order-sensitive event hash: position, size, first and last pattern */
    s->hash = fold(s->hash, (uint64_t)m->position);
    s->hash = fold(s->hash, (uint64_t)m->size);
    if (m->size) {
        s->hash = fold(s->hash, (uint64_t)(uintptr_t)m->patterns[0].aux);
        s->hash = fold(s->hash, (uint64_t)(uintptr_t)m->patterns[m->size - 1].aux);
    }
    for (j = 0; j < m->size; j++) {
        if (s->n_hits < s->cap) {
            if (s->hit_pos) s->hit_pos[s->n_hits] = (uint64_t)m->position;
            if (s->hit_pat) s->hit_pat[s->n_hits] = (uint32_t)((uintptr_t)m->patterns[j].aux - 1);
            if (s->hit_len) s->hit_len[s->n_hits] = (uint32_t)m->patterns[j].ptext.length;
        }
        s->n_hits++;
    }
    return s->first_only ? 1 : 0;
}

drv_t *drv_create(void)
{
    drv_t *d = (drv_t *)calloc(1, sizeof(drv_t));
    d->trie = ac_trie_create();
    return d;
}

static int add_one(drv_t *d, const char *bytes, size_t len, size_t ordinal)
{
    AC_PATTERN_t p;
    memset(&p, 0, sizeof(p));
    p.ptext.astring = bytes;
    p.ptext.length = len;
    p.rtext.astring = NULL;
    p.rtext.length = 0;
    p.id.type = AC_PATTID_TYPE_NUMBER;
    p.id.u.number = (long)ordinal;
    p.aux = (void *)(uintptr_t)(ordinal + 1);
    return (int)ac_trie_add(d->trie, &p, 1);
}

/* one pattern, ordinal = running count; returns the AC_STATUS_t */
int drv_add(drv_t *d, const char *bytes, size_t len)
{
    return add_one(d, bytes, len, d->n_added++);
}

/* one PHP-level init()/add_patterns() call: array element i gets ordinal
 * base+i, but elements are added LAST FIRST; statuses are ignored. */
void drv_add_php_order(drv_t *d, const char *flat, const uint64_t *off, size_t n)
{
    size_t k;
    const size_t base = d->n_added;
    for (k = n; k-- > 0;) add_one(d, flat + off[k], (size_t)(off[k + 1] - off[k]), base + k);
    d->n_added += n;
}

void drv_finalize(drv_t *d) { ac_trie_finalize(d->trie); }

/* Returns the number of hits (may exceed cap). */
long drv_search(drv_t *d, const char *text, size_t len, int keep, int first_only,
                uint64_t *hit_pos, uint32_t *hit_pat, uint32_t *hit_len, size_t cap,
                uint64_t *n_events, uint64_t *hash)
{
    sink_t s;
    AC_TEXT_t t;
    memset(&s, 0, sizeof(s));
    s.hit_pos = hit_pos; s.hit_pat = hit_pat; s.hit_len = hit_len; s.cap = cap;
    s.first_only = first_only;
    t.astring = text; t.length = len;
    d->last_rc = ac_trie_search(d->trie, &t, keep, on_match, &s);
    if (n_events) *n_events = s.n_events;
    if (hash) *hash = s.hash;
    return (long)s.n_hits;
}

int drv_last_rc(const drv_t *d) { return d->last_rc; }

void drv_release(drv_t *d)
{
    ac_trie_release(d->trie);
    free(d);
}

/* ---------------------------------------------------------------------- *
 * CPU baseline timing: `threads` workers, each with a PRIVATE trie replica
 * (the reference trie carries mutable search state,
 * src/multifast/ahocorasick.h:49-65) that it builds and finalizes itself
 * (side by side: the reference's finalize takes 16 s for 100 k signatures).
 * A batch is split into contiguous blocks of haystacks; ONE large haystack
 * is split into disjoint byte slices, each walked from the root over the
 * `halo` = Lmax-1 bytes before it, events counted only when they end inside
 * the slice (SURVEY.md 8d "CPU baseline").  Only the searches are timed
 * (finalize excluded).  Returns the wall-clock seconds of the best
 * repetition; *events = events seen in one pass.  With hay_events /
 * hay_hash (n_hay entries each, batches only) every haystack's event count
 * and order-sensitive hash (the fold of on_match) are stored as well.
 * ---------------------------------------------------------------------- */
typedef struct work {
    drv_t *d;
    const char *pat_flat; const uint64_t *pat_off; size_t n_pat;
    const char *text; const uint64_t *hay_off; size_t h0, h1;
    uint64_t b0, b1, skip;          /* slice mode: bytes [b0, b1) of haystack 0, events ending at <= skip ignored */
    int slice_mode;
    uint64_t events;
    uint64_t *hay_events, *hay_hash;
    pthread_barrier_t *start, *stop;
    int reps;
} work_t;

typedef struct tally { uint64_t events, hash, skip; } tally_t;

static int on_count(AC_MATCH_t *m, void *param)
{
    tally_t *t = (tally_t *)param;
    if ((uint64_t)m->position > t->skip) t->events++;
    return 0;
}

static int on_digest(AC_MATCH_t *m, void *param)
{
    tally_t *t = (tally_t *)param;
    t->events++;
    t->hash = fold(t->hash, (uint64_t)m->position);
    t->hash = fold(t->hash, (uint64_t)m->size);
    if (m->size) {
        t->hash = fold(t->hash, (uint64_t)(uintptr_t)m->patterns[0].aux);
        t->hash = fold(t->hash, (uint64_t)(uintptr_t)m->patterns[m->size - 1].aux);
    }
    return 0;
}

static void *worker(void *arg)
{
    work_t *w = (work_t *)arg;
    int r;
    size_t h;
    w->d = drv_create();
    drv_add_php_order(w->d, w->pat_flat, w->pat_off, w->n_pat);
    drv_finalize(w->d);
    for (r = 0; r < w->reps; r++) {
        pthread_barrier_wait(w->start);
        w->events = 0;
        if (w->slice_mode) {
            tally_t t = {0, 0, w->skip};
            AC_TEXT_t x;
            x.astring = w->text + (w->b0 - w->skip);
            x.length = (size_t)(w->b1 - w->b0 + w->skip);
            ac_trie_search(w->d->trie, &x, 0, on_count, &t);
            w->events = t.events;
        } else
        for (h = w->h0; h < w->h1; h++) {
            tally_t t = {0, 0, 0};
            AC_TEXT_t x;
            x.astring = w->text + w->hay_off[h];
            x.length = (size_t)(w->hay_off[h + 1] - w->hay_off[h]);
            ac_trie_search(w->d->trie, &x, 0, w->hay_hash ? on_digest : on_count, &t);
            w->events += t.events;
            if (w->hay_events) w->hay_events[h] = t.events;
            if (w->hay_hash) w->hay_hash[h] = t.hash;
        }
        pthread_barrier_wait(w->stop);
    }
    return NULL;
}

double drv_bench2(const char *pat_flat, const uint64_t *pat_off, size_t n_pat,
                  const char *text, const uint64_t *hay_off, size_t n_hay,
                  int threads, int reps, uint64_t halo, uint64_t *events,
                  uint64_t *hay_events, uint64_t *hay_hash)
{
    pthread_t *th;
    work_t *w;
    pthread_barrier_t start, stop;
    double best = 1e30;
    int i, r;
    const int slice_mode = (n_hay == 1 && threads > 1);
    if (threads < 1) threads = 1;
    if (!slice_mode && (size_t)threads > n_hay && n_hay) threads = (int)n_hay;
    th = (pthread_t *)calloc(threads, sizeof(pthread_t));
    w = (work_t *)calloc(threads, sizeof(work_t));
    pthread_barrier_init(&start, NULL, threads + 1);
    pthread_barrier_init(&stop, NULL, threads + 1);
    for (i = 0; i < threads; i++) {
        w[i].pat_flat = pat_flat; w[i].pat_off = pat_off; w[i].n_pat = n_pat;
        w[i].text = text; w[i].hay_off = hay_off;
        w[i].h0 = n_hay * (size_t)i / threads;
        w[i].h1 = n_hay * (size_t)(i + 1) / threads;
        w[i].slice_mode = slice_mode;
        if (slice_mode) {
            const uint64_t total = hay_off[1] - hay_off[0];
            w[i].b0 = hay_off[0] + total * (uint64_t)i / threads;
            w[i].b1 = hay_off[0] + total * (uint64_t)(i + 1) / threads;
            w[i].skip = (w[i].b0 - hay_off[0] < halo) ? w[i].b0 - hay_off[0] : halo;
        }
        w[i].hay_events = slice_mode ? NULL : hay_events;
        w[i].hay_hash = slice_mode ? NULL : hay_hash;
        w[i].start = &start; w[i].stop = &stop; w[i].reps = reps;
        pthread_create(&th[i], NULL, worker, &w[i]);
    }
    for (r = 0; r < reps; r++) {
        struct timespec a, b;
        double dt;
        pthread_barrier_wait(&start);
        clock_gettime(CLOCK_MONOTONIC, &a);
        pthread_barrier_wait(&stop);
        clock_gettime(CLOCK_MONOTONIC, &b);
        dt = (b.tv_sec - a.tv_sec) + 1e-9 * (b.tv_nsec - a.tv_nsec);
        if (dt < best) best = dt;
    }
    if (events) *events = 0;
    for (i = 0; i < threads; i++) {
        pthread_join(th[i], NULL);
        if (events) *events += w[i].events;
        drv_release(w[i].d);
    }
    pthread_barrier_destroy(&start);
    pthread_barrier_destroy(&stop);
    free(th); free(w);
    return best;
}

double drv_bench(const char *pat_flat, const uint64_t *pat_off, size_t n_pat,
                 const char *text, const uint64_t *hay_off, size_t n_hay,
                 int threads, int reps, uint64_t *events)
{
    return drv_bench2(pat_flat, pat_off, n_pat, text, hay_off, n_hay, threads, reps, 0, events, NULL, NULL);
}
