/*
 * acb200.h — C-ABI of the B200-native Aho-Corasick matcher (libacb200.so).
 *
 * This is the drop-in boundary for the hot path of ph4r05/php_aho_corasick:
 * ahocorasick_match() -> ac_trie_search().  The PHP extension's Zend glue
 * (reference src/php_ahocorasick.c) reaches its bundled matcher through exactly
 * five calls; this header declares those five with the same names, argument
 * meaning, return codes and value-type layouts, so the glue compiles against it
 * unchanged.  Every declaration cites the reference
 * interface it replaces as  file:line  relative to the reference tree.
 *
 * The handle (AC_TRIE_t) is opaque here: the glue never dereferences it
 * (reference src/php_ahocorasick.c only stores it in ahocorasick_master_t.acap,
 * src/php_ahocorasick.h:181).
 *
 * Plain C, plain pointers and sizes; no CUDA or torch types cross this line.
 */
#ifndef ACB200_H_
#define ACB200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------ *
 * Value types crossing the seam.  Layouts are identical to the reference's  *
 * (src/multifast/actypes.h:41-138) so that callers built against either     *
 * header are binary compatible.                                             *
 * ------------------------------------------------------------------------ */

/* replaces src/multifast/actypes.h:41 */
typedef char AC_ALPHABET_t;

/* replaces src/multifast/actypes.h:47-51 — borrowed byte string, binary safe */
typedef struct ac_text
{
    const AC_ALPHABET_t *astring;
    size_t length;
} AC_TEXT_t;

/* replaces src/multifast/actypes.h:57-62 */
enum ac_pattid_type
{
    AC_PATTID_TYPE_DEFAULT = 0,
    AC_PATTID_TYPE_NUMBER,
    AC_PATTID_TYPE_STRING
};

/* replaces src/multifast/actypes.h:68-77 */
typedef struct ac_pattid
{
    union
    {
        const char *stringy;
        long number;
    } u;
    enum ac_pattid_type type;
} AC_PATTID_t;

/* replaces src/multifast/actypes.h:83-89 (56 bytes on LP64) */
typedef struct ac_pattern
{
    AC_TEXT_t ptext;   /* the search string                                  */
    AC_TEXT_t rtext;   /* replacement string — carried, never interpreted     */
    AC_PATTID_t id;    /* returned verbatim in match events                   */
    void *aux;         /* caller's opaque pointer, returned verbatim          */
} AC_PATTERN_t;

/* replaces src/multifast/actypes.h:106-113 — one event: every pattern that
 * ends at `position` (exclusive end offset), longest pattern first.         */
typedef struct ac_match
{
    AC_PATTERN_t *patterns;
    size_t size;
    size_t position;
} AC_MATCH_t;

/* replaces src/multifast/actypes.h:118-125 */
typedef enum ac_status
{
    ACERR_SUCCESS = 0,
    ACERR_DUPLICATE_PATTERN,
    ACERR_LONG_PATTERN,
    ACERR_ZERO_PATTERN,
    ACERR_TRIE_CLOSED
} AC_STATUS_t;

/* replaces src/multifast/actypes.h:138 — non-zero return stops the search   */
typedef int (*AC_MATCH_CALBACK_f)(AC_MATCH_t *, void *);

/* replaces src/multifast/actypes.h:148 — longer patterns are rejected; this
 * constant changes results, so it is part of the contract.                  */
#define AC_PATTRN_MAX_LENGTH 1024

/* replaces src/multifast/ahocorasick.h:37-66 — opaque in this library       */
typedef struct ac_trie AC_TRIE_t;

/* ------------------------------------------------------------------------ *
 * The five entry points the Zend glue binds.                                *
 * ------------------------------------------------------------------------ */

/* replaces src/multifast/ahocorasick.h:73 (called at php_ahocorasick.c:812).
 * Returns an open (not finalized) automaton; NULL only on host OOM.         */
AC_TRIE_t *ac_trie_create(void);

/* replaces src/multifast/ahocorasick.h:74 (called at php_ahocorasick.c:484).
 * Acceptance rules of src/multifast/ahocorasick.c:91-131: TRIE_CLOSED after
 * finalize, ZERO_PATTERN for length 0, LONG_PATTERN for length > 1024,
 * DUPLICATE_PATTERN when the same bytes were accepted before (first wins).
 * copy!=0: bytes and string id are copied (binary safe); copy==0: borrowed.  */
AC_STATUS_t ac_trie_add(AC_TRIE_t *thiz, AC_PATTERN_t *patt, int copy);

/* replaces src/multifast/ahocorasick.h:75 (called at php_ahocorasick.c:140).
 * Computes failure/output links, flattens the automaton into the dense
 * transition table and uploads it to the GPU (once).  On a CUDA failure the
 * automaton is closed but unusable; see acb200_last_error().                */
void ac_trie_finalize(AC_TRIE_t *thiz);

/* replaces src/multifast/ahocorasick.h:79-80 (called at php_ahocorasick.c:745).
 * Returns -1 if not finalized (or on a device error — acb200_last_error()
 * tells which), 0 if the text was scanned to its end, 1 if the callback
 * stopped the search.  keep==0 starts from the root at offset 0; keep!=0
 * continues from the state and base offset the previous call ended at
 * (src/multifast/ahocorasick.c:162-163,191-194,236-238).  Events arrive in
 * ascending position, at most one per position.                             */
int ac_trie_search(AC_TRIE_t *thiz, AC_TEXT_t *text, int keep,
                   AC_MATCH_CALBACK_f callback, void *user);

/* replaces src/multifast/ahocorasick.h:76 (called at php_ahocorasick.c:504,821).
 * Frees host and device memory of the automaton.                            */
void ac_trie_release(AC_TRIE_t *thiz);

/* ------------------------------------------------------------------------ *
 * Additions (not in the reference).                                         *
 * ------------------------------------------------------------------------ */

/* Batched callback: like AC_MATCH_CALBACK_f plus the index of the haystack.
 * Non-zero return stops the search of THAT haystack only.                   */
typedef int (*ACB200_BATCH_CALLBACK_f)(size_t text_idx, AC_MATCH_t *, void *);

/* Searches n independent haystacks in one call; every haystack
 * starts at the root at offset 0 (the keep=0 rule of php_ahocorasick.c:745).
 * Large batches are gathered slab by slab into pinned staging memory and
 * pipelined over every GPU of the handle (acb200_set_devices).
 * first_only!=0 reports only the first event of each haystack
 * (php_ahocorasick.c:588 — findAll=false).  Callbacks arrive ordered by
 * (text_idx, position).  Returns 0, or -1 on error.                         */
int ac_trie_search_batch(AC_TRIE_t *thiz, const AC_TEXT_t *texts, size_t n,
                         int first_only, ACB200_BATCH_CALLBACK_f callback,
                         void *user);

/* Same, for haystacks already laid end to end in one host buffer:
 * haystack i is bytes [offsets[i], offsets[i+1]) of `bytes` (n+1 offsets).
 * Avoids the gather copy; `bytes` may be pinned (acb200_host_alloc).        */
int ac_trie_search_flat(AC_TRIE_t *thiz, const char *bytes,
                        const uint64_t *offsets, size_t n, int first_only,
                        ACB200_BATCH_CALLBACK_f callback, void *user);

/* One raw match event as the device produces it. `end` is the exclusive end
 * offset inside haystack `text_idx`; `state` indexes the automaton's output
 * lists (acb200_state_patterns).                                            */
typedef struct acb200_event
{
    uint64_t end;
    uint32_t state;
    uint32_t text_idx;
} ACB200_EVENT_t;

/* Event-level search (no callback replay): fills up to `cap` events sorted by
 * (text_idx, end) and stores the total number found in *n_events (which may
 * exceed cap).  Inputs as ac_trie_search_flat.  Returns 0 / -1.             */
int acb200_search_events(AC_TRIE_t *thiz, const char *bytes,
                         const uint64_t *offsets, size_t n, int first_only,
                         ACB200_EVENT_t *events, size_t cap, size_t *n_events);

/* Device-resident variant: `d_bytes` is a CUDA device pointer on the
 * automaton's device holding total = offsets[n] bytes (offsets is a HOST
 * array).  Events stay on the device: *d_events receives a device pointer to
 * packed {uint32 end_in_buffer, uint32 state} records (library-owned, valid
 * until the next call on this handle), sorted by buffer offset.  `stream` is
 * a cudaStream_t passed as void* (NULL = the handle's own stream).  The call
 * returns after the kernels are enqueued and the event count has been read
 * back.  Used by bench.py's kernel-only leg and by callers that already hold
 * haystacks in HBM.  Returns 0 / -1.                                        */
int acb200_search_device(AC_TRIE_t *thiz, const void *d_bytes,
                         const uint64_t *offsets, size_t n, int first_only,
                         void *stream, const void **d_events,
                         size_t *n_events);

/* Same for a batch of n haystacks of EQUAL length hay_len laid end to end (no offsets array:
 * nothing on the host is proportional to n).                                          */
int acb200_search_device_uniform(AC_TRIE_t *thiz, const void *d_bytes, size_t n,
                                 size_t hay_len, int first_only, void *stream,
                                 const void **d_events, size_t *n_events);

/* Asynchronous variant for callers that chain more device work behind the scan on `stream` (the multi-GPU event
 * gather sends the rows on without a host round trip): returns after the kernels are enqueued, waits for nothing.
 * `d_rows` is a DEVICE buffer of 1 + max_events 8-byte rows: row 0 receives {uint32 event count, uint32 densely
 * flagged tiles}, rows 1.. the first max_events packed events in ascending order (a count above max_events means
 * the caller has to repeat the call with more rows).  Only the prefilter path supports this: -1 (and an error
 * text) if the dictionary, the batch size or the automatic density rule would choose the full walk — use the
 * synchronous call then.  `stream` = NULL means the legacy default stream here (the kernels must be ordered with
 * the caller's work, so the library's private stream is never used).  After the caller has waited for `stream`, acb200_async_finish(count) completes the
 * statistics of acb200_last_stats(); the library's own event buffer is not touched.                      */
int acb200_search_device_uniform_async(AC_TRIE_t *thiz, const void *d_bytes, size_t n, size_t hay_len,
                                       void *d_rows, size_t max_events, void *stream);
int acb200_async_finish(AC_TRIE_t *thiz, size_t n_events, size_t dense_tiles);

/* One reported pattern occurrence, as the reference's callback would have recorded it
 * (src/php_ahocorasick.c:555-584): haystack index, exclusive end offset inside the haystack
 * ("pos"), start offset ("start_postion" = pos - length) and the pattern's index in acceptance
 * order (acb200_pattern() returns its AC_PATTERN_t: id, aux, bytes).                       */
typedef struct acb200_hit
{
    uint32_t text_idx;
    uint32_t end;
    uint32_t start;
    uint32_t pattern;
} ACB200_HIT_t;

/* Hit-level search: the events are expanded to hits ON THE DEVICE (every pattern of every event,
 * longest first inside an event, events in ascending order) and only the hit columns come back.
 * Fills up to `cap` hits and stores the total in *n_hits (which may exceed cap).  findAll only.
 * Inputs as ac_trie_search_flat.  Returns 0 / -1.                                           */
int acb200_search_hits(AC_TRIE_t *thiz, const char *bytes, const uint64_t *offsets, size_t n,
                       ACB200_HIT_t *hits, size_t cap, size_t *n_hits);

/* The hits of the most recent acb200_search_hits() call again (its events are still on the device): for a caller
 * whose buffer was too small the first time — the haystack is not scanned again.                       */
int acb200_last_hits(AC_TRIE_t *thiz, ACB200_HIT_t *hits, size_t cap, size_t *n_hits);

/* Accepted pattern number `index` (acceptance order = the order of successful ac_trie_add calls);
 * NULL if out of range or not finalized.                                                     */
const AC_PATTERN_t *acb200_pattern(const AC_TRIE_t *thiz, size_t index);

/* Writes the finalized automaton (flat description, output lists, prefilter tables, accepted patterns with
 * their ids; aux pointers are stored as opaque integers) to `path`.  The reference has no counterpart: a trie
 * cannot be serialised, so every request pays init + finalize again.  Returns 0 / -1.              */
int acb200_save(const AC_TRIE_t *thiz, const char *path);

/* Loads such a file: a finalized automaton, ready to search (only the device part of finalize is replayed).
 * Returns NULL on a missing / corrupt file (acb200_last_error()); like after ac_trie_finalize, a handle whose
 * device setup failed is returned but unusable (acb200_info().device < 0).                           */
AC_TRIE_t *acb200_load(const char *path);

/* Copies up to max_events packed {uint32 end_in_buffer, uint32 state} records of the most recent
 * device-resident search into the caller's DEVICE buffer `d_dst` (async on `stream`, NULL = the
 * handle's stream).  Returns the number of records copied, or -1.                      */
long acb200_copy_events(AC_TRIE_t *thiz, void *d_dst, size_t max_events, void *stream);

/* A ready-made consumer for callers that only need totals: pass acb200_tally_cb as the
 * ACB200_BATCH_CALLBACK_f (or acb200_tally_match_cb as the AC_MATCH_CALBACK_f) and an
 * ACB200_TALLY_t as `user`.  `hash` folds (text_idx, position, size, first and last pattern's
 * aux) in callback order, so equal hashes mean equal event sequences.                 */
typedef struct acb200_tally
{
    uint64_t events;
    uint64_t hits;
    uint64_t hash;
} ACB200_TALLY_t;
int acb200_tally_cb(size_t text_idx, AC_MATCH_t *m, void *tally);
int acb200_tally_match_cb(AC_MATCH_t *m, void *tally);

/* Per-haystack digest of an event list (as returned by acb200_search_events): counts[h] = events of haystack h,
 * hashes[h] = fold over its events, in order, of (position, number of patterns, first and last pattern's aux) —
 * the value a caller of the reference gets by folding the same fields inside its AC_MATCH_CALBACK_f, so two
 * implementations can be compared haystack by haystack without keeping every hit.  Returns 0 / -1.          */
int acb200_event_digest(const AC_TRIE_t *thiz, const ACB200_EVENT_t *events, size_t n_events, size_t n_texts,
                        uint64_t *counts, uint64_t *hashes);

/* Patterns reported by automaton state `state` (longest first); returns the
 * count and stores a library-owned array in *patterns (NULL if none).       */
size_t acb200_state_patterns(const AC_TRIE_t *thiz, uint32_t state,
                             const AC_PATTERN_t **patterns);

/* Automaton facts after finalize. */
typedef struct acb200_info
{
    uint64_t n_patterns;      /* accepted patterns                           */
    uint64_t n_states;        /* automaton states incl. root                 */
    uint32_t n_classes;       /* byte classes = table columns                */
    uint32_t entry_bytes;     /* 2 or 4                                      */
    uint32_t max_pattern_len; /* Lmax of accepted patterns                   */
    uint32_t final_bound;     /* event states are in [1, final_bound)        */
    uint32_t root;            /* id of the root state                        */
    uint64_t table_bytes;     /* dense transition table in HBM               */
    int32_t device;           /* CUDA device ordinal                         */
    int32_t finalized;
    int32_t filter_word;      /* prefilter word size W (8 or 4), 0 = dictionary not eligible */
    uint32_t min_pattern_len; /* shortest accepted pattern                   */
    float filter_l1_fill;     /* fraction of level-1 prefilter bits set      */
    uint32_t filter_l2_log2;  /* log2(bits) of the level-2 bitmap, 0 = none  */
    uint32_t reserved_;
    uint32_t direct_keys;      /* distinct grams in the exact gram table (0 = no table)                  */
    uint32_t direct_walk_keys; /* of those, grams that cannot be decided by one comparison (are walked)  */
} ACB200_INFO_t;
int acb200_info(const AC_TRIE_t *thiz, ACB200_INFO_t *out);

/* Statistics of the most recent search on this handle. */
typedef struct acb200_stats
{
    uint64_t bytes;           /* haystack bytes scanned                      */
    uint64_t events;          /* events found                                */
    uint64_t kernel_launches; /* this library's kernels launched             */
    uint32_t chunk_bytes;     /* bytes per thread slice                      */
    uint32_t halo_bytes;      /* overlap re-read before each slice           */
    float kernel_ms;          /* device time of the scan kernel(s)           */
    float h2d_ms, d2h_ms;     /* copies, 0 for the device-resident entry     */
    uint32_t devices;         /* GPUs that took part in the call (acb200_set_devices) */
    uint32_t filtered;        /* 1: gram prefilter + verify kernels, 0: full automaton walk */
    float filter_ms;          /* device time of the prefilter kernel (0 if unused) */
    float verify_ms;          /* device time of the verify kernel (0 if unused) */
    uint64_t flagged_words;   /* aligned words the prefilter handed to verification */
    uint64_t dense_tiles;     /* 16 KiB tiles handed to verification as whole spans */
    float reorder_ms;         /* device time of the offsets + emit kernels (0 if unused) */
    float expand_ms;          /* device time of the hit expansion kernels (acb200_search_hits) */
} ACB200_STATS_t;
int acb200_last_stats(const AC_TRIE_t *thiz, ACB200_STATS_t *out);

/* Out-of-band error text of the last failed call on this thread ("" if none).
 * The reference's void/ignored returns leave no room for CUDA errors.       */
const char *acb200_last_error(void);

/* Primary device of automata CREATED afterwards by this thread
 * (default: env ACB200_DEVICE, else the current CUDA device).               */
int acb200_set_device(int device);
int acb200_device_count(void);

/* The GPUs one host call may use.  After finalize (or acb200_load) the automaton lives on its primary device
 * (acb200_set_device at create time); this call replicates it onto the other listed devices — the expanded
 * table is copied from the primary's HBM over NVLink — and from then on ac_trie_search, ac_trie_search_batch,
 * ac_trie_search_flat and acb200_search_events cut large inputs into slabs spread over all of them, balanced by
 * bytes, one pinned double-buffered H2D pipeline per GPU, events returned in the same global order as on one
 * GPU.  A cut may fall inside a haystack (the slab then carries the Lmax-1 bytes before it), so ONE large
 * haystack is spread as well.  The list replaces any earlier one; the primary device always takes part, and an
 * ordinal listed k times gets k pipelines (its own scratch and streams each).
 * Environment ACB200_DEVICES=all|0,1,... does the same at finalize for callers that cannot make this call
 * (the PHP extension).  Returns 0 / -1.                                                                      */
int acb200_set_devices(AC_TRIE_t *thiz, const int *devices, size_t n);

/* Environment read by the library — the complete list: ACB200_DEVICE (primary GPU of new handles), ACB200_DEVICES
 * (GPUs a host call may use, above), ACB200_L2_MIN_FILL (test knob: level-1 fill above which finalize builds the
 * level-2 bitmap of the prefilter, default 0.10; results do not depend on it), ACB200_GATHER_THREADS (threads that
 * gather the scattered strings of an ac_trie_search_batch() slab into pinned staging, 1..64; default
 * min(12, 3/4 cores / GPUs); results do not depend on it), ACB200_GATHER_NT (0: that gather uses plain memcpy
 * instead of non-temporal stores, for A/B measurements), ACB200_STAGE_MIN (bytes from which a pageable text or the
 * strings of a batch that fit one launch are copied into pinned staging by the handle's helper threads, piece by
 * piece with each piece's DMA queued behind it, instead of by one cudaMemcpyAsync on the pageable pointer; default
 * 3 MiB; results do not depend on it). */

/* Bytes per slab of the host pipeline (0 = default 64 MiB).  Tests use small slabs to force cuts. */
int acb200_set_slab_bytes(AC_TRIE_t *thiz, uint64_t bytes);

/* Diagnostic: the slab plan of a call (csrc/shard.hpp) evaluated on the host — no GPU involved.  Slab i reports
 * the events that end in stream bytes (begin, end], travels with `halo` bytes in front, runs on device slot
 * `device_slot` and touches the haystacks [first_text, end_text).                                            */
typedef struct acb200_slab
{
    uint64_t begin, end;
    uint32_t halo, device_slot;
    uint64_t first_text, end_text;
} ACB200_SLAB_t;
int acb200_plan_slabs(const uint64_t *offsets, size_t n, uint32_t halo_max, int n_devices, uint64_t slab_bytes,
                      ACB200_SLAB_t *out, size_t cap, size_t *n_slabs);

/* Peer-memory plumbing for callers that run ONE PROCESS PER GPU (torchrun) and collect every rank's events on one
 * of them: a device buffer of the collecting rank is mapped into the other processes through CUDA IPC
 * (export -> 64-byte handle -> open), each rank copies its rows into its slot of it with the copy engines
 * (acb200_copy_async on a side stream: no collective kernel competes with the scan for SMs, the transfer of step k
 * overlaps the scan of step k+1) and then writes the step number into its mailbox word, behind the rows on the same
 * stream; acb200_mailbox_wait_async enqueues a one-warp kernel that returns once all n mailboxes (stride_words
 * apart) have reached `seq` (step numbers only grow; the words may live in a peer's memory).  php_aho_corasick_b200/dist.py (MailboxGatherer) is the user.  0 / -1, NULL on failure.   */
void *acb200_device_alloc(int device, size_t bytes);            /* cudaMalloc + zero-fill on `device` */
int acb200_device_free(int device, void *p);
int acb200_ipc_export(const void *dptr, unsigned char handle[64]);
void *acb200_ipc_open(int device, const unsigned char handle[64]);
int acb200_ipc_close(int device, void *p);
int acb200_copy_async(void *dst, const void *src, size_t bytes, void *stream);
int acb200_mailbox_wait_async(int device, const void *mailboxes, uint32_t n, uint32_t stride_words, uint32_t seq,
                              void *stream);

/* The mailbox gather built from those pieces.  Every rank creates one object over the collector's two buffers as they
 * are mapped in ITS process (rows: 2 x world x cap_rows x 8 bytes; mailboxes: (2 x world x 4 + 2) words, zeroed).
 * acb200_mailbox_step() is one step of one rank: scans n equal-length haystacks resident at d_bytes (the kernels of
 * acb200_search_device_uniform_async, on `stream`), waits for this rank's own event count, then puts its rows and
 * {step, count} on their way to the collector on a copy stream and returns the count (-1: error, e.g. more events
 * than cap_rows).  The transfer overlaps whatever the caller enqueues next; a sender never overwrites a slot the
 * collector still holds (per-parity acknowledgement).  On the collector acb200_mailbox_result(step) waits until
 * every rank's rows of that step have landed and returns their counts: rank r's rows are rows[step & 1][r][0..count)
 * and stay valid until step + 2 is sent.  acb200_mailbox_drain() makes `stream` wait for this rank's pending copies. */
typedef struct acb200_mailbox ACB200_MAILBOX_t;
ACB200_MAILBOX_t *acb200_mailbox_create(AC_TRIE_t *thiz, int rank, int world, int collector, size_t cap_rows,
                                        void *rows, void *mailboxes);
long acb200_mailbox_step(ACB200_MAILBOX_t *m, const void *d_bytes, size_t n, size_t hay_len, void *stream);
int acb200_mailbox_result(ACB200_MAILBOX_t *m, uint32_t step, uint32_t *counts);
int acb200_mailbox_drain(ACB200_MAILBOX_t *m, void *stream);
void acb200_mailbox_free(ACB200_MAILBOX_t *m);

/* Pinned host memory for haystacks (optional; pageable memory works too).   */
void *acb200_host_alloc(size_t bytes);
void acb200_host_free(void *p);

/* Tuning knobs (0 = automatic). chunk_bytes: bytes per thread slice. */
int acb200_set_tuning(AC_TRIE_t *thiz, uint32_t chunk_bytes,
                      uint32_t smem_table_bytes);

/* Gram prefilter (dictionaries whose accepted patterns are all >= 8 bytes): 0 = automatic,
 * 1 = use it whenever the dictionary allows, -1 = never (always walk the full automaton).
 * Results are identical either way.                                                     */
int acb200_set_filter(AC_TRIE_t *thiz, int mode);

/* Full walk: how the haystack text reaches the walking lanes.  0 = automatic (TMA boxes into shared memory where the
 * shape allows it and the previous call found few events, else 16-byte loads), 1 = TMA wherever the shape allows,
 * -1 = loads only.  Results are identical either way. */
int acb200_set_tma(AC_TRIE_t *thiz, int mode);

/* Diagnostic: the prefilter's decision for one aligned word, evaluated on the HOST from the tables finalize
 * built (the same hashes the filter kernel uses).  `word` = the W bytes little-endian (W from
 * acb200_info().filter_word; upper bytes zero for W = 4), `next_byte` = the byte after the word, or 0x100 for
 * "unknown".  Returns 1 if the word would be handed to verification, 0 if not, -1 if the dictionary has no
 * prefilter.  Not a matching path: tests use it to check that no occurrence can be filtered away.        */
int acb200_filter_probe(const AC_TRIE_t *thiz, uint64_t word, unsigned next_byte);

/* Direct verification of flagged words (gram_table.hpp): a flagged word whose (word, next byte) belongs to exactly
 * one (pattern, alignment) is decided by comparing the haystack with that pattern instead of walking the
 * automaton.  0 = automatic (currently 1), 1 = on, -1 = off (every flagged word is walked).  Results are
 * identical in every mode. */
int acb200_set_direct(AC_TRIE_t *thiz, int mode);

/* Diagnostic: the direct verification of the aligned word `word_index` (bytes [W*word_index, W*word_index + W) of
 * the stream `bytes`) evaluated on the HOST with the code the kernels run; the haystack that contains the W end
 * offsets after the word starts at stream offset `hay_begin` (bytes before it belong to another haystack).
 * Returns 0 = nothing ends at those offsets, 1 = exactly one event (*end = its exclusive end offset in the
 * stream, *state = the automaton state the event carries), 2 = undecided (the kernel walks the automaton),
 * -1 = not applicable (no table, or the window [W*(word_index+1) - warm, W*(word_index+2)) does not lie inside
 * the stream).  Not a matching path: tests check the table construction against the CPU oracle with it. */
int acb200_direct_probe(const AC_TRIE_t *thiz, const char *bytes, size_t length, size_t hay_begin, size_t word_index,
                        uint32_t *end, uint32_t *state);

const char *acb200_version(void);

#ifdef __cplusplus
}
#endif

#endif /* ACB200_H_ */
