/*
 * acb200_php.h — host-side mirror of the reference's PHP-level API for the hot path.
 *
 * The reference's L3/L2 layer is a Zend extension (src/php_ahocorasick.c); PHP
 * and its headers are not available where this library is built, so the six
 * userland functions are restated here in C++ behind a C interface with the
 * same names, argument meaning, return values, warnings and exception texts.
 * A PHP array is modelled by aho_array_t (ordered key/value pairs), a zval by
 * aho_value_t.  The real extension glue a maintainer would build against
 * libacb200.so is shown in INTEGRATION.md; this mirror exists so that parity
 * tests can be written exactly like the reference's the phpt files under tests/.
 *
 * Everything below sits ABOVE the C-ABI of acb200.h and reaches the matcher
 * only through ac_trie_create/add/finalize/release and the batched search.
 */
#ifndef ACB200_PHP_H_
#define ACB200_PHP_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* zval type tags, named after php_aho_type_str() (src/php_ahocorasick.c:98-117) */
enum aho_type {
    AHO_T_NULL = 0, AHO_T_FALSE, AHO_T_TRUE, AHO_T_LONG, AHO_T_DOUBLE,
    AHO_T_STRING, AHO_T_ARRAY, AHO_T_OBJECT, AHO_T_RESOURCE
};

struct aho_array;

typedef struct aho_value {
    int type;                 /* enum aho_type                                        */
    long lval;                /* AHO_T_LONG                                           */
    double dval;              /* AHO_T_DOUBLE                                         */
    const char *sval;         /* AHO_T_STRING: bytes (binary safe)                    */
    size_t slen;
    const struct aho_array *aval; /* AHO_T_ARRAY: nested array (pattern element)      */
    void *opaque;             /* caller's handle for the value; returned verbatim as
                                 "aux" / "key" / "value" in results                   */
} aho_value_t;

typedef struct aho_entry {
    const char *key;          /* string key, or NULL for an integer key               */
    size_t key_len;
    long index;               /* integer key when key == NULL                          */
    aho_value_t val;
} aho_entry_t;

typedef struct aho_array {
    const aho_entry_t *entries;
    size_t n;
} aho_array_t;

/* E_WARNING texts and the AhoException message produced by a call */
typedef struct aho_diag {
    char warnings[8][256];
    int n_warnings;
    char exception[512];      /* non-empty: the reference would throw AhoException     */
} aho_diag_t;

/* "AhoCorasick search" resource (src/php_ahocorasick.h:179-190) */
typedef struct aho_master aho_master_t;

/* one element of the array ahocorasick_match() returns
 * (src/php_ahocorasick.c:555-584; key order pos, key|keyIdx, aux, start_postion, value) */
typedef struct aho_hit {
    long pos;
    int key_type;             /* 0 none, 1 keyIdx (long), 2 key (string)               */
    long key_idx;
    void *key_opaque;         /* the caller's handle of the 'key' string value         */
    int has_aux;
    void *aux_opaque;         /* the caller's handle of the 'aux' value                */
    long start_postion;       /* sic — the reference's spelling                        */
    void *value_opaque;       /* the caller's handle of the 'value' string             */
    const char *value;        /* library-owned copy of the value bytes                 */
    size_t value_len;
} aho_hit_t;

typedef struct aho_result {
    int is_false;             /* the PHP function returned false                       */
    aho_hit_t *hits;
    size_t n;
} aho_result_t;

/* ahocorasick_init(array $data): resource|false   — src/php_ahocorasick.c:798-838 */
aho_master_t *ahocorasick_init(const aho_array_t *data, aho_diag_t *diag);
/* ahocorasick_add_patterns(resource, array): bool   — :882-925 */
int ahocorasick_add_patterns(aho_master_t *m, const aho_array_t *data, aho_diag_t *diag);
/* ahocorasick_finalize(resource): bool              — :845-875 */
int ahocorasick_finalize(aho_master_t *m, aho_diag_t *diag);
/* ahocorasick_match(string, resource, bool findAll = true): array|false — :664-746.
 * The result is owned by the caller; free with aho_result_free(). */
aho_result_t *ahocorasick_match(const char *haystack, size_t len, aho_master_t *m, int find_all,
                                aho_diag_t *diag);
/* ahocorasick_match_batch(array $haystacks, resource, bool findAll = true): array of arrays — new.
 * results[i] receives the match array of haystack i. Returns 0, or -1 (false + warning). */
int ahocorasick_match_batch(const char *const *haystacks, const size_t *lens, size_t n,
                            aho_master_t *m, int find_all, aho_result_t **results, aho_diag_t *diag);
/* ahocorasick_isValid(resource): bool               — :623-655 */
int ahocorasick_isValid(const aho_master_t *m);
/* ahocorasick_deinit(resource): bool                — :754-791.  After a successful deinit the
 * handle stays allocated as a closed resource (isValid false, second deinit false) until
 * aho_resource_free(), which stands in for the Zend resource destructor. */
int ahocorasick_deinit(aho_master_t *m, aho_diag_t *diag);
void aho_resource_free(aho_master_t *m);
void aho_result_free(aho_result_t *r);
/* the underlying AC_TRIE_t* (for stats / tuning in tests and bench) */
void *aho_master_trie(aho_master_t *m);

#ifdef __cplusplus
}
#endif

#endif /* ACB200_PHP_H_ */
