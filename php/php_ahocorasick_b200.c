/*
 * php_ahocorasick_b200.c — the PHP extension ("ahocorasick") on top of libacb200.so.
 *
 * Same userland API as ph4r05/php_aho_corasick (reference src/php_ahocorasick.stub.php:12-37):
 *   ahocorasick_init(array): resource            ahocorasick_add_patterns(resource, array): bool
 *   ahocorasick_finalize(resource): bool          ahocorasick_match(string, resource, bool findAll = true): array|false
 *   ahocorasick_isValid(resource): bool           ahocorasick_deinit(resource): bool
 * plus   ahocorasick_match_batch(array haystacks, resource, bool findAll = true): array|false
 * which hands an array of haystacks to ONE ac_trie_search_batch() call (all GPUs of the box, see
 * include/acb200.h acb200_set_devices / ACB200_DEVICES) and returns one result array per haystack.
 *
 * What it keeps from the reference's glue, because scripts and tests/test1..6.phpt observe it:
 *   - pattern arrays: keys key | value | id | aux | ignoreCase, case-insensitive, a bare element is the value;
 *     warning / AhoException texts (reference src/php_ahocorasick.c:222,243,255,281,305,311,317,329,404)
 *   - the patterns of one call reach the automaton LAST ELEMENT FIRST, add status ignored (:410-421, 457-486)
 *   - lazy finalize on the first match (:707), add_patterns refused afterwards (:914-917), finalize true once (:130-142)
 *   - result record: pos, key | keyIdx, aux, start_postion (sic), value — in that order (:555-584)
 *   - findAll = false: every pattern of the FIRST event (:588)
 * What it does not keep: the PHP 5 compatibility layer (PHP >= 7.0 only) and the bundled MultiFast sources — the
 * five ac_trie_* calls go to libacb200.so, whose header declares them with the reference's names and layouts.
 *
 * Build (needs php-dev; see php/config.m4):   phpize && ./configure --with-acb200=/path/to/repo && make && make test
 * This image has no PHP toolchain: tests/test_php_extension_source.py compiles this file against the declaration-
 * only headers in php/zend_stub/ (syntax and types), and the same rules are exercised on the GPU through the C++
 * mirror csrc/php_host.cpp (tests/test_php_api_gpu.py replays the six .phpt goldens).
 */
#ifdef HAVE_CONFIG_H
#include "config.h"
#endif

#include "php.h"
#include "ext/standard/info.h"
#include "zend_exceptions.h"

#include "acb200.h"

#define PHP_AHOCORASICK_EXTNAME "ahocorasick"
#define PHP_AHOCORASICK_VERSION "0.0.8-b200"
#define AHO_RES_NAME "AhoCorasick search"

static int le_aho;                       /* resource type */
static zend_class_entry *aho_exception_ce;

/* one pattern as the script gave it; the automaton carries a pointer to it in AC_PATTERN_t.aux */
typedef struct aho_pat {
    zend_string *value;                  /* never NULL once accepted */
    zend_string *key;                    /* string id or NULL */
    zend_long id;                        /* numeric id */
    zval aux;                            /* IS_UNDEF if none */
    enum ac_pattid_type id_type;
} aho_pat_t;

typedef struct aho_handle {
    AC_TRIE_t *trie;
    aho_pat_t **pats;
    size_t n_pats, cap_pats;
    zend_bool init_ok, finalized;
} aho_handle_t;

/* ------------------------------------------------------------------ patterns ---- */

static void aho_pat_free(aho_pat_t *p)
{
    if (!p) return;
    if (p->value) zend_string_release(p->value);
    if (p->key) zend_string_release(p->key);
    if (Z_TYPE(p->aux) != IS_UNDEF) zval_ptr_dtor(&p->aux);
    efree(p);
}

enum { F_KEY = 1, F_VALUE = 2, F_IGNORECASE = 4, F_ID = 8, F_AUX = 16 };

static const struct { const char *name; size_t len; int field; } aho_fields[] = {
    {"key", 3, F_KEY}, {"value", 5, F_VALUE}, {"ignoreCase", 10, F_IGNORECASE}, {"id", 2, F_ID}, {"aux", 3, F_AUX},
};

static int aho_field_of(const zend_string *k)
{
    size_t i;
    if (!k) return F_VALUE;                                  /* a bare element is the value */
    for (i = 0; i < sizeof(aho_fields) / sizeof(aho_fields[0]); i++)
        if (ZSTR_LEN(k) == aho_fields[i].len && zend_binary_strcasecmp(ZSTR_VAL(k), ZSTR_LEN(k), aho_fields[i].name, aho_fields[i].len) == 0)
            return aho_fields[i].field;
    return 0;
}

/* one element of the pattern array -> record; NULL (after a warning or exception) if it is not acceptable */
static aho_pat_t *aho_pat_from_array(zend_long index, HashTable *ht)
{
    aho_pat_t *p = ecalloc(1, sizeof(*p));
    zend_string *k;
    zval *v;
    int seen = 0;
    ZVAL_UNDEF(&p->aux);
    p->id_type = AC_PATTID_TYPE_DEFAULT;
    ZEND_HASH_FOREACH_STR_KEY_VAL(ht, k, v) {
        const int f = aho_field_of(k);
        if (!f) {
            php_error_docref(NULL, E_WARNING, "Invalid structure (unrecognized sub-array key)! Only allowed are: {key, id, value, aux, ignoreCase}. "
                             "Cannot initialize. Pattern index: %ld", (long)index);
            goto fail;
        }
        seen |= f;
        ZVAL_DEREF(v);
        switch (f) {
        case F_ID:
            if (Z_TYPE_P(v) != IS_LONG) {
                zend_throw_exception_ex(aho_exception_ce, 0, "Invalid type of pattern ID given (long required), type: %s, pattern index: %ld",
                                        zend_zval_type_name(v), (long)index);
                goto fail;
            }
            p->id = Z_LVAL_P(v);
            p->id_type = AC_PATTID_TYPE_NUMBER;
            break;
        case F_AUX:
            if (Z_TYPE(p->aux) != IS_UNDEF) zval_ptr_dtor(&p->aux);
            ZVAL_COPY(&p->aux, v);
            break;
        case F_KEY:
        case F_VALUE:
            if (Z_TYPE_P(v) != IS_STRING) {
                zend_throw_exception_ex(aho_exception_ce, 0, "Pattern %s has to be a string, type: %s, pattern index: %ld",
                                        f == F_KEY ? "key" : "value", zend_zval_type_name(v), (long)index);
                goto fail;
            }
            if (f == F_KEY) {
                if (p->key) zend_string_release(p->key);
                p->key = zend_string_copy(Z_STR_P(v));
                p->id_type = AC_PATTID_TYPE_STRING;
            } else {
                if (p->value) zend_string_release(p->value);
                p->value = zend_string_copy(Z_STR_P(v));
            }
            break;
        default:
            break;                                           /* ignoreCase: reported below */
        }
    } ZEND_HASH_FOREACH_END();
    if (!p->value) {
        php_error_docref(NULL, E_WARNING, "No value was specified for pattern index: %ld", (long)index);
        goto fail;
    }
    if ((seen & F_KEY) && (seen & F_ID)) {
        php_error_docref(NULL, E_WARNING, "Pattern can have either numeric or string identifier, not both! Pattern index: %ld", (long)index);
        goto fail;
    }
    if (seen & F_IGNORECASE)
        php_error_docref(NULL, E_WARNING, "ignoreCase attribute is deprecated and is ignored. Pattern index: %ld", (long)index);
    return p;
fail:
    aho_pat_free(p);
    return NULL;
}

/* all patterns of one init()/add_patterns() call: all or nothing */
static int aho_add_call(aho_handle_t *h, HashTable *data)
{
    const uint32_t n = zend_hash_num_elements(data);
    aho_pat_t **list = n ? ecalloc(n, sizeof(*list)) : NULL;
    uint32_t got = 0, k;
    zend_ulong idx;
    zend_string *key;
    zval *v;
    ZEND_HASH_FOREACH_KEY_VAL(data, idx, key, v) {
        ZVAL_DEREF(v);
        if (Z_TYPE_P(v) != IS_ARRAY) {
            php_error_docref(NULL, E_WARNING, "Invalid pattern structure! Cannot initialize.");
            goto fail;
        }
        list[got] = aho_pat_from_array(key ? (zend_long)got : (zend_long)idx, Z_ARRVAL_P(v));
        if (!list[got]) goto fail;
        got++;
    } ZEND_HASH_FOREACH_END();
    if (h->n_pats + got > h->cap_pats) {
        h->cap_pats = (h->n_pats + got) * 2;
        h->pats = h->pats ? erealloc(h->pats, h->cap_pats * sizeof(*h->pats)) : emalloc(h->cap_pats * sizeof(*h->pats));
    }
    /* last element first; what the automaton answers (duplicate, too long, empty) is not looked at */
    for (k = got; k-- > 0;) {
        aho_pat_t *p = list[k];
        AC_PATTERN_t patt;
        memset(&patt, 0, sizeof(patt));
        patt.ptext.astring = ZSTR_VAL(p->value);
        patt.ptext.length = ZSTR_LEN(p->value);
        patt.id.type = p->id_type;
        if (p->id_type == AC_PATTID_TYPE_NUMBER) patt.id.u.number = (long)p->id;
        else if (p->id_type == AC_PATTID_TYPE_STRING) patt.id.u.stringy = ZSTR_VAL(p->key);
        patt.aux = p;
        ac_trie_add(h->trie, &patt, 1);
    }
    for (k = 0; k < got; k++) h->pats[h->n_pats++] = list[k];
    if (list) efree(list);
    return SUCCESS;
fail:
    for (k = 0; k < got; k++) aho_pat_free(list[k]);
    if (list) efree(list);
    return FAILURE;
}

/* ------------------------------------------------------------------ resource ---- */

static void aho_handle_dtor(zend_resource *rsrc)
{
    aho_handle_t *h = (aho_handle_t *)rsrc->ptr;
    size_t i;
    if (!h) return;
    if (h->trie) ac_trie_release(h->trie);               /* frees the device tables too */
    for (i = 0; i < h->n_pats; i++) aho_pat_free(h->pats[i]);
    if (h->pats) efree(h->pats);
    efree(h);
    rsrc->ptr = NULL;
}

static aho_handle_t *aho_fetch(zval *zid)
{
    if (Z_TYPE_P(zid) != IS_RESOURCE || Z_RES_TYPE_P(zid) != le_aho) return NULL;
    return (aho_handle_t *)zend_fetch_resource(Z_RES_P(zid), AHO_RES_NAME, le_aho);
}

static zend_bool aho_finalize_once(aho_handle_t *h)
{
    if (!h->init_ok || h->finalized) return 0;
    h->finalized = 1;
    ac_trie_finalize(h->trie);                            /* flattens the automaton and uploads it to the GPU(s) */
    return 1;
}

/* ------------------------------------------------------------------ results ---- */

/* one reported pattern -> the record array the reference builds (key order is part of the contract) */
static void aho_append_hit(zval *list, const aho_pat_t *p, zend_long pos)
{
    zval rec;
    array_init_size(&rec, 5);
    add_assoc_long(&rec, "pos", pos);
    if (p->id_type == AC_PATTID_TYPE_STRING) add_assoc_str(&rec, "key", zend_string_copy(p->key));
    else if (p->id_type == AC_PATTID_TYPE_NUMBER) add_assoc_long(&rec, "keyIdx", p->id);
    if (Z_TYPE(p->aux) != IS_UNDEF) {
        zval aux;
        ZVAL_COPY(&aux, &p->aux);
        add_assoc_zval(&rec, "aux", &aux);
    }
    add_assoc_long(&rec, "start_postion", pos - (zend_long)ZSTR_LEN(p->value));
    add_assoc_str(&rec, "value", zend_string_copy(p->value));
    add_next_index_zval(list, &rec);
}

typedef struct { zval *result; int stop_after_first; } aho_sink_t;

static int aho_on_match(AC_MATCH_t *m, void *user)
{
    aho_sink_t *s = (aho_sink_t *)user;
    size_t j;
    for (j = 0; j < m->size; j++)
        if (m->patterns[j].aux) aho_append_hit(s->result, (const aho_pat_t *)m->patterns[j].aux, (zend_long)m->position);
    return s->stop_after_first;
}

typedef struct { zval *per_text; } aho_batch_sink_t;

static int aho_on_batch_match(size_t text_idx, AC_MATCH_t *m, void *user)
{
    aho_batch_sink_t *s = (aho_batch_sink_t *)user;
    size_t j;
    for (j = 0; j < m->size; j++)
        if (m->patterns[j].aux) aho_append_hit(&s->per_text[text_idx], (const aho_pat_t *)m->patterns[j].aux, (zend_long)m->position);
    return 0;
}

/* ------------------------------------------------------------------ functions ---- */

ZEND_BEGIN_ARG_INFO_EX(arginfo_aho_init, 0, 0, 1)
    ZEND_ARG_ARRAY_INFO(0, data, 0)
ZEND_END_ARG_INFO()
ZEND_BEGIN_ARG_INFO_EX(arginfo_aho_id, 0, 0, 1)
    ZEND_ARG_INFO(0, id)
ZEND_END_ARG_INFO()
ZEND_BEGIN_ARG_INFO_EX(arginfo_aho_add, 0, 0, 2)
    ZEND_ARG_INFO(0, id)
    ZEND_ARG_ARRAY_INFO(0, patterns, 0)
ZEND_END_ARG_INFO()
ZEND_BEGIN_ARG_INFO_EX(arginfo_aho_match, 0, 0, 2)
    ZEND_ARG_INFO(0, needle)
    ZEND_ARG_INFO(0, id)
    ZEND_ARG_INFO(0, findAll)
ZEND_END_ARG_INFO()
ZEND_BEGIN_ARG_INFO_EX(arginfo_aho_match_batch, 0, 0, 2)
    ZEND_ARG_ARRAY_INFO(0, haystacks, 0)
    ZEND_ARG_INFO(0, id)
    ZEND_ARG_INFO(0, findAll)
ZEND_END_ARG_INFO()

PHP_FUNCTION(ahocorasick_init)
{
    zval *data;
    aho_handle_t *h;
    if (zend_parse_parameters(ZEND_NUM_ARGS(), "a", &data) == FAILURE) RETURN_FALSE;
    h = ecalloc(1, sizeof(*h));
    h->trie = ac_trie_create();
    if (!h->trie || aho_add_call(h, Z_ARRVAL_P(data)) != SUCCESS) {
        if (h->trie) ac_trie_release(h->trie);
        efree(h);
        RETURN_FALSE;
    }
    h->init_ok = 1;
    RETURN_RES(zend_register_resource(h, le_aho));
}

PHP_FUNCTION(ahocorasick_add_patterns)
{
    zval *zid, *data;
    aho_handle_t *h;
    if (zend_parse_parameters(ZEND_NUM_ARGS(), "ra", &zid, &data) == FAILURE) RETURN_FALSE;
    h = aho_fetch(zid);
    if (!h || !h->init_ok) {
        php_error_docref(NULL, E_WARNING, "Cannot add a new pattern, not initialized");
        RETURN_FALSE;
    }
    if (h->finalized) {
        php_error_docref(NULL, E_WARNING, "Cannot add a new pattern to finalized search structure");
        RETURN_FALSE;
    }
    RETURN_BOOL(aho_add_call(h, Z_ARRVAL_P(data)) == SUCCESS);
}

PHP_FUNCTION(ahocorasick_finalize)
{
    zval *zid;
    aho_handle_t *h;
    if (zend_parse_parameters(ZEND_NUM_ARGS(), "r", &zid) == FAILURE) RETURN_FALSE;
    h = aho_fetch(zid);
    if (!h) RETURN_FALSE;
    RETURN_BOOL(aho_finalize_once(h));
}

/* common entry checks of match / match_batch: warnings of the reference, lazy finalize, a usable device */
static aho_handle_t *aho_ready(zval *zid)
{
    aho_handle_t *h = aho_fetch(zid);
    if (!h) { php_error_docref(NULL, E_WARNING, "Invalid resource."); return NULL; }
    if (!h->init_ok) { php_error_docref(NULL, E_WARNING, "Not initialized."); return NULL; }
    aho_finalize_once(h);
    return h;
}

PHP_FUNCTION(ahocorasick_match)
{
    zend_string *hay;
    zval *zid;
    zend_bool find_all = 1;
    aho_handle_t *h;
    AC_TEXT_t text;
    aho_sink_t sink;
    if (zend_parse_parameters(ZEND_NUM_ARGS(), "Sr|b", &hay, &zid, &find_all) == FAILURE) RETURN_FALSE;
    if (!(h = aho_ready(zid))) RETURN_FALSE;
    array_init(return_value);
    text.astring = ZSTR_VAL(hay);                         /* borrowed, binary safe */
    text.length = ZSTR_LEN(hay);
    sink.result = return_value;
    sink.stop_after_first = find_all ? 0 : 1;
    if (ac_trie_search(h->trie, &text, 0, aho_on_match, &sink) < 0) {
        /* the reference's search cannot fail; a GPU can (no device, out of memory): say so instead of returning "no match" */
        php_error_docref(NULL, E_WARNING, "GPU search failed: %s", acb200_last_error());
        zval_ptr_dtor(return_value);
        RETURN_FALSE;
    }
}

PHP_FUNCTION(ahocorasick_match_batch)
{
    zval *hays, *zid, *v;
    zend_bool find_all = 1;
    aho_handle_t *h;
    AC_TEXT_t *texts;
    zend_string **held;
    aho_batch_sink_t sink;
    uint32_t n, i = 0;
    int rc;
    if (zend_parse_parameters(ZEND_NUM_ARGS(), "ar|b", &hays, &zid, &find_all) == FAILURE) RETURN_FALSE;
    if (!(h = aho_ready(zid))) RETURN_FALSE;
    n = zend_hash_num_elements(Z_ARRVAL_P(hays));
    texts = ecalloc(n ? n : 1, sizeof(*texts));
    held = ecalloc(n ? n : 1, sizeof(*held));
    sink.per_text = ecalloc(n ? n : 1, sizeof(zval));
    ZEND_HASH_FOREACH_VAL(Z_ARRVAL_P(hays), v) {
        held[i] = zval_get_string(v);                     /* non-strings are converted like (string) would */
        texts[i].astring = ZSTR_VAL(held[i]);
        texts[i].length = ZSTR_LEN(held[i]);
        array_init(&sink.per_text[i]);
        i++;
    } ZEND_HASH_FOREACH_END();
    rc = ac_trie_search_batch(h->trie, texts, n, find_all ? 0 : 1, aho_on_batch_match, &sink);
    if (rc == 0) {
        zend_ulong idx;
        zend_string *key;
        array_init_size(return_value, n);
        i = 0;
        ZEND_HASH_FOREACH_KEY(Z_ARRVAL_P(hays), idx, key) {   /* the result keeps the keys of the haystack array */
            if (key) zend_hash_update(Z_ARRVAL_P(return_value), key, &sink.per_text[i]);
            else zend_hash_index_update(Z_ARRVAL_P(return_value), idx, &sink.per_text[i]);
            i++;
        } ZEND_HASH_FOREACH_END();
    } else {
        php_error_docref(NULL, E_WARNING, "GPU search failed: %s", acb200_last_error());
        for (i = 0; i < n; i++) zval_ptr_dtor(&sink.per_text[i]);
        ZVAL_FALSE(return_value);
    }
    for (i = 0; i < n; i++) zend_string_release(held[i]);
    efree(held);
    efree(texts);
    efree(sink.per_text);
}

PHP_FUNCTION(ahocorasick_isValid)
{
    zval *zid;
    aho_handle_t *h;
    if (zend_parse_parameters(ZEND_NUM_ARGS(), "r", &zid) == FAILURE) RETURN_FALSE;
    h = aho_fetch(zid);
    RETURN_BOOL(h && h->init_ok);
}

PHP_FUNCTION(ahocorasick_deinit)
{
    zval *zid;
    aho_handle_t *h;
    if (zend_parse_parameters(ZEND_NUM_ARGS(), "r", &zid) == FAILURE) RETURN_FALSE;
    h = aho_fetch(zid);                                   /* NULL for a handle that was closed before: false, like the reference */
    if (!h) RETURN_FALSE;
    h->init_ok = 0;
    zend_list_close(Z_RES_P(zid));                        /* runs aho_handle_dtor: device and host memory are gone now */
    RETURN_TRUE;
}

static const zend_function_entry ahocorasick_functions[] = {
    PHP_FE(ahocorasick_match, arginfo_aho_match)
    PHP_FE(ahocorasick_match_batch, arginfo_aho_match_batch)
    PHP_FE(ahocorasick_init, arginfo_aho_init)
    PHP_FE(ahocorasick_deinit, arginfo_aho_id)
    PHP_FE(ahocorasick_isValid, arginfo_aho_id)
    PHP_FE(ahocorasick_finalize, arginfo_aho_id)
    PHP_FE(ahocorasick_add_patterns, arginfo_aho_add)
    PHP_FE_END
};

PHP_MINIT_FUNCTION(ahocorasick)
{
    zend_class_entry ce;
    le_aho = zend_register_list_destructors_ex(aho_handle_dtor, NULL, AHO_RES_NAME, module_number);
    INIT_CLASS_ENTRY(ce, "AhoException", NULL);
    aho_exception_ce = zend_register_internal_class_ex(&ce, zend_ce_exception);
    return SUCCESS;
}

PHP_MINFO_FUNCTION(ahocorasick)
{
    php_info_print_table_start();
    php_info_print_table_row(2, "ahocorasick support", "enabled (B200 / libacb200)");
    php_info_print_table_row(2, "extension version", PHP_AHOCORASICK_VERSION);
    php_info_print_table_row(2, "matcher", acb200_version());
    php_info_print_table_end();
}

zend_module_entry ahocorasick_module_entry = {
    STANDARD_MODULE_HEADER,
    PHP_AHOCORASICK_EXTNAME,
    ahocorasick_functions,
    PHP_MINIT(ahocorasick),
    NULL, NULL, NULL,
    PHP_MINFO(ahocorasick),
    PHP_AHOCORASICK_VERSION,
    STANDARD_MODULE_PROPERTIES
};

#ifdef COMPILE_DL_AHOCORASICK
ZEND_GET_MODULE(ahocorasick)
#endif
