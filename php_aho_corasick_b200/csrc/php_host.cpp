// php_host.cpp — mirror of the reference's PHP-level functions (see include/acb200_php.h).
//
// Restates the observable rules of the Zend glue in src/php_ahocorasick.c: pattern
// array validation (:195-336), reverse insertion order (:410-421, 457-486), resource
// life cycle (:130-142, 494-512, 754-925) and the result record (:542-589).  It talks to
// the matcher only through the C-ABI of acb200.h.
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <string>
#include <strings.h>
#include <vector>

#include "acb200.h"
#include "acb200_php.h"

namespace {

// ahocorasick_pattern_t (src/php_ahocorasick.h:159-174)
struct PatternRec {
    int key_type = AC_PATTID_TYPE_DEFAULT;
    long key_id = 0;
    std::string key; void *key_opaque = nullptr; bool has_key = false;
    std::string value; void *value_opaque = nullptr; bool has_value = false;
    bool has_aux = false; void *aux_opaque = nullptr;
};

const char *type_str(int t)   // php_aho_type_str, src/php_ahocorasick.c:98-117
{
    switch (t) {
        case AHO_T_NULL: return "null";
        case AHO_T_FALSE: return "false";
        case AHO_T_TRUE: return "true";
        case AHO_T_LONG: return "long";
        case AHO_T_DOUBLE: return "double";
        case AHO_T_STRING: return "string";
        case AHO_T_ARRAY: return "array";
        case AHO_T_OBJECT: return "object";
        case AHO_T_RESOURCE: return "resource";
        default: return "undef";
    }
}

void warn(aho_diag_t *d, const char *fmt, ...)
{
    if (!d || d->n_warnings >= 8) return;
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(d->warnings[d->n_warnings], sizeof(d->warnings[0]), fmt, ap);
    va_end(ap);
    d->n_warnings++;
}

void throw_aho(aho_diag_t *d, const char *fmt, ...)
{
    if (!d) return;
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(d->exception, sizeof(d->exception), fmt, ap);
    va_end(ap);
}

// The fields a pattern array may carry (keys are matched case-insensitively; an element without a string key is
// the value).  Rules observed from the reference (src/php_ahocorasick.c:195-336): unknown key -> warning, the call
// fails; id must be an integer and key / value strings, else AhoException; a value is mandatory; key and id exclude
// each other; ignoreCase is accepted with a deprecation warning and has no effect.
enum Field { FIELD_NONE, FIELD_KEY, FIELD_VALUE, FIELD_IGNORECASE, FIELD_ID, FIELD_AUX };

Field field_of(const aho_entry_t &e)
{
    static const struct { const char *name; Field f; } table[] = {
        {"key", FIELD_KEY}, {"value", FIELD_VALUE}, {"ignoreCase", FIELD_IGNORECASE}, {"id", FIELD_ID}, {"aux", FIELD_AUX}};
    if (!e.key) return FIELD_VALUE;
    for (const auto &t : table)
        if (e.key_len == strlen(t.name) && strncasecmp(e.key, t.name, e.key_len) == 0) return t.f;
    return FIELD_NONE;
}

// one element of the pattern array -> record; 0 on success
int process_pattern(long pidx, PatternRec &p, const aho_array_t *sub, aho_diag_t *diag)
{
    bool seen[6] = {false, false, false, false, false, false};
    int rc = 0;
    for (size_t k = 0; k < sub->n && rc == 0; ++k) {
        const aho_entry_t &e = sub->entries[k];
        const Field f = field_of(e);
        seen[f] = true;
        switch (f) {
        case FIELD_NONE:
            warn(diag, "Invalid structure (unrecognized sub-array key)! Only allowed are: {key, id, value, aux, "
                       "ignoreCase}. Cannot initialize. Pattern index: %ld", pidx);
            rc = -2;
            break;
        case FIELD_ID:
            if (e.val.type != AHO_T_LONG) {
                throw_aho(diag, "Invalid type of pattern ID given (long required), type: %s, pattern index: %ld",
                          type_str(e.val.type), pidx);
                rc = -5;
                break;
            }
            p.key_id = e.val.lval;
            p.key_type = AC_PATTID_TYPE_NUMBER;
            break;
        case FIELD_AUX:
            p.has_aux = true;
            p.aux_opaque = e.val.opaque;
            break;
        case FIELD_KEY:
        case FIELD_VALUE: {
            if (e.val.type != AHO_T_STRING) {
                throw_aho(diag, "Pattern %s has to be a string, type: %s, pattern index: %ld",
                          f == FIELD_KEY ? "key" : "value", type_str(e.val.type), pidx);
                rc = -5;
                break;
            }
            std::string &dst = (f == FIELD_KEY) ? p.key : p.value;
            dst.assign(e.val.sval ? e.val.sval : "", e.val.slen);
            if (f == FIELD_KEY) { p.key_opaque = e.val.opaque; p.has_key = true; p.key_type = AC_PATTID_TYPE_STRING; }
            else { p.value_opaque = e.val.opaque; p.has_value = true; }
            break;
        }
        case FIELD_IGNORECASE:
            break;
        }
    }
    if (rc == 0 && !p.has_value) {
        warn(diag, "No value was specified for pattern index: %ld", pidx);
        rc = -2;
    } else if (rc == 0 && seen[FIELD_KEY] && seen[FIELD_ID]) {
        warn(diag, "Pattern can have either numeric or string identifier, not both! Pattern index: %ld", pidx);
        rc = -3;
    }
    if (seen[FIELD_IGNORECASE])
        warn(diag, "ignoreCase attribute is deprecated and is ignored. Pattern index: %ld", pidx);
    return rc;
}

} // namespace

struct aho_master {
    AC_TRIE_t *acap = nullptr;
    bool ac_finalized = false;
    bool init_ok = false;
    bool closed = false;                       // zend_list_close() happened
    std::vector<PatternRec *> patterns;        // all accepted calls' records (owned)
};

namespace {

void release_master(aho_master *m)
{
    if (m->acap) { ac_trie_release(m->acap); m->acap = nullptr; }
    for (PatternRec *p : m->patterns) delete p;
    m->patterns.clear();
}

// php_ahocorasick_process_patterns, src/php_ahocorasick.c:389-489
int process_patterns(aho_master *m, const aho_array_t *data, aho_diag_t *diag)
{
    std::vector<PatternRec *> list;            // array order
    int status = 0;
    for (size_t k = 0; k < data->n; ++k) {
        const aho_entry_t &e = data->entries[k];
        if (e.val.type != AHO_T_ARRAY || !e.val.aval) {
            warn(diag, "Invalid pattern structure! Cannot initialize.");
            status = -4;
            break;
        }
        PatternRec *p = new PatternRec();
        list.push_back(p);
        const long pidx = e.key ? (long)k : e.index;
        if (process_pattern(pidx, *p, e.val.aval, diag) != 0) { status = -1; break; }
    }
    if (status != 0) {
        for (PatternRec *p : list) delete p;
        return status;
    }
    // The reference links every element at the list head and then walks the list from the head,
    // so the trie receives the patterns of one call LAST ELEMENT FIRST; the add status is ignored.
    for (size_t k = list.size(); k-- > 0;) {
        PatternRec *p = list[k];
        AC_PATTERN_t patt;
        memset(&patt, 0, sizeof(patt));
        patt.ptext.astring = p->value.data();
        patt.ptext.length = p->value.size();
        patt.rtext.astring = nullptr;
        patt.rtext.length = 0;
        patt.id.type = (enum ac_pattid_type)p->key_type;
        if (p->key_type == AC_PATTID_TYPE_NUMBER) patt.id.u.number = p->key_id;
        else if (p->key_type == AC_PATTID_TYPE_STRING) patt.id.u.stringy = p->key.c_str();
        patt.aux = p;
        ac_trie_add(m->acap, &patt, 1);
    }
    m->patterns.insert(m->patterns.end(), list.begin(), list.end());
    return 0;
}

// php_ahocorasick_finalize, src/php_ahocorasick.c:130-142
int finalize_once(aho_master *m)
{
    if (!m || !m->init_ok || m->ac_finalized) return 0;
    m->ac_finalized = true;
    ac_trie_finalize(m->acap);
    return 1;
}

// ahocorasick_match() on one haystack between these sizes fills its records from device-expanded hit columns
constexpr size_t HITS_MIN_BYTES = 32u << 10, HITS_MAX_BYTES = 32u << 20;

// php_ahocorasick_match_handler, src/php_ahocorasick.c:542-589
void append_hits(std::vector<aho_hit_t> &out, const AC_MATCH_t *mt)
{
    for (size_t j = 0; j < mt->size; ++j) {
        const PatternRec *p = static_cast<const PatternRec *>(mt->patterns[j].aux);
        if (!p) continue;
        aho_hit_t h;
        memset(&h, 0, sizeof(h));
        h.pos = (long)mt->position;
        if (mt->patterns[j].id.type == AC_PATTID_TYPE_STRING) { h.key_type = 2; h.key_opaque = p->key_opaque; }
        else if (mt->patterns[j].id.type == AC_PATTID_TYPE_NUMBER) { h.key_type = 1; h.key_idx = mt->patterns[j].id.u.number; }
        h.has_aux = p->has_aux ? 1 : 0;
        h.aux_opaque = p->aux_opaque;
        h.start_postion = (long)mt->position - (long)p->value.size();
        h.value_opaque = p->value_opaque;
        h.value = p->value.data();
        h.value_len = p->value.size();
        out.push_back(h);
    }
}

int batch_cb(size_t idx, AC_MATCH_t *mt, void *user)
{
    auto *per = static_cast<std::vector<std::vector<aho_hit_t>> *>(user);
    append_hits((*per)[idx], mt);
    return 0;
}

aho_result_t *make_result(const std::vector<aho_hit_t> &hits, bool is_false)
{
    aho_result_t *r = new aho_result_t();
    r->is_false = is_false ? 1 : 0;
    r->n = hits.size();
    r->hits = hits.empty() ? nullptr : new aho_hit_t[hits.size()];
    for (size_t i = 0; i < hits.size(); ++i) r->hits[i] = hits[i];
    return r;
}

} // namespace

extern "C" {

aho_master_t *ahocorasick_init(const aho_array_t *data, aho_diag_t *diag)
{
    if (diag) { diag->n_warnings = 0; diag->exception[0] = 0; }
    aho_master *m = new aho_master();
    m->acap = ac_trie_create();
    if (process_patterns(m, data, diag) != 0) {     // :819-824
        release_master(m);
        delete m;
        return nullptr;
    }
    m->init_ok = true;
    return m;
}

int ahocorasick_add_patterns(aho_master_t *m, const aho_array_t *data, aho_diag_t *diag)
{
    if (diag) { diag->n_warnings = 0; diag->exception[0] = 0; }
    if (!m || m->closed || !m->init_ok) {
        warn(diag, "Cannot add a new pattern, not initialized");
        return 0;
    }
    if (m->ac_finalized) {
        warn(diag, "Cannot add a new pattern to finalized search structure");
        return 0;
    }
    return process_patterns(m, data, diag) == 0 ? 1 : 0;
}

int ahocorasick_finalize(aho_master_t *m, aho_diag_t *diag)
{
    if (diag) { diag->n_warnings = 0; diag->exception[0] = 0; }
    if (!m || m->closed) return 0;
    return finalize_once(m);
}

int ahocorasick_match_batch(const char *const *haystacks, const size_t *lens, size_t n, aho_master_t *m,
                            int find_all, aho_result_t **results, aho_diag_t *diag)
{
    if (diag) { diag->n_warnings = 0; diag->exception[0] = 0; }
    if (!m || m->closed) { warn(diag, "Invalid resource."); return -1; }       // :696-699
    if (!m->init_ok) { warn(diag, "Not initialized."); return -1; }            // :701-704
    finalize_once(m);                                                          // :707
    std::vector<AC_TEXT_t> texts(n);
    for (size_t i = 0; i < n; ++i) { texts[i].astring = haystacks[i]; texts[i].length = lens[i]; }
    std::vector<std::vector<aho_hit_t>> per(n);
    // findAll=false: the reference's callback returns 1 after the first event (:588)
    int rc;
    if (n == 1 && find_all && lens[0] > HITS_MIN_BYTES && lens[0] <= HITS_MAX_BYTES) {
        // One haystack that is neither tiny (one-launch path) nor huge (slab pipeline): the device expands the events
        // into {pos, start_postion, pattern} columns and the records are filled from those — no callback per event
        // (what php_ahocorasick_match_handler does per reported pattern, src/php_ahocorasick.c:555-584).
        const uint64_t offs[2] = {0, (uint64_t)lens[0]};
        std::vector<ACB200_HIT_t> hits(4096);
        size_t total = 0;
        rc = acb200_search_hits(m->acap, haystacks[0], offs, 1, hits.data(), hits.size(), &total);
        if (rc == 0 && total > hits.size()) {
            hits.resize(total);
            rc = acb200_last_hits(m->acap, hits.data(), hits.size(), &total);
        }
        if (rc == 0) {
            per[0].reserve(total);
            for (size_t i = 0; i < total; ++i) {
                const AC_PATTERN_t *pat = acb200_pattern(m->acap, hits[i].pattern);
                const PatternRec *p = pat ? static_cast<const PatternRec *>(pat->aux) : nullptr;
                if (!p) continue;
                aho_hit_t h;
                memset(&h, 0, sizeof(h));
                h.pos = (long)hits[i].end;
                if (pat->id.type == AC_PATTID_TYPE_STRING) { h.key_type = 2; h.key_opaque = p->key_opaque; }
                else if (pat->id.type == AC_PATTID_TYPE_NUMBER) { h.key_type = 1; h.key_idx = pat->id.u.number; }
                h.has_aux = p->has_aux ? 1 : 0;
                h.aux_opaque = p->aux_opaque;
                h.start_postion = (long)hits[i].start;
                h.value_opaque = p->value_opaque;
                h.value = p->value.data();
                h.value_len = p->value.size();
                per[0].push_back(h);
            }
        }
    } else if (n == 1) {
        const uint64_t offs[2] = {0, (uint64_t)lens[0]};     // single haystack: no gather copy
        rc = ac_trie_search_flat(m->acap, haystacks[0], offs, 1, find_all ? 0 : 1, batch_cb, &per);
    } else {
        rc = ac_trie_search_batch(m->acap, texts.data(), n, find_all ? 0 : 1, batch_cb, &per);
    }
    if (rc != 0) {
        warn(diag, "GPU search failed: %s", acb200_last_error());
        return -1;
    }
    for (size_t i = 0; i < n; ++i) results[i] = make_result(per[i], false);
    return 0;
}

aho_result_t *ahocorasick_match(const char *haystack, size_t len, aho_master_t *m, int find_all, aho_diag_t *diag)
{
    aho_result_t *r = nullptr;
    const char *hs[1] = {haystack};
    const size_t ls[1] = {len};
    if (ahocorasick_match_batch(hs, ls, 1, m, find_all, &r, diag) != 0)
        return make_result(std::vector<aho_hit_t>(), true);
    return r;
}

int ahocorasick_isValid(const aho_master_t *m)
{
    return (m && !m->closed && m->init_ok) ? 1 : 0;
}

int ahocorasick_deinit(aho_master_t *m, aho_diag_t *diag)
{
    if (diag) { diag->n_warnings = 0; diag->exception[0] = 0; }
    if (!m || m->closed) return 0;
    // The reference finalizes here too (:782); building a device table only to free it would be
    // pointless, so an automaton that was never finalized is simply marked closed.
    m->ac_finalized = true;
    m->init_ok = false;
    release_master(m);                        // resource destructor, :494-512
    m->closed = true;
    return 1;
}

void aho_resource_free(aho_master_t *m)
{
    if (!m) return;
    if (!m->closed) release_master(m);
    delete m;
}

void aho_result_free(aho_result_t *r)
{
    if (!r) return;
    delete[] r->hits;
    delete r;
}

void *aho_master_trie(aho_master_t *m) { return m ? m->acap : nullptr; }

} // extern "C"
