/*
 * php.h — DECLARATION-ONLY stand-in for the Zend / PHP 7-8 headers, just wide enough to type-check
 * php/php_ahocorasick_b200.c in an image without php-dev (tests/test_php_extension_source.py runs
 * `gcc -fsyntax-only` and a link-free compile against it).  Nothing here is an implementation and nothing here
 * is used when the extension is built for real (phpize puts the real headers first on the include path).
 * Names, argument orders and macro shapes follow the public Zend API of PHP 7.0 - 8.3.
 */
#ifndef ACB200_ZEND_STUB_PHP_H
#define ACB200_ZEND_STUB_PHP_H

#include <stddef.h>
#include <stdint.h>
#include <string.h>

typedef int64_t zend_long;
typedef uint64_t zend_ulong;
typedef unsigned char zend_bool;
typedef unsigned char zend_uchar;

#define SUCCESS 0
#define FAILURE (-1)
#define E_WARNING 2

enum { IS_UNDEF = 0, IS_NULL, IS_FALSE, IS_TRUE, IS_LONG, IS_DOUBLE, IS_STRING, IS_ARRAY, IS_OBJECT, IS_RESOURCE, IS_REFERENCE };

typedef struct _zend_string { uint32_t refcount; size_t len; char val[1]; } zend_string;
typedef struct _zend_array HashTable;
typedef struct _zend_resource { uint32_t refcount; int handle; int type; void *ptr; } zend_resource;
typedef struct _zend_class_entry { const char *name; } zend_class_entry;
typedef struct _zval_struct zval;
struct _zval_struct {
    union { zend_long lval; double dval; zend_string *str; HashTable *arr; zend_resource *res; zval *ref; void *ptr; } value;
    union { uint32_t type_info; struct { zend_uchar type; } v; } u1;
    uint32_t u2;
};
typedef struct _Bucket { zval val; zend_ulong h; zend_string *key; } Bucket;
struct _zend_array { uint32_t refcount; uint32_t nNumUsed, nNumOfElements; Bucket *arData; };

#define ZSTR_VAL(s) ((s)->val)
#define ZSTR_LEN(s) ((s)->len)
#define Z_TYPE(z) ((z).u1.v.type)
#define Z_TYPE_P(zp) Z_TYPE(*(zp))
#define Z_LVAL_P(zp) ((zp)->value.lval)
#define Z_STR_P(zp) ((zp)->value.str)
#define Z_ARRVAL_P(zp) ((zp)->value.arr)
#define Z_RES_P(zp) ((zp)->value.res)
#define Z_RES_TYPE_P(zp) ((zp)->value.res->type)
#define ZVAL_UNDEF(zp) do { Z_TYPE_P(zp) = IS_UNDEF; } while (0)
#define ZVAL_FALSE(zp) do { Z_TYPE_P(zp) = IS_FALSE; } while (0)
#define ZVAL_DEREF(zp) do { if (Z_TYPE_P(zp) == IS_REFERENCE) (zp) = (zp)->value.ref; } while (0)
#define ZVAL_COPY(dst, src) do { *(dst) = *(src); zend_stub_addref(dst); } while (0)

void zend_stub_addref(zval *z);
void zval_ptr_dtor(zval *z);
zend_string *zend_string_copy(zend_string *s);
void zend_string_release(zend_string *s);
zend_string *zval_get_string(zval *z);
const char *zend_zval_type_name(const zval *z);
int zend_binary_strcasecmp(const char *s1, size_t len1, const char *s2, size_t len2);

void *emalloc(size_t n);
void *ecalloc(size_t n, size_t size);
void *erealloc(void *p, size_t n);
void efree(void *p);

uint32_t zend_hash_num_elements(const HashTable *ht);
zval *zend_hash_update(HashTable *ht, zend_string *key, zval *v);
zval *zend_hash_index_update(HashTable *ht, zend_ulong h, zval *v);

/* iteration macros: same variable contract as the real ones (key may be NULL, idx is the integer key) */
#define ZEND_HASH_FOREACH_BODY_(ht) { Bucket *_p = (ht)->arData, *_end = _p + (ht)->nNumUsed; for (; _p != _end; _p++) { zval *_z = &_p->val; if (Z_TYPE_P(_z) == IS_UNDEF) continue;
#define ZEND_HASH_FOREACH_VAL(ht, _val) ZEND_HASH_FOREACH_BODY_(ht) _val = _z;
#define ZEND_HASH_FOREACH_KEY(ht, _h, _key) ZEND_HASH_FOREACH_BODY_(ht) _h = _p->h; _key = _p->key; (void)_z;
#define ZEND_HASH_FOREACH_STR_KEY_VAL(ht, _key, _val) ZEND_HASH_FOREACH_BODY_(ht) _key = _p->key; _val = _z;
#define ZEND_HASH_FOREACH_KEY_VAL(ht, _h, _key, _val) ZEND_HASH_FOREACH_BODY_(ht) _h = _p->h; _key = _p->key; _val = _z;
#define ZEND_HASH_FOREACH_END() } }

void array_init(zval *arr);
void array_init_size(zval *arr, uint32_t size);
void add_assoc_long(zval *arr, const char *key, zend_long n);
void add_assoc_str(zval *arr, const char *key, zend_string *str);
void add_assoc_zval(zval *arr, const char *key, zval *value);
void add_next_index_zval(zval *arr, zval *value);

void php_error_docref(const char *docref, int type, const char *format, ...) __attribute__((format(printf, 3, 4)));

typedef void (*rsrc_dtor_func_t)(zend_resource *res);
int zend_register_list_destructors_ex(rsrc_dtor_func_t ld, rsrc_dtor_func_t pld, const char *type_name, int module_number);
zend_resource *zend_register_resource(void *rsrc_pointer, int rsrc_type);
void *zend_fetch_resource(zend_resource *res, const char *resource_type_name, int resource_type);
int zend_list_close(zend_resource *res);

/* functions, argument info, module entry */
typedef struct _zend_execute_data zend_execute_data;
#define INTERNAL_FUNCTION_PARAMETERS zend_execute_data *execute_data, zval *return_value
typedef void (*zif_handler)(INTERNAL_FUNCTION_PARAMETERS);
#define PHP_FUNCTION(name) void zif_##name(INTERNAL_FUNCTION_PARAMETERS)
#define ZEND_NUM_ARGS() zend_stub_num_args(execute_data)
uint32_t zend_stub_num_args(zend_execute_data *ex);
int zend_parse_parameters(uint32_t num_args, const char *type_spec, ...);

#define RETURN_FALSE do { ZVAL_FALSE(return_value); return; } while (0)
#define RETURN_TRUE do { Z_TYPE_P(return_value) = IS_TRUE; return; } while (0)
#define RETURN_BOOL(b) do { Z_TYPE_P(return_value) = (b) ? IS_TRUE : IS_FALSE; return; } while (0)
#define RETURN_RES(r) do { return_value->value.res = (r); Z_TYPE_P(return_value) = IS_RESOURCE; return; } while (0)

typedef struct _zend_internal_arg_info { const char *name; int type; int pass_by_reference; } zend_internal_arg_info;
#define ZEND_BEGIN_ARG_INFO_EX(name, unused, return_reference, required_num_args) static const zend_internal_arg_info name[] = { {(const char *)(uintptr_t)(required_num_args), 0, return_reference},
#define ZEND_ARG_INFO(pass_by_ref, name) {#name, 0, pass_by_ref},
#define ZEND_ARG_ARRAY_INFO(pass_by_ref, name, allow_null) {#name, IS_ARRAY, pass_by_ref},
#define ZEND_END_ARG_INFO() };

typedef struct _zend_function_entry { const char *fname; zif_handler handler; const zend_internal_arg_info *arg_info; uint32_t num_args; uint32_t flags; } zend_function_entry;
#define PHP_FE(name, arg_info) {#name, zif_##name, arg_info, (uint32_t)(sizeof(arg_info) / sizeof(arg_info[0]) - 1), 0},
#define PHP_FE_END {NULL, NULL, NULL, 0, 0}

#define INIT_FUNC_ARGS int type, int module_number
#define ZEND_MODULE_INFO_FUNC_ARGS struct _zend_module_entry *zend_module
#define PHP_MINIT_FUNCTION(module) int zm_startup_##module(INIT_FUNC_ARGS)
#define PHP_MINFO_FUNCTION(module) void zm_info_##module(ZEND_MODULE_INFO_FUNC_ARGS)
#define PHP_MINIT(module) zm_startup_##module
#define PHP_MINFO(module) zm_info_##module
typedef struct _zend_module_entry {
    unsigned short size; unsigned int zend_api; unsigned char zend_debug, zts; const void *ini_entry; const void *deps;
    const char *name; const zend_function_entry *functions;
    int (*module_startup_func)(INIT_FUNC_ARGS); int (*module_shutdown_func)(INIT_FUNC_ARGS);
    int (*request_startup_func)(INIT_FUNC_ARGS); int (*request_shutdown_func)(INIT_FUNC_ARGS);
    void (*info_func)(ZEND_MODULE_INFO_FUNC_ARGS); const char *version;
    size_t globals_size; void *globals_ptr; void (*globals_ctor)(void *); void (*globals_dtor)(void *); int (*post_deactivate_func)(void);
    int module_started; unsigned char type; void *handle; int module_number; const char *build_id;
} zend_module_entry;
#define STANDARD_MODULE_HEADER sizeof(zend_module_entry), 20200930, 0, 0, NULL, NULL
#define STANDARD_MODULE_PROPERTIES 0, NULL, NULL, NULL, NULL, 0, 0, NULL, 0, "stub"
#define ZEND_GET_MODULE(name) zend_module_entry *get_module(void) { return &name##_module_entry; }

zend_class_entry *zend_register_internal_class_ex(zend_class_entry *class_entry, zend_class_entry *parent_ce);
#define INIT_CLASS_ENTRY(ce, class_name, functions) do { (ce).name = (class_name); } while (0)

#endif
