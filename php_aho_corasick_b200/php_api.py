"""The reference's PHP userland API, mirrored in Python over libacb200.so.

Same six functions, argument meaning, return values, warning texts and exception
as src/php_ahocorasick.stub.php:12-37 / src/php_ahocorasick.c:623-925, plus
ahocorasick_match_batch().  PHP arrays are Python lists (integer keys) or dicts
(string keys, insertion order); PHP strings are str (encoded as UTF-8) or bytes.
E_WARNINGs surface as Python warnings of category AhoWarning.

All matching runs in the CUDA library; there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import warnings

from . import native

T_NULL, T_FALSE, T_TRUE, T_LONG, T_DOUBLE, T_STRING, T_ARRAY, T_OBJECT, T_RESOURCE = range(9)


class AhoException(Exception):
    """class AhoException extends Exception (src/php_ahocorasick.c:601-605)"""


class AhoWarning(UserWarning):
    """PHP E_WARNING raised by the extension"""


class _Array(C.Structure):
    pass


class _Value(C.Structure):
    _fields_ = [("type", C.c_int), ("lval", C.c_long), ("dval", C.c_double), ("sval", C.c_char_p),
                ("slen", C.c_size_t), ("aval", C.POINTER(_Array)), ("opaque", C.c_void_p)]


class _Entry(C.Structure):
    _fields_ = [("key", C.c_char_p), ("key_len", C.c_size_t), ("index", C.c_long), ("val", _Value)]


_Array._fields_ = [("entries", C.POINTER(_Entry)), ("n", C.c_size_t)]


class _Diag(C.Structure):
    _fields_ = [("warnings", (C.c_char * 256) * 8), ("n_warnings", C.c_int), ("exception", C.c_char * 512)]


class _Hit(C.Structure):
    _fields_ = [("pos", C.c_long), ("key_type", C.c_int), ("key_idx", C.c_long), ("key_opaque", C.c_void_p),
                ("has_aux", C.c_int), ("aux_opaque", C.c_void_p), ("start_postion", C.c_long),
                ("value_opaque", C.c_void_p), ("value", C.c_void_p), ("value_len", C.c_size_t)]


class _Result(C.Structure):
    _fields_ = [("is_false", C.c_int), ("hits", C.POINTER(_Hit)), ("n", C.c_size_t)]


_bound = False


def _lib():
    global _bound
    L = native.lib()
    if not _bound:
        L.ahocorasick_init.argtypes = [C.POINTER(_Array), C.POINTER(_Diag)]
        L.ahocorasick_init.restype = C.c_void_p
        L.ahocorasick_add_patterns.argtypes = [C.c_void_p, C.POINTER(_Array), C.POINTER(_Diag)]
        L.ahocorasick_add_patterns.restype = C.c_int
        L.ahocorasick_finalize.argtypes = [C.c_void_p, C.POINTER(_Diag)]
        L.ahocorasick_finalize.restype = C.c_int
        L.ahocorasick_match.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_int, C.POINTER(_Diag)]
        L.ahocorasick_match.restype = C.POINTER(_Result)
        L.ahocorasick_match_batch.argtypes = [C.POINTER(C.c_char_p), C.POINTER(C.c_size_t), C.c_size_t, C.c_void_p,
                                              C.c_int, C.POINTER(C.POINTER(_Result)), C.POINTER(_Diag)]
        L.ahocorasick_match_batch.restype = C.c_int
        L.ahocorasick_isValid.argtypes = [C.c_void_p]
        L.ahocorasick_isValid.restype = C.c_int
        L.ahocorasick_deinit.argtypes = [C.c_void_p, C.POINTER(_Diag)]
        L.ahocorasick_deinit.restype = C.c_int
        L.aho_resource_free.argtypes = [C.c_void_p]
        L.aho_resource_free.restype = None
        L.aho_result_free.argtypes = [C.POINTER(_Result)]
        L.aho_result_free.restype = None
        L.aho_master_trie.argtypes = [C.c_void_p]
        L.aho_master_trie.restype = C.c_void_p
        _bound = True
    return L


class Resource:
    """resource(…) of type (AhoCorasick search)"""

    def __init__(self, handle, objects):
        self._h = handle
        self._objects = objects      # Python values referenced by opaque handles (key/value/aux)

    def trie(self):
        return _lib().aho_master_trie(self._h)

    def __del__(self):
        try:
            if self._h:
                _lib().aho_resource_free(self._h)
                self._h = None
        except Exception:
            pass


class _Marshal:
    """Builds aho_array_t trees; keeps every ctypes buffer and Python object alive."""

    def __init__(self, objects):
        self.keep = []
        self.objects = objects

    def _opaque(self, obj):
        self.objects.append(obj)
        return len(self.objects)          # 1-based so that NULL means "none"

    def value(self, v) -> _Value:
        out = _Value()
        out.opaque = self._opaque(v)
        if v is None:
            out.type = T_NULL
        elif v is True:
            out.type = T_TRUE
        elif v is False:
            out.type = T_FALSE
        elif isinstance(v, int):
            out.type = T_LONG
            out.lval = v
        elif isinstance(v, float):
            out.type = T_DOUBLE
            out.dval = v
        elif isinstance(v, (str, bytes, bytearray)):
            b = v.encode("utf-8") if isinstance(v, str) else bytes(v)
            buf = C.create_string_buffer(b, len(b))
            self.keep.append(buf)
            out.type = T_STRING
            out.sval = C.cast(buf, C.c_char_p)
            out.slen = len(b)
        elif isinstance(v, (list, tuple, dict)):
            out.type = T_ARRAY
            arr = self.array(v)
            out.aval = C.pointer(arr)
        else:
            out.type = T_OBJECT
        return out

    def array(self, data) -> _Array:
        items = list(data.items()) if isinstance(data, dict) else list(enumerate(data))
        ents = (_Entry * max(1, len(items)))()
        for i, (k, v) in enumerate(items):
            if isinstance(k, str):
                kb = k.encode("utf-8")
                buf = C.create_string_buffer(kb, len(kb))
                self.keep.append(buf)
                ents[i].key = C.cast(buf, C.c_char_p)
                ents[i].key_len = len(kb)
            else:
                ents[i].key = None
                ents[i].index = int(k)
            ents[i].val = self.value(v)
        arr = _Array()
        arr.entries = C.cast(ents, C.POINTER(_Entry))
        arr.n = len(items)
        self.keep.append(ents)
        self.keep.append(arr)
        return arr


def _report(diag: _Diag):
    for i in range(diag.n_warnings):
        warnings.warn(bytes(diag.warnings[i].value).decode("utf-8", "replace"), AhoWarning, stacklevel=3)
    if diag.exception:
        raise AhoException(diag.exception.decode("utf-8", "replace"))


def _hits(res_ptr, objects):
    r = res_ptr.contents
    if r.is_false:
        return False
    out = []
    for i in range(r.n):
        h = r.hits[i]
        d = {"pos": h.pos}
        if h.key_type == 2:
            d["key"] = objects[h.key_opaque - 1]
        elif h.key_type == 1:
            d["keyIdx"] = h.key_idx
        if h.has_aux:
            d["aux"] = objects[h.aux_opaque - 1]
        d["start_postion"] = h.start_postion          # sic, src/php_ahocorasick.c:578
        d["value"] = objects[h.value_opaque - 1]
        out.append(d)
    return out


def ahocorasick_init(data):
    """ahocorasick_init(array $data): resource|false"""
    if not isinstance(data, (list, tuple, dict)):
        raise TypeError("ahocorasick_init(): Argument #1 ($data) must be of type array")
    L = _lib()
    objects = []
    m = _Marshal(objects)
    arr = m.array(data)
    diag = _Diag()
    h = L.ahocorasick_init(C.byref(arr), C.byref(diag))
    _report(diag)
    if not h:
        return False
    return Resource(h, objects)


def ahocorasick_add_patterns(resource, data):
    """ahocorasick_add_patterns(resource $id, array $patterns): bool"""
    L = _lib()
    m = _Marshal(resource._objects)
    arr = m.array(data)
    diag = _Diag()
    ok = L.ahocorasick_add_patterns(resource._h, C.byref(arr), C.byref(diag))
    _report(diag)
    return bool(ok)


def ahocorasick_finalize(resource):
    """ahocorasick_finalize(resource $id): bool — true only the first time"""
    diag = _Diag()
    ok = _lib().ahocorasick_finalize(resource._h, C.byref(diag))
    _report(diag)
    return bool(ok)


def _bytes(h):
    return h.encode("utf-8") if isinstance(h, str) else bytes(h)


def ahocorasick_match(haystack, resource, findAll=True):
    """ahocorasick_match(string $needle, resource $id, bool $findAll = true): array|false"""
    L = _lib()
    b = _bytes(haystack)
    buf = C.create_string_buffer(b, len(b))
    diag = _Diag()
    res = L.ahocorasick_match(C.cast(buf, C.c_void_p), len(b), resource._h, 1 if findAll else 0, C.byref(diag))
    try:
        out = _hits(res, resource._objects)
    finally:
        L.aho_result_free(res)
    _report(diag)
    return out


def ahocorasick_match_batch(haystacks, resource, findAll=True):
    """ahocorasick_match_batch(array $haystacks, resource $id, bool $findAll = true): array|false
    One device launch for the whole array; element i is what ahocorasick_match($haystacks[i]) returns."""
    L = _lib()
    bs = [_bytes(h) for h in haystacks]
    n = len(bs)
    bufs = [C.create_string_buffer(b, len(b)) for b in bs]
    ptrs = (C.c_char_p * max(1, n))(*[C.cast(x, C.c_char_p) for x in bufs])
    lens = (C.c_size_t * max(1, n))(*[len(b) for b in bs])
    results = (C.POINTER(_Result) * max(1, n))()
    diag = _Diag()
    rc = L.ahocorasick_match_batch(ptrs, lens, n, resource._h, 1 if findAll else 0, results, C.byref(diag))
    out = False
    if rc == 0:
        out = []
        for i in range(n):
            out.append(_hits(results[i], resource._objects))
            L.aho_result_free(results[i])
    _report(diag)
    return out


def ahocorasick_isValid(resource):
    """ahocorasick_isValid(resource $id): bool"""
    if not isinstance(resource, Resource):
        return False
    return bool(_lib().ahocorasick_isValid(resource._h))


def ahocorasick_deinit(resource):
    """ahocorasick_deinit(resource $id): bool — false on an already closed resource"""
    if not isinstance(resource, Resource):
        return False
    diag = _Diag()
    ok = _lib().ahocorasick_deinit(resource._h, C.byref(diag))
    _report(diag)
    return bool(ok)
