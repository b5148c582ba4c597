/* declaration-only stand-in, see php.h in this directory */
#ifndef ACB200_ZEND_STUB_EXCEPTIONS_H
#define ACB200_ZEND_STUB_EXCEPTIONS_H
#include "php.h"
extern zend_class_entry *zend_ce_exception;
void *zend_throw_exception_ex(zend_class_entry *exception_ce, zend_long code, const char *format, ...) __attribute__((format(printf, 3, 4)));
#endif
