dnl config.m4 — phpize build of the "ahocorasick" extension on top of libacb200.so.
dnl
dnl   cd php && phpize && ./configure --with-acb200=/path/to/this/repo && make && make test
dnl
dnl The reference's config.m4 (config.m4:4-9) compiles its bundled MultiFast sources into the extension; this one
dnl compiles the single glue file and links the CUDA library, which exports the same five ac_trie_* symbols.
PHP_ARG_WITH(acb200, for the B200 Aho-Corasick matcher,
[  --with-acb200[=DIR]       Enable ahocorasick support; DIR = checkout of this repository (default: ..)])

if test "$PHP_ACB200" != "no"; then
  ACB200_DIR="$PHP_ACB200"
  if test "$ACB200_DIR" = "yes"; then
    ACB200_DIR=".."
  fi
  if test ! -f "$ACB200_DIR/include/acb200.h"; then
    AC_MSG_ERROR([include/acb200.h not found under $ACB200_DIR])
  fi
  if test ! -f "$ACB200_DIR/php_aho_corasick_b200/libacb200.so"; then
    AC_MSG_ERROR([libacb200.so is not built: run make -C $ACB200_DIR/php_aho_corasick_b200/csrc first])
  fi
  PHP_ADD_INCLUDE($ACB200_DIR/include)
  PHP_ADD_LIBRARY_WITH_PATH(acb200, $ACB200_DIR/php_aho_corasick_b200, AHOCORASICK_SHARED_LIBADD)
  PHP_SUBST(AHOCORASICK_SHARED_LIBADD)
  AC_DEFINE(HAVE_AHOCORASICK, 1, [Whether you have Aho Corasick])
  PHP_NEW_EXTENSION(ahocorasick, php_ahocorasick_b200.c, $ext_shared)
fi
