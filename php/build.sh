#!/bin/sh
# Builds the PHP extension where a PHP toolchain exists and runs the reference's six .phpt files UNCHANGED against it.
#   php/build.sh [reference checkout with tests/*.phpt]
# Without php-config (this image) it says so and exits 0: the extension source is still type-checked against the
# declaration-only headers in php/zend_stub/ by tests/test_php_extension_source.py.
set -e
here=$(cd "$(dirname "$0")" && pwd)
root=$(dirname "$here")
ref=${1:-/root/reference}
if ! command -v php-config >/dev/null 2>&1 || ! command -v phpize >/dev/null 2>&1; then
    echo "php-config / phpize not found: skipping the extension build (nothing to do in this image)"
    exit 0
fi
make -C "$root/php_aho_corasick_b200/csrc"
cd "$here"
phpize
./configure --with-acb200="$root"
make
# the reference's own tests, byte for byte; they SKIP themselves when the extension is not loaded
if [ -d "$ref/tests" ]; then
    mkdir -p tests
    cp "$ref"/tests/*.phpt tests/
    LD_LIBRARY_PATH="$root/php_aho_corasick_b200:$LD_LIBRARY_PATH" NO_INTERACTION=1 REPORT_EXIT_STATUS=1 make test
else
    echo "no reference tests under $ref/tests: built only"
fi
