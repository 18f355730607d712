#!/bin/bash
# AddressSanitizer + UndefinedBehaviorSanitizer over the host-side code ("-fsanitize=address,undefined on the host oracle harness"):
#   * the oracle restatement (oracle/x266_oracle.c) under the whole CPU test suite that drives it,
#   * the product's packed-SATD arithmetic header (satd_packed.h) inside its CPU model test, and the host copy pool of the pageable
#     path (x266_b200/csrc/hostcopy.cpp) as a stand-alone binary.
# Writes profiles/r02_asan_ubsan.log.  No GPU needed.
set -u
cd "$(dirname "$0")/.."
LOG=profiles/r02_asan_ubsan.log
TMP=$(mktemp -d)
SAN="-fsanitize=address,undefined -fno-sanitize-recover=undefined -fno-omit-frame-pointer -g -O1"
{
  echo "# ASan + UBSan run, $(gcc --version | head -1), $(date -u +%Y-%m-%dT%H:%MZ)"
  echo "## oracle/x266_oracle.c under tests/test_oracle.py (liboracle built with: $SAN)"
  gcc $SAN -fPIC -shared -Wall -Wextra -Wno-maybe-uninitialized oracle/x266_oracle.c -o $TMP/liboracle_san.so -lpthread
  X266_TEST_CXXFLAGS="$SAN" X266_ORACLE_LIB=$TMP/liboracle_san.so LD_PRELOAD="$(gcc -print-file-name=libasan.so) $(gcc -print-file-name=libubsan.so)" \
    ASAN_OPTIONS=detect_leaks=0:abort_on_error=0 UBSAN_OPTIONS=print_stacktrace=1 \
    python -m pytest tests/test_oracle.py -q -x -p no:cacheprovider 2>&1 | tail -6
  echo "## x266_b200/csrc/hostcopy.cpp + tests/c/hostcopy_test.cpp"
  g++ $SAN -std=c++17 -I x266_b200/csrc tests/c/hostcopy_test.cpp x266_b200/csrc/hostcopy.cpp -o $TMP/hostcopy_san -lpthread && $TMP/hostcopy_san 2>&1 | tail -12
  echo "## (x266_b200/csrc/satd_packed.h is compiled with the same flags inside test_packed_search_arithmetic_model above)"
  echo "## done (any sanitizer report would appear above as 'ERROR: AddressSanitizer' or 'runtime error:')"
} > $LOG 2>&1
rm -rf $TMP
grep -c "ERROR: AddressSanitizer\|runtime error:" $LOG; tail -30 $LOG
