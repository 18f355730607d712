#!/bin/bash
# oracle/ref_conv_slice.sh -- TEST INFRASTRUCTURE.  Compiles the UNMODIFIED xConvInputFmt / xConvOutput420 of the reference
# (src/x266.cpp:415-492) and its ref_block_t (x266.cpp:55-63) into oracle/_ref/libx266conv.so.
#
# x266.cpp as a whole does not build with g++ (MSVC-only _aligned_malloc / _stricmp, goto over initialisers), but these two
# functions are plain C: the recipe cuts them out of the source WHERE IT LIES by their first and last lines, pipes them -- between a
# four-line prologue (the reference's own REF_BLOCK_SZ / PACKED definitions, x266.cpp:36,42) and two extern "C" forwarders -- into
# the compiler's stdin.  No reference source is written anywhere; only the .so lands in the git-ignored oracle/_ref/.
set -e
REF=${1:-/root/reference}
OUT=$(dirname "$0")/_ref
SRC=$REF/src/x266.cpp
[ -f "$SRC" ] || { echo "reference tree absent: keeping prebuilt oracle/_ref/libx266conv.so (if any)"; exit 0; }
mkdir -p "$OUT"
{
  cat <<'HEAD'
#include <assert.h>
#include <stdint.h>
#include <string.h>
#define REF_BLOCK_SZ        (16)
#define PACKED( class_to_pack ) class_to_pack __attribute__((__packed__))
HEAD
  sed -n '/^PACKED(struct _ref_block_t/,/^typedef struct _ref_block_t ref_block_t;/p' "$SRC"
  sed -n '/^void xConvInputFmt(ref_block_t/,/^int xCodecInit(/p' "$SRC" | sed '$d'
  cat <<'TAIL'
extern "C" void ref_xConvInputFmt(void* blk, const uint8_t* y, const uint8_t* u, const uint8_t* v, intptr_t strdY, int w, int h)
{ xConvInputFmt((ref_block_t*)blk, y, u, v, strdY, w, h); }
extern "C" void ref_xConvOutput420(const void* blk, uint8_t* y, intptr_t strdY, uint8_t* u, uint8_t* v, intptr_t strdC, int w, int h)
{ xConvOutput420((const ref_block_t*)blk, y, strdY, u, v, strdC, w, h); }
extern "C" int ref_sizeof_ref_block_t(void) { return (int)sizeof(ref_block_t); }
TAIL
} | g++ -O2 -fPIC -shared -w -x c++ - -o "$OUT/libx266conv.so"
echo "built oracle/_ref/libx266conv.so from $SRC"
