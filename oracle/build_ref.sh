#!/usr/bin/env bash
# Builds oracle/_ref/ from the REFERENCE's own sources where they lie
# (/root/reference, read-only).  Sources are extracted to a scratch directory
# and compiled there; only binaries land in oracle/_ref/ (git-ignored).
#   libs3ref.so  : the reference's vendored, patched libbz2 1.0.6 + ref_harness.c
#   starch3_ref  : the reference starch3 binary (makefile flags, mk:2, mk:18)
# Usage: oracle/build_ref.sh [reference_dir]   (default /root/reference)
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${1:-/root/reference}"
OUT="$HERE/_ref"
if [ ! -f "$REF/third-party/bzip2-1.0.6.tar.gz" ]; then
  echo "build_ref: reference not present at $REF; keeping prebuilt $OUT" >&2
  exit 0
fi
mkdir -p "$OUT"
SCR="$(mktemp -d /tmp/s3ref.XXXXXX)"
trap 'rm -rf "$SCR"' EXIT
tar xzf "$REF/third-party/bzip2-1.0.6.tar.gz" -C "$SCR"
BZ="$SCR/bzip2-1.0.6"
CFLAGS="-O2 -fPIC -D_FILE_OFFSET_BITS=64 -w"
for f in blocksort huffman crctable randtable decompress bzlib; do
  gcc $CFLAGS -c "$BZ/$f.c" -o "$SCR/$f.o"
done
gcc $CFLAGS -fvisibility=default -I"$BZ" -c "$HERE/ref_harness.c" -o "$SCR/ref_harness.o"
gcc -shared -o "$OUT/libs3ref.so" "$SCR"/{ref_harness,blocksort,huffman,crctable,randtable,decompress,bzlib}.o
echo "build_ref: built $OUT/libs3ref.so"
# the reference's vendored jansson 2.9: pins the archive metadata text (jansson_harness.c), and the reference
# binary needs it to satisfy its #include / link line
if [ ! -f "$OUT/libs3jansson.so" ] || { [ "${S3_SKIP_REF_BINARY:-0}" != "1" ] && [ ! -x "$OUT/starch3_ref" ]; }; then
  tar xzf "$REF/third-party/jansson-2.9.tar.gz" -C "$SCR"
  ( cd "$SCR/jansson-2.9" && CFLAGS="-O2 -fPIC" ./configure --prefix="$SCR/jansson" --disable-shared >/dev/null 2>&1 \
      && make -j8 >/dev/null 2>&1 && make install >/dev/null 2>&1 )
  gcc -O2 -fPIC -shared -fvisibility=hidden -I"$SCR/jansson/include" "$HERE/jansson_harness.c" \
      "$SCR/jansson/lib/libjansson.a" -o "$OUT/libs3jansson.so"
  echo "build_ref: built $OUT/libs3jansson.so"
fi
if [ "${S3_SKIP_REF_BINARY:-0}" != "1" ] && [ ! -x "$OUT/starch3_ref" ]; then
  ( cd "$BZ" && make libbz2.a CC=gcc >/dev/null 2>&1 )
  g++ -std=c++11 -O3 -D_LARGEFILE64_SOURCE -D_FILE_OFFSET_BITS=64 -DDEBUG -w \
      -I"$REF/include" -I"$BZ" -I"$SCR/jansson/include" \
      "$REF/src/starch3.cpp" -o "$OUT/starch3_ref" \
      "$BZ/libbz2.a" -lpthread "$SCR/jansson/lib/libjansson.a"
  echo "build_ref: built $OUT/starch3_ref"
fi
