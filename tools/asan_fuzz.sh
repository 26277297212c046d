#!/bin/sh
# Fuzz every front-end parser under AddressSanitizer + UBSan (host only, no GPU):
#   tools/asan_fuzz.sh [mutants per format, default 1500]
# Builds loos_b200/frontend into a sanitized library under /tmp and runs tests/host/fuzz_readers.py on it
# for each trajectory format and for the text parsers (PDB, selection, frame ranges, matrix text).
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
N=${1:-1500}
OUT=$(mktemp -d /tmp/loosb200_fuzz.XXXXXX)
F=$ROOT/loos_b200/frontend
g++ -O1 -g -std=c++17 -pthread -fPIC -shared -fsanitize=address,undefined -fno-omit-frame-pointer \
    -o "$OUT/libfe_asan.so" "$F/frontend_capi.cpp" "$F/frontend.cpp" "$F/hdf5_mini.cpp" -lz
for fmt in dcd nc xtc h5 h5v2 text; do
  mkdir -p "$OUT/$fmt"
  LOOSB200_FE_LIB="$OUT/libfe_asan.so" ASAN_OPTIONS=detect_leaks=0:abort_on_error=1 \
  LD_PRELOAD="$(gcc -print-file-name=libasan.so):$(gcc -print-file-name=libubsan.so)" \
    python "$ROOT/tests/host/fuzz_readers.py" "$fmt" 100000 "$N" "$OUT/$fmt" > "$OUT/$fmt.log" 2>&1 \
    && echo "$fmt: $(tail -1 "$OUT/$fmt.log")" || { echo "$fmt: FAILED, see $OUT/$fmt.log"; exit 1; }
  if grep -q "runtime error" "$OUT/$fmt.log"; then echo "$fmt: UBSan findings in $OUT/$fmt.log"; exit 1; fi
done
