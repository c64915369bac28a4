#!/bin/bash
# Build tools/preview_latency.c against libmdzcuda and against the reference's own pool
# (oracle/_ref/libmdzref.so) and run both.  The reference's pool has a known intermittent
# hang (reference BUGS:1-6), so its runs are short and guarded by `timeout`.
set -u
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
OUT="${1:-/tmp}"
L="-l:libmpfr.so.6 -l:libgmp.so.10 -lpthread -lm"
gcc -std=gnu99 -O1 -o "$OUT/preview_cuda" "$ROOT/tools/preview_latency.c" -L"$ROOT/mdz_b200" -lmdzcuda -Wl,-rpath,"$ROOT/mdz_b200" $L || exit 1
gcc -std=gnu99 -O1 -o "$OUT/preview_ref" "$ROOT/tools/preview_latency.c" -L"$ROOT/oracle/_ref" -lmdzref -Wl,-rpath,"$ROOT/oracle/_ref" $L || exit 1
for cfg in "300 64" "300 128" "3000 64"; do
  echo "== libmdzcuda (B200), depth/precision $cfg"
  "$OUT/preview_cuda" 200 $cfg
  echo "== reference pool ($(nproc) host threads), depth/precision $cfg"
  ok=0
  for try in 1 2 3 4 5 6; do
    if timeout 10 "$OUT/preview_ref" 20 $cfg 1; then ok=1; break; else echo "   (reference run $try hung: killed after 10 s)"; fi
  done
  timeout 20 "$OUT/preview_ref" 200 $cfg 2 || echo "   (reference interrupt run hung: killed after 20 s)"
done
