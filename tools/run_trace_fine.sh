#!/bin/bash
# per-tile clock trace with the fine-grained epilogue stamps (tools/probe_fine = probe built with -DNIND_TRACE_FINE=1)
export NIND_TRACE=1
for cfg in "9 8 64 32 252 252 0 0 0" "1 128 256 32 126 126 0 0 1" "9 64 64 32 250 250 0 0 0 0 -1 0 0 -1 1 1" "9 128 64 32 252 252 0 0 0" "9 64 128 32 126 126 0 0 0"; do
  echo "--- conv $cfg"
  timeout 60 ./tools/probe_fine conv $cfg | grep -v "^  mismatch"
done
