#!/bin/bash
export NIND_TRACE=1
for cfg in "9 64 64 1 510 510 0 0 0" "1 64 64 1 506 506 0 0 0" "9 128 64 1 508 508 0 0 0" "9 128 128 2 252 252 0 0 0" "1 128 256 1 252 252 0 0 1" "9 512 256 4 124 124 0 0 0"; do
  timeout 60 ./tools/probe conv $cfg
done
