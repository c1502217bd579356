#!/bin/bash
for swap in 0 1; do
  for cfg in "0 8 16" "0 1 10" "3 1 10" "12 8 10" "21 1 10" "5 20 10"; do
    timeout 30 ./tools/probe desc0 $cfg $swap
  done
done
