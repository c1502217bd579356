#!/bin/bash
# every UtNet layer shape (cs 248) with forced N_TILE / CTA-group variants; args: taps cin n B Hs Ws 0 0 epi n_tile ws ctas cg
cp tools/probe_epi tools/probe 2>/dev/null
while read -r name cfg; do
  for nt in 64 128 256; do
    for cg in 1 2; do
      out=$(timeout 60 ./tools/probe conv $cfg $nt -1 0 $cg 2>&1 | grep -E 'TFLOP|FAIL|failed' | tr '\n' ' ' | cut -c1-60)
      [ -n "$out" ] && echo "$name n_tile $nt cg $cg: $out"
    done
  done
done <<'CFG'
convs2.0 9 64 128 32 124 124 0 0 0
convs2.2 9 128 128 32 122 122 0 0 0
convs3.0 9 128 256 48 60 60 0 0 0
convs3.2 9 256 256 48 58 58 0 0 0
convs4.0 9 256 512 64 28 28 0 0 0
convs4.2 9 512 512 64 26 26 0 0 0
bottom.0 9 512 1024 96 12 12 0 0 0
bottom.2 9 1024 1024 96 14 14 0 0 0
tconvs1.0 9 1024 512 64 28 28 0 0 0
tconvs1.2 9 512 512 64 30 30 0 0 0
tconvs2.0 9 512 256 48 60 60 0 0 0
tconvs3.0 9 256 128 32 124 124 0 0 0
tconvs4.0 9 128 64 8 252 252 0 0 0
CFG
