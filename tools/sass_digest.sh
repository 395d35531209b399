#!/bin/bash
# SASS evidence that the production kernels are Blackwell-native: per-kernel counts of tcgen05.mma (UTC*MMA), tcgen05.ld/st
# (LDTM / STTM), TMA (UTMALDG), mma.sync (HMMA, the legacy tensor path) and peer-memory instructions, from the in-tree .so.
#   tools/sass_digest.sh > profiles/r02_sass_digest.txt
SO=${1:-hi3d_official_b200/libhi3d_b200.so}
echo "# cuobjdump -sass $SO  ($(date -u +%F)); counts per kernel (mangled names abbreviated)"
printf "%-64s %8s %8s %6s %6s %8s %6s %6s\n" kernel UTCHMMA UTC2CTA LDTM STTM UTMALDG HMMA MUFU
cuobjdump -sass "$SO" | awk '
  /Function :/ { if (f != "") out(); f=$3; u=0; u2=0; l=0; s=0; t=0; h=0; m=0 }
  /UTCHMMA/ { u++; if ($0 ~ /2CTA/) u2++ }
  /LDTM/ { l++ } /STTM/ { s++ } /UTMALDG/ { t++ } /HMMA/ { if ($0 !~ /UTC/) h++ } /MUFU/ { m++ }
  function out() { n=f; gsub(/^_ZN4hi3d[0-9]*/, "", n); printf "%-64s %8d %8d %6d %6d %8d %6d %6d\n", substr(n,1,64), u, u2, l, s, t, h, m }
  END { out() }' | sort
