#!/bin/bash
# SASS evidence: which Blackwell-native instructions each object of libpdr.so contains
# (B200_PROFILING.md "What proves a Blackwell-native kernel").  Usage: tools/sass_summary.sh > profiles/rNN_sass_summary.txt
cd "$(dirname "$0")/.."
echo "# cuobjdump -sass of build/pdr/*.o (nvcc 12.9, -gencode arch=compute_100a,code=sm_100a), mnemonic counts"
for o in build/pdr/*.o; do
  n=$(basename $o .o)
  cuobjdump -sass $o 2>/dev/null > /tmp/_sass_$$.txt
  echo "== $n.cu"
  grep -oE "\b(UTC[A-Z]*MMA(\.2CTA)?|UTMALDG(\.[0-9]D)?(\.2CTA)?|UTMASTG(\.[0-9]D)?|UBLKCP|LDTM(\.[0-9x]+)*|STTM|UTCBAR(\.2CTA)?(\.MULTICAST)?|UTCCP|HMMA\.[0-9]+|LDSM|LDGSTS|SYNCS\.[A-Z_.]+|MUFU\.[A-Z0-9]+|DFMA|ATOM[GS]?\.[A-Z.0-9_]+|RED\.[A-Z.0-9_]+)" /tmp/_sass_$$.txt \
    | sort | uniq -c | sort -rn | awk '{printf "   %6d  %s\n", $1, $2}' | head -24
  echo "   kernels:" $(grep -c "Function :" /tmp/_sass_$$.txt)
done
rm -f /tmp/_sass_$$.txt
