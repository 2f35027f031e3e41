#!/bin/bash
# SASS evidence of the Blackwell-native paths in the shipped library: counts of the mnemonics B200_PROFILING.md names
# (tcgen05.mma -> UTC*MMA, tcgen05.ld -> LDTM, TMA -> UTMALDG, tcgen05.commit -> UTCBAR, mbarrier -> SYNCS) per kernel.
#   tools/sass_summary.sh > profiles/r2_sass_summary.txt
cd "$(dirname "$0")/.."
LIB=outline_rag_b200/liborx.so
echo "# cuobjdump -sass $LIB  ($(date -u +%Y-%m-%dT%H:%MZ), nvcc $(nvcc --version | grep -o 'release [0-9.]*'), $(stat -c %s $LIB) bytes)"
echo "# whole library"
cuobjdump -sass $LIB > /tmp/orx_sass.txt
grep -oE "UTCHMMA[.A-Z0-9_]*|UTMALDG[.A-Z0-9_]*|UTCBAR[.A-Z0-9_]*|LDTM[.A-Z0-9_x]*|SYNCS[.A-Z0-9_]*|UBLKCP[.A-Z0-9_]*" /tmp/orx_sass.txt | sort | uniq -c
echo "# legacy tensor path (mma.sync / wmma -> bare HMMA): $(grep -cE '[^C]HMMA' /tmp/orx_sass.txt) instructions (UTCHMMA above is tcgen05)"
echo "# binary64 canonical rescoring: DADD $(grep -c DADD /tmp/orx_sass.txt), DMUL $(grep -c DMUL /tmp/orx_sass.txt)"
echo "# per kernel (Function : name -> tensor / TMA / TMEM instruction counts)"
awk '/Function :/ {name=$3} /UTCHMMA|UTMALDG|LDTM|UTCBAR/ {split($0,a," "); c[name]++} END {for (n in c) printf "%6d  %s\n", c[n], n}' /tmp/orx_sass.txt | sort -k2 | c++filt | cut -c1-200
echo "# PTX-level: the library is compiled for sm_100a only"
cuobjdump -lelf $LIB | head -20
