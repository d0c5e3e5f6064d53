#!/bin/bash
# SASS opcode evidence for the shipped library: counts of the Blackwell-specific instructions per kernel family
# (tcgen05.mma = UTCHMMA / UTCQMMA..., tcgen05.ld = LDTM, bulk-async TMA copies = UBLKCP, L2 prefetch = UBLKPF,
# multicast commits = UTCBAR, legacy warp MMA of the LSTM = HMMA).   usage: tools/sass_histogram.sh [lib.so] > profiles/sass_opcodes_rNN.txt
LIB=${1:-ml_audio_restoration_b200/libaudiorestore_sm100.so}
echo "# cuobjdump -sass $LIB  (sm_100a), $(date -u +%Y-%m-%dT%H:%MZ)"
/usr/local/cuda/bin/cuobjdump -sass "$LIB" > /tmp/sass_all.txt
echo "## whole library"
for op in UTCHMMA.2CTA UTCHMMA LDTM UBLKCP.S.G UBLKPF.L2 UTCBAR UTCATOMSWS HMMA.16816.F32 UTMALDG SYNCS.ARRIVE F2FP.SATFINITE FFMA2 FADD2 FMUL2 MUFU.EX2 MUFU.RCP ELECT; do
  printf "%-18s %6d\n" "$op" "$(grep -c "$op" /tmp/sass_all.txt)"
done
echo "## per kernel (Function : name -> UTCHMMA / LDTM / UBLKCP / HMMA counts)"
awk '/Function :/ {name=$3} /UTCHMMA/ {m[name]++} /LDTM/ {l[name]++} /UBLKCP/ {b[name]++} /HMMA\.16816/ {h[name]++} END {for (n in m) printf "%-110s UTCHMMA %4d LDTM %4d UBLKCP %4d\n", n, m[n], l[n], b[n]; for (n in h) printf "%-110s HMMA %4d\n", n, h[n]}' /tmp/sass_all.txt | c++filt | sed 's/(ar::[A-Za-z]*, ar::[A-Za-z0-9]*, int)//' | sort
