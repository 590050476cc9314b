#!/bin/bash
# Per-kernel SASS evidence of the Blackwell-native path (B200_PROFILING.md "What proves a Blackwell-native kernel"):
# counts of UTCHMMA (tcgen05.mma kind::f16), LDTM (tcgen05.ld), UTCBAR (tcgen05.commit), UBLKCP (cp.async.bulk),
# SYNCS (mbarrier), LDGSTS (cp.async), HMMA (legacy mma.sync - must be 0), FFMA / FFMA2 (packed fp32) per kernel of libnisqa_b200.so.
#   bash tools/sass_summary.sh > profiles/sass_summary.txt          (no GPU needed)
cd "$(dirname "$0")/.."
LIB=nisqa_b200/libnisqa_b200.so
echo "# cuobjdump -sass $LIB ($(date -u +%Y-%m-%dT%H:%MZ)), source digest $(python -c 'from nisqa_b200 import build; print(build.kernel_digest()[:16])')"
cuobjdump -sass $LIB | awk '
  /Function :/ { name=$3; order[++n]=name }
  /UTCHMMA/ {a[name]++} /LDTM/ {b[name]++} /UTCBAR/ {c[name]++} /UBLKCP/ {d[name]++} /SYNCS/ {e[name]++}
  /LDGSTS/ {f[name]++} / HMMA/ {g[name]++} /FFMA /  {h[name]++} /FFMA2/ {p2[name]++} /UTMALDG|UTMASTG/ {t[name]++} /STTM/ {u[name]++}
  END {
    printf "%-8s %-6s %-6s %-7s %-6s %-6s %-5s %-7s %-6s %-6s %-6s %s\n", "UTCHMMA","LDTM","STTM","UTCBAR","UBLKCP","SYNCS","HMMA","UTMA*","LDGSTS","FFMA","FFMA2","kernel";
    for (i=1;i<=n;i++) { k=order[i];
      printf "%-8d %-6d %-6d %-7d %-6d %-6d %-5d %-7d %-6d %-6d %-6d %s\n", a[k],b[k],u[k],c[k],d[k],e[k],g[k],t[k],f[k],h[k],p2[k],k; }
  }' | (read hdr; echo "$hdr"; c++filt | sed 's/(.*//' )
