#!/bin/bash
# SASS evidence for the Blackwell-specific paths: counts of tcgen05 / TMEM / TMA mnemonics per kernel of libb200ov.so.
#   tools/sass_summary.sh > profiles/sass_summary.txt
LIB=${1:-pyopenvino_b200/libb200ov.so}
echo "# cuobjdump -sass $LIB | per-kernel counts of UTCHMMA (tcgen05.mma) / UTCBAR (tcgen05.commit) / LDTM / STTM (tcgen05.ld/st),"
echo "# UTMALDG / UTMASTG (TMA tensor load / store), UBLKCP (bulk copy), SYNCS (mbarrier), FFMA2 / FADD2 / FMUL2 (packed FP32), FHFMA"
echo "# built: $(date -u +%Y-%m-%dT%H:%MZ)  nvcc: $(nvcc --version | tail -1)"
cuobjdump -sass "$LIB" | awk '
  /Function :/ { if (name != "") flush(); name=$3; delete c; next }
  { for (m in want) if (index($0, m)) c[m]++ }
  function flush(   line, any) {
    line=""; any=0
    for (i = 1; i <= n; i++) { m=order[i]; if (c[m] > 0) { line = line sprintf(" %s=%d", m, c[m]); any=1 } }
    if (any) printf "%-110s%s\n", name, line
  }
  BEGIN { n=split("UTCHMMA UTCBAR LDTM STTM UTMALDG UTMASTG UTMAPF UBLKCP SYNCS FFMA2 FADD2 FMUL2 FHFMA F2FP", order, " "); for (i=1;i<=n;i++) want[order[i]]=1 }
  END { if (name != "") flush() }
' | c++filt | sed 's/b200ov:://g' | sort
