#!/bin/bash
# Per-kernel counts of the Blackwell-specific SASS mnemonics in the shipped library (B200_PROFILING.md: UTCHMMA = tcgen05.mma,
# LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG = TMA tensor load/store, UBLKCP = cp.async.bulk, UBLKPF = bulk L2 prefetch,
# HMMA = legacy mma.sync).   tools/sass_summary.sh [lib] > profiles/rN/sass_summary.txt
LIB=${1:-gnn-formation-control_b200/libgfc.so}
echo "# $(basename $LIB): $(stat -c %s $LIB) bytes, $(date -u +%F)"
cuobjdump -sass "$LIB" | awk '
  /Function :/ { fn=$3; sub(/^_ZN3gfc[0-9]*/,"",fn); sub(/I[LN].*$/,"",fn); sub(/E?v?P[KF].*$/,"",fn); names[fn]=1; next }
  { for (m in pat) if ($0 ~ pat[m]) cnt[fn,m]++ }
  BEGIN { pat["UTCHMMA"]="UTCHMMA"; pat["UTCQMMA"]="UTCQMMA"; pat["LDTM"]="LDTM"; pat["STTM"]="STTM"; pat["UTMALDG"]="UTMALDG"; pat["UTMASTG"]="UTMASTG";
          pat["UBLKCP"]="UBLKCP"; pat["UBLKPF"]="UBLKPF"; pat["HMMA"]="[^C]HMMA"; pat["SYNCS"]="SYNCS"; pat["REDG"]="REDG"; pat["UTCBAR"]="UTCBAR" }
  END { printf "%-70s", "kernel"; n=split("UTCHMMA LDTM STTM UTMALDG UTMASTG UBLKCP UBLKPF UTCBAR SYNCS HMMA REDG", ord, " ");
        for (i=1;i<=n;i++) printf "%8s", ord[i]; printf "\n";
        for (f in names) { tot=0; for (i=1;i<=n;i++) tot+=cnt[f,ord[i]]; if (tot==0) continue;
          printf "%-70s", substr(f,1,70); for (i=1;i<=n;i++) printf "%8d", cnt[f,ord[i]]; printf "\n" } }' | (read -r l1; echo "$l1"; read -r l2; echo "$l2"; sort)
