#!/usr/bin/env bash
# Counts the Blackwell-native instructions per kernel in the built library (no GPU needed):
# UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UBLKCP = cp.async.bulk (TMA 1-D), UTMALDG/UTMASTG = tensor-map TMA;
# "legacy" = HMMA (mma.sync / wmma) or *GMMA (wgmma): must be zero.
set -euo pipefail
cd "$(dirname "$0")/.."
so=defensive-model-vae_b200/csrc/libdmvae.so
cuobjdump -sass "$so" | c++filt | awk '
  /Function :/ { name=$0; sub(/.*Function : /, "", name); sub(/\(.*/, "", name); sub(/^void /, "", name) }
  / UTC[A-Z]*MMA/ { mma[name]++ }
  / LDTM| STTM/ { tm[name]++ }
  / UBLKCP/ { blk[name]++ }
  / UTMALDG| UTMASTG/ { tma[name]++ }
  / HMMA| HGMMA| QGMMA| IGMMA/ { legacy[name]++ }
  /^ +\/\*[0-9a-f]+\*\/ +[A-Z@]/ { n[name]++ }
  END { printf "%-40s %8s %8s %9s %8s %8s %7s\n", "kernel", "instrs", "UTC*MMA", "LDTM/STTM", "UBLKCP", "UTMA*", "legacy";
        for (k in n) printf "%-40s %8d %8d %9d %8d %8d %7d\n", k, n[k], mma[k], tm[k], blk[k], tma[k], legacy[k] }' | sort -r
