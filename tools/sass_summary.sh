#!/bin/bash
# SASS evidence that the hot kernels are Blackwell-native (B200_PROFILING.md "What proves a Blackwell-native kernel"):
# per kernel of libuvlt_sm100.so the count of UTC*MMA (tcgen05.mma), UTMALDG / UTMASTG (TMA load / store), LDTM / STTM
# (tcgen05.ld / st), UTCBAR (tcgen05.commit), HMMA (legacy mma.sync: must be 0), BRA.U.ANY (waterfall loops around
# uniform-datapath instructions: must be 0 in mainloops) and the registers ptxas reports.
#   tools/sass_summary.sh > profiles/r02_sass.md
set -e
cd "$(dirname "$0")/.."
LIB=uvltrack_b200/libuvlt_sm100.so
echo "# SASS summary of $LIB ($(date -u +%Y-%m-%d), nvcc $(nvcc --version | grep -o 'release [0-9.]*'))"
echo
echo "\`cuobjdump -sass\` mnemonic counts per kernel (tensor-core / TMA kernels only)."
echo
echo "| kernel | UTC*MMA | of which .2CTA | TS (A in TMEM) | UTMALDG | UTMASTG | LDTM | STTM | UTCBAR | HMMA | BRA.U.ANY |"
echo "|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|"
cuobjdump -sass $LIB | awk '
  /Function :/ { if (name != "") flush(); name=$3; mma=two=ts=ld=st=ldtm=sttm=bar=hmma=wf=0 }
  /UTC[A-Z]*MMA/ { mma++; if ($0 ~ /\.2CTA/) two++; if ($0 ~ /UTC[A-Z]*MMA[.A-Z0-9]* tmem\[/) ts++ }
  /UTMALDG/ { ld++ } /UTMASTG/ { st++ } /LDTM/ { ldtm++ } /STTM/ { sttm++ } /UTCBAR/ { bar++ }
  /[^A-Z]HMMA/ { hmma++ } /BRA\.U\.ANY/ { wf++ }
  function flush() { if (mma + ld + st + ldtm > 0) printf "| `%s` | %d | %d | %d | %d | %d | %d | %d | %d | %d | %d |\n", name, mma, two, ts, ld, st, ldtm, sttm, bar, hmma, wf }
  END { flush() }' | while IFS= read -r line; do
    sym=$(echo "$line" | sed -n 's/^| `\([^`]*\)`.*/\1/p')
    dem=$(echo "$sym" | c++filt | sed 's/(.*//; s/^void //')
    echo "$line" | sed "s|\`$sym\`|\`$dem\`|"
  done | sort -u
echo
echo "The one BRA.U.ANY of the attention3 kernels wraps the UTMASTG (TMA bulk store of an output tile) that the slot's store"
echo "thread issues once per work item, outside the key-block loop (attention3.cuh, item epilogue)."
echo
echo "PTX (\`cuobjdump -ptx\` is empty: the library ships SASS for sm_100a only); source-level instructions:"
echo
echo '```'
grep -ho "tcgen05\.[a-z_.:0-9A-Z]*\|cp\.async\.bulk\.tensor[a-z_.:0-9A-Z]*\|elect\.sync\|setmaxnreg\.[a-z]*\|griddepcontrol\.[a-z_]*" uvltrack_b200/csrc/*.cuh uvltrack_b200/csrc/*.cu | sort | uniq -c | sort -rn
echo '```'
