#!/bin/bash
# SASS evidence: which Blackwell / Hopper+ instructions the kernels of libfastq_gpu.so use (cuobjdump runs without a GPU)
cd "$(dirname "$0")/.."
SO=fastq_utils_b200/libfastq_gpu.so
echo "# $(cuobjdump -lelf $SO | head -3 | tr '\n' ' ')"
cuobjdump -sass $SO > /tmp/fqg_sass.txt
echo "# kernels:"; grep -E "Function :" /tmp/fqg_sass.txt | sed -E 's/.*Function : //' | sed -E 's/_ZN[0-9]+_GLOBAL__N__[0-9a-f_]+cu_[0-9a-f]+[0-9]+//' | sort | uniq -c | awk '{print "  " $2}' | head -60
echo "# instruction counts over all kernels:"
for m in UBLKCP "UBLKPF" "SYNCS" "ATOMG.E.CAS.128" "ATOMG" "ATOMS" "RED.E" "IDP.4A" "VIMNMX" "MATCH.ANY" "REDUX" "LOP3.LUT" "PRMT" "SHF" "IMAD.HI" "LDS.128" "STG.E.128" "LDG.E.128" "STRONG.SYS" "NANOSLEEP" "HMMA" "UTC.*MMA"; do
  printf "  %-40s %s\n" "$m" "$(grep -cE "$m" /tmp/fqg_sass.txt)"
done
