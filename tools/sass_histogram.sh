#!/bin/bash
# SASS opcode histogram of the shipped library (what proves the Blackwell-native paths: UTMALDG / UBLKCP / UBLKPF = TMA and bulk
# copies, SYNCS = mbarrier, LDG/STG.E.128 with cache hints, DFMA/DADD/DMUL = the FP64 butterflies; no HMMA/UTC*MMA: no tensor cores)
LIB=${1:-mixed-radix-fast-fourier-transform_b200/libhsfft_b200.so}
echo "# $(basename $LIB)  $(stat -c %s $LIB) bytes  $(date -u +%Y-%m-%dT%H:%MZ)"
cuobjdump -lelf $LIB | grep -c "sm_100a" | sed 's/^/# sm_100a cubins: /'
cuobjdump -sass $LIB | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed -E 's/^\s+\/\*[0-9a-f]+\*\/\s+(@!?U?P[0-9T]+\s+)?//' | awk '{print $1}' | sed -E 's/;$//' | awk -F. '{print $1}' | sort | uniq -c | sort -rn
