#!/bin/sh
# tools/sass_tma.sh > profiles/r02_sass_tma.txt : mnemonic counts in the shipped library's SASS (what proves TMA, bulk
# copies, mbarriers and packed FP32 are really there; no tensor-core instruction is expected)
lib=${1:-picha_b200/libpicha_b200.so}
echo "# cuobjdump -sass $lib | grep -c <mnemonic>   ($(date -u +%Y-%m-%d), $(cuobjdump --version | tail -1))"
echo "# cubins: $(cuobjdump -lelf $lib | grep -c sm_100a) sm_100a, $(cuobjdump -lelf $lib | grep -vc sm_100a) other"
cuobjdump -sass $lib > /tmp/picha_sass.txt
for m in UTMALDG UTMAPF UBLKCP SYNCS.ARRIVE SYNCS.PHASECHK SYNCS.EXCH FFMA2 FMUL2 FFMA PRMT LDS.128 STS.128 LDCU "BRA.U" ACQBULK HMMA UTCHMMA UTCQMMA QGMMA; do
  printf "%-16s %s\n" "$m" "$(grep -c "$m" /tmp/picha_sass.txt)"
done
echo "# kernels (entry functions):"
grep "Function :" /tmp/picha_sass.txt | sed 's/.*Function : //' | c++filt | sed 's/(.*//' | sed 's/<.*//' | sort | uniq -c | sort -rn
