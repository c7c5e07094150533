#!/bin/bash
# Evidence that libffr_b200.so is Blackwell-native (B200_PROFILING.md, "What proves a Blackwell-native kernel"): SASS mnemonic
# counts of the in-tree library, and the ptxas register / spill table of every shipped kernel instantiation.
#   tools/sass_proof.sh > profiles/r02_sass_proof.txt
set -e
cd "$(dirname "$0")/.."
LIB=face_detection_and_recognition_b200/libffr_b200.so
echo "== cuobjdump -sass $LIB | grep -c <mnemonic>  ($(date -u +%Y-%m-%dT%H:%MZ), $(nvcc --version | tail -1))"
cuobjdump -sass $LIB > /tmp/ffr_sass.txt
for m in 'UTCHMMA' 'UTCHMMA\.2CTA' 'LDTM' 'UTMALDG' 'UTMALDG\.2D\.2CTA' 'UTCBAR' 'SYNCS\.' '[^C]HMMA' 'HGMMA' 'ACQBULK|UBLKCP'; do
  printf "%-22s %s\n" "$m" "$(grep -cE "$m" /tmp/ffr_sass.txt || true)"
done
echo "(UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG = cp.async.bulk.tensor; [^C]HMMA = legacy mma.sync: must be 0)"
echo "code objects: $(grep -c 'arch = sm_100a' /tmp/ffr_sass.txt) x sm_100a"
echo
echo "== ptxas -v (registers / spills) per kernel, nvcc -gencode arch=compute_100a,code=sm_100a -O3"
for f in ffr_filter_mma ffr_recheck ffr_l2norm ffr_filter_fp32 ffr_gallery ffr_dedup; do
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xptxas -v -c face_detection_and_recognition_b200/csrc/$f.cu -o /tmp/ffr_tmp.o 2>&1 \
   | grep -E "Compiling entry|Used|spill" | paste - - - \
   | sed -E 's/ptxas info    : Compiling entry function .(_ZN3ffr[0-9]+_GLOBAL__N__[0-9a-f_]+cu_[0-9a-f]+)?//; s/. for .sm_100a.//; s/ptxas info    ://g' \
   | awk -v f=$f '{print f ": " $0}'
done | c++filt 2>/dev/null | sed -E 's/\(CUtensorMap_st[^)]*\)//; s/ffr::\(anonymous namespace\):://g; s/\([^()]*\)//g; s/[ \t]+/ /g' 
