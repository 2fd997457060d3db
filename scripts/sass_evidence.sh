#!/usr/bin/env bash
# SASS excerpts proving what the shipped binaries execute (profiles/r2_sass_*.txt): tcgen05 / TMEM / TMA mnemonics of the
# hand-written GEMM kernels in libptk.so, and the 128-bit global accesses of the generated cfg2 kernel (NVRTC cubin cache).
set -eu
cd "$(dirname "$0")/.."
out=profiles/r2_sass_libptk_gemm.txt
{
  echo "# cuobjdump -sass pytensor_b200/libptk.so | tensor-core / TMEM / TMA / mbarrier / cluster instructions per kernel (count mnemonic)"
  echo "# UTCHMMA[.2CTA] = tcgen05.mma, LDTM = tcgen05.ld (TMEM -> registers), UTMALDG = cp.async.bulk.tensor (TMA), UTCBAR = tcgen05.commit,"
  echo "# UTCATOMSWS = tcgen05.alloc/dealloc, SYNCS = mbarrier, UCGABAR = barrier.cluster"
  cuobjdump -sass pytensor_b200/libptk.so | awk '/Function :/{f=$3} /UTCHMMA|UTCBAR|LDTM|UTMALDG|UTCATOMSWS|UCGABAR|SYNCS\./{m=$2; sub(/^\/\*[0-9a-f]+\*\//,"",m); c[f" "m]++} END{for(k in c) print c[k], k}' | sort -k2,2 -k1,1nr | grep -i gemm | c++filt | sed 's/(anonymous namespace):://'
  echo
  echo "# first tcgen05.mma / tcgen05.ld / TMA lines of the cta_group::2 kernel, verbatim:"
  cuobjdump -sass pytensor_b200/libptk.so | awk '/Function :.*pair_kernel/{p=1} /Function :/{if($0!~/pair_kernel/)p=0} p&&/UTCHMMA|LDTM|UTMALDG|UTCBAR/{print}' | head -14
} > "$out"
echo "wrote $out"
k3=$(grep -l "ptk_ew_red_row" pytensor_b200/_kcache/*.cu | xargs ls -t | head -1)
out2=profiles/r2_sass_cfg2_k3.txt
{
  echo "# generated kernel of the cfg2 graph (${k3##*/}): global memory instructions of the main loop (cuobjdump -sass of the NVRTC cubin)"
  cuobjdump --dump-resource-usage "${k3%.cu}.cubin" | grep -E "Function|REG"
  cuobjdump -sass "${k3%.cu}.cubin" | grep -E "LDG|STG|CCTL|DADD|F2F|MUFU|BAR|SHFL" | awk '{print $2, $3, $4}' | sort | uniq -c | sort -rn
} > "$out2"
echo "wrote $out2"
