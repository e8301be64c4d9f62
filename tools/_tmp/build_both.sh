set -e
WWZ_NVCC_EXTRA="-DWWZ_TRACE" python -m pyleoclim_util_b200.build --force >/dev/null
cp pyleoclim_util_b200/lib/libwwz_b200.so tools/_tmp/lib_trace.so
python -m pyleoclim_util_b200.build --force >/dev/null
grep -A2 "wwz_tile_kernelILi" pyleoclim_util_b200/lib/ptxas.log | grep "spill\|Used"
