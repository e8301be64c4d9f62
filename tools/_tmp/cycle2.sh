python tools/regress_hash.py gpurun_out/hash_new.json tools/_tmp/hash_base.json 2>/dev/null | grep -c same
python tools/regress_hash.py gpurun_out/hash_new.json tools/_tmp/hash_base.json 2>/dev/null | grep DIFF
SEGS=0 python tools/devbench9.py 2>&1 | grep -v dense
WWZB200_NO_FUSED_EPILOGUE=1 SEGS=0 python tools/devbench9.py 2>&1 | grep -v dense
