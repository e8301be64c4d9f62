# developer loop: regression hashes + timings with the production build, then the item trace with the -DWWZ_TRACE build
python tools/regress_hash.py gpurun_out/hash_new.json tools/_tmp/hash_base.json 2>/dev/null | grep -c same
python tools/regress_hash.py gpurun_out/hash_new.json tools/_tmp/hash_base.json 2>/dev/null | grep DIFF
SEGS=${SEGS:-0} python tools/devbench9.py 2>&1 | grep -v dense
cp tools/_tmp/lib_trace.so pyleoclim_util_b200/lib/libwwz_b200.so
SEGS=${TSEGS:-2} python tools/trace_items.py 2>&1 | grep -A8 "^c=" | grep "^c=\|per chunk\|point loops\|pipe share\|first-chunk"
