for s in 6 8 12 16; do echo "== segs $s"; WWZB200_POINT_SEGMENTS=$s python tools/stream_bench.py 2>&1 | sed -n 3,5p; done
