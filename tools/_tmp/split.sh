for s in 1 2 3 4; do echo "== TAU_SPLIT=$s"; WWZB200_TAU_SPLIT=$s SEGS=0 python tools/devbench9.py 2>&1 | grep -v dense; done
