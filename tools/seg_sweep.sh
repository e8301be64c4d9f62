for s in 1 2 3 5 8; do WWZB200_POINT_SEGMENTS=$s python - <<'PY'
import os,sys
sys.path.insert(0,'.')
import tools.devbench as d
from pyleoclim_util_b200 import synth
t,y,tau,freq=synth.config_grid("C2")
d.run("C2 dense segs="+os.environ["WWZB200_POINT_SEGMENTS"], t,y,tau,freq,flags=1)
d.run("C2 window segs="+os.environ["WWZB200_POINT_SEGMENTS"], t,y,tau,freq,flags=0)
PY
done
