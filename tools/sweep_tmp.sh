python -m pytest tests/test_gpu_batch.py -m gpu -x -q 2>&1 | tail -3
WWZB200_STORED_WEIGHTS=1 python -m pytest tests/test_gpu_batch.py -m gpu -x -q 2>&1 | tail -3
WWZ_FLAGS=1 python tools/signif_prof.py
python tools/signif_prof.py
WWZB200_STORED_WEIGHTS=0 python tools/signif_prof.py
