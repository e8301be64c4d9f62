python -m pytest tests/test_gpu_batch.py -m gpu -x -q 2>&1 | tail -3
python tools/signif_prof.py
WWZB200_QUANTILE_RADIX=1 python tools/signif_prof.py
