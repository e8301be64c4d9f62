python -m pytest tests/test_gpu_batch.py -m gpu -x -q 2>&1 | tail -3
python tools/signif_prof.py
REPS=1 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_signif.csv python tools/signif_prof.py > gpurun_out/ncu_signif.log 2>&1; tail -1 gpurun_out/ncu_signif.log
