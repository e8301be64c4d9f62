set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python -c "import __graft_entry__ as g; g.smoke()"
python bench.py --steps 20 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err; cat gpurun_out/bench_ref.json
python bench.py --steps 5 --warmup 3 > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches.csv python bench.py --steps 5 --warmup 3 > gpurun_out/ncu.log 2>&1
tail -2 gpurun_out/ncu.log
python tools/ncu_one.py > gpurun_out/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:wwz_cells -s 2 -c 2 -o gpurun_out/prof_cells -f python tools/ncu_one.py > gpurun_out/ncu2.log 2>&1
tail -3 gpurun_out/ncu2.log
WWZ_FLAGS=0 python tools/ncu_one.py > gpurun_out/plain3.log 2>&1 && WWZ_FLAGS=0 ncu --set full --clock-control none --import-source on -k regex:wwz_rec -s 2 -c 2 -o gpurun_out/prof_rec -f python tools/ncu_one.py > gpurun_out/ncu3.log 2>&1
tail -3 gpurun_out/ncu3.log
REPS=1 S=2048 python tools/signif_prof.py > gpurun_out/plain4.log 2>&1 && REPS=1 S=2048 ncu --set full --clock-control none --import-source on -k regex:shared_mma -s 1 -c 1 -o gpurun_out/prof_mma -f python tools/signif_prof.py > gpurun_out/ncu4.log 2>&1
tail -3 gpurun_out/ncu4.log
