# round-2 evidence run on one B200: tests, smoke, bench (both arms), launch list, ncu captures of the three hot kernels
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/r2_gputests.log
python -c "import __graft_entry__ as g; g.smoke()"
python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; tail -2 gpurun_out/r2_bench_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref.json 2>> gpurun_out/r2_bench_n1.err
python bench.py --steps 5 --warmup 3 > gpurun_out/plain.log 2>&1 && WWZ_BENCH_SKIP_NUMBA=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 5 --warmup 3 > gpurun_out/ncu.log 2>&1
tail -2 gpurun_out/ncu.log
WWZ_FLAGS=0 python tools/ncu_one.py > gpurun_out/plain3.log 2>&1 && WWZ_FLAGS=0 ncu --set full --clock-control none --import-source on -k regex:wwz_tile_kernel -s 2 -c 2 -o gpurun_out/r2_prof_rec -f python tools/ncu_one.py > gpurun_out/ncu3.log 2>&1
tail -2 gpurun_out/ncu3.log
python tools/ncu_one.py > gpurun_out/plain2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:wwz_cells -s 2 -c 2 -o gpurun_out/r2_prof_cells -f python tools/ncu_one.py > gpurun_out/ncu2.log 2>&1
tail -2 gpurun_out/ncu2.log
REPS=1 S=2048 python tools/signif_prof.py > gpurun_out/plain4.log 2>&1 && REPS=1 S=2048 ncu --set full --clock-control none --import-source on -k regex:shared_mma -s 1 -c 1 -o gpurun_out/r2_prof_mma -f python tools/signif_prof.py > gpurun_out/ncu4.log 2>&1
tail -2 gpurun_out/ncu4.log
M=250 python tools/ensemble_prof.py > gpurun_out/plain5.log 2>&1 && M=250 ncu --set full --clock-control none --import-source on -k regex:wwz_tile_kernel -s 4 -c 2 -o gpurun_out/r2_prof_ens -f python tools/ensemble_prof.py > gpurun_out/ncu5.log 2>&1
tail -2 gpurun_out/ncu5.log
