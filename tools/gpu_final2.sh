# final single-GPU evidence of round 2 (session 3): whole GPU test suite, smoke, bench (both arms), launch list of the
# bench, ncu capture of the fused recurrence kernel behind the streamed host call
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/r2c_gputests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py --steps 20 --warmup 3 > gpurun_out/r2c_bench_n1.json 2> gpurun_out/r2c_bench_n1.err; tail -2 gpurun_out/r2c_bench_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2c_bench_ref.json 2>> gpurun_out/r2c_bench_n1.err
WWZ_BENCH_SKIP_NUMBA=1 python bench.py --steps 5 --warmup 3 > gpurun_out/plain.log 2>&1 && WWZ_BENCH_SKIP_NUMBA=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2c_launches.csv python bench.py --steps 5 --warmup 3 > gpurun_out/ncu.log 2>&1
tail -2 gpurun_out/ncu.log
WWZ_FLAGS=0 python tools/ncu_one.py > gpurun_out/plain3.log 2>&1 && WWZ_FLAGS=0 ncu --set full --clock-control none --import-source on -k regex:wwz_tile_kernel -s 2 -c 2 -o gpurun_out/r2c_prof_rec_fused -f python tools/ncu_one.py > gpurun_out/ncu3.log 2>&1
tail -2 gpurun_out/ncu3.log
python tools/stream_bench.py 2>&1 | tee gpurun_out/r2c_stream_bench.log
