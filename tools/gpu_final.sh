# final evidence of the round on one B200: whole GPU test suite, smoke, bench (both arms)
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/r2_gputests.log
python -c "import __graft_entry__ as g; g.smoke()"
python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; tail -2 gpurun_out/r2_bench_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref.json 2>> gpurun_out/r2_bench_n1.err
python tools/ensemble_prof.py 2>&1 | tail -8
