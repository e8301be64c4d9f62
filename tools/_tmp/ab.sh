# A/B: 8-cell lanes vs lane pairs on the bench (C2 step, C5, ensemble), numba leg skipped
for pair in 0 1; do
  echo "== WWZB200_REC_PAIR=$pair"
  WWZB200_REC_PAIR=$pair WWZ_BENCH_SKIP_NUMBA=1 python bench.py --steps 20 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('C2 ms/step %.4f  e2e ms %.3f  frac %.3f  C3 s %.4f  C5 ms %.1f  ens members/s %.0f' % (d['ms_per_step'], d['e2e']['ms_per_step'], d['roofline']['frac'], d['signif_test']['seconds'], d['config5']['default']['ms'], d['ensemble']['members_per_s']))
print('parity', json.dumps(d.get('parity'))[:600])
"
done
