for wl in c4 c5; do python bench.py --workload $wl --steps 3 --warmup 3 --cpu-budget 0 --no-t2 --e2e-steps 1 --t0-sample 32768 --tier $( [ $wl = c5 ] && echo lu || echo auto ) 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$wl', d['config']['tier'], round(d['ms_per_step'],3), 'frac', round(d['roofline']['frac'],4), 'achieved', round(d['roofline']['achieved'],2), 'T0 sample TF/s', d['kernels'].get('t0_lu_per_pair',{}).get('achieved_tflops'), d['parity']['max_rel_err'])
"; done
