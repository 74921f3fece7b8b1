#!/bin/bash
# refresh of the bench lines / captures that the last T0 change touches (the rest of scratch/round_profiles.sh stands)
O=gpurun_out
R=r02
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > $O/${R}_gputests.log
python bench.py --steps 20 --warmup 3 > $O/${R}_bench_c4.json 2> $O/${R}_bench_c4.err
python bench.py --workload c3 --tier lu --steps 5 --warmup 3 --cpu-budget 0 --no-t2 > $O/${R}_bench_c3_lu.json 2>> $O/${R}_bench_c3.err
python bench.py --workload c5 --tier lu --steps 5 --warmup 3 > $O/${R}_bench_c5_lu.json 2> $O/${R}_bench_c5.err
for k in 32 100 316; do python bench.py --workload c5 --tier lu --k-exc $k --steps 5 --warmup 3 --cpu-budget 0 --no-t2 --t0-sample 0 > $O/${R}_bench_c5_lu_k$k.json 2>> $O/${R}_bench_c5.err; done
python scratch/one_call.py c4 lu 1 > $O/plain3.log 2>&1 && \
ncu --set full --import-source on --clock-control none -k regex:t0_fused -c 1 -s 2 -o /tmp/${R}_t0n100 -f python scratch/one_call.py c4 lu 1 > $O/ncu3.log 2>&1
python scratch/one_call.py c5 lu 1 > $O/plain4.log 2>&1 && \
ncu --set full --import-source on --clock-control none -k regex:t0_fused -c 1 -s 2 -o /tmp/${R}_t0n64 -f python scratch/one_call.py c5 lu 1 > $O/ncu4.log 2>&1
for k in t0n100 t0n64; do ncu -i /tmp/${R}_$k.ncu-rep --page raw --csv > $O/${R}_${k}_raw.csv 2>/dev/null; done
cat $O/${R}_gputests.log
