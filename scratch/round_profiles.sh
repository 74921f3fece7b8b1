#!/bin/bash
# One gpurun call: bench lines of every workload + the ncu evidence that goes under profiles/ (round 2).
set -x
O=gpurun_out
R=r02
python bench.py --steps 20 --warmup 3 > $O/${R}_bench_c4.json 2> $O/${R}_bench_c4.err
python bench.py --impl reference --steps 2 --warmup 1 > $O/${R}_bench_c4_reference.json 2> $O/${R}_bench_ref.err
python bench.py --workload c3 --steps 20 --warmup 3 > $O/${R}_bench_c3.json 2> $O/${R}_bench_c3.err
python bench.py --workload c2 --steps 20 --warmup 3 > $O/${R}_bench_c2.json 2> $O/${R}_bench_c2.err
python bench.py --workload c3 --tier lu --steps 5 --warmup 3 --cpu-budget 0 --no-t2 > $O/${R}_bench_c3_lu.json 2>> $O/${R}_bench_c3.err
python bench.py --workload c5 --tier lu --steps 5 --warmup 3 > $O/${R}_bench_c5_lu.json 2> $O/${R}_bench_c5.err
for k in 32 100 316; do python bench.py --workload c5 --tier lu --k-exc $k --steps 5 --warmup 3 --cpu-budget 0 --no-t2 --t0-sample 0 > $O/${R}_bench_c5_lu_k$k.json 2>> $O/${R}_bench_c5.err; done
python bench.py --workload traj --steps 5 --warmup 2 > $O/${R}_bench_traj.json 2> $O/${R}_bench_traj.err
# launch list (cold-cache serialised per-launch times: compare shares only)
python bench.py --steps 2 --warmup 3 --e2e-steps 1 --cpu-budget 0 --t0-sample 2368 --no-check > $O/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/${R}_launches.csv \
    python bench.py --steps 2 --warmup 3 --e2e-steps 1 --cpu-budget 0 --t0-sample 2368 --no-check > $O/ncu1.log 2>&1
# full captures of the dominant kernels
python scratch/one_call.py c4 auto 1 > $O/plain2.log 2>&1 && \
ncu --set full --import-source on --clock-control none -k regex:t1_fused -c 1 -s 2 -o /tmp/${R}_t1 -f python scratch/one_call.py c4 auto 1 > $O/ncu2.log 2>&1
python scratch/one_call.py c4 lu 1 > $O/plain3.log 2>&1 && \
ncu --set full --import-source on --clock-control none -k regex:t0_fused -c 1 -s 2 -o /tmp/${R}_t0n100 -f python scratch/one_call.py c4 lu 1 > $O/ncu3.log 2>&1
python scratch/one_call.py c5 lu 1 > $O/plain4.log 2>&1 && \
ncu --set full --import-source on --clock-control none -k regex:t0_fused -c 1 -s 2 -o /tmp/${R}_t0n64 -f python scratch/one_call.py c5 lu 1 > $O/ncu4.log 2>&1
for k in t1 t0n100 t0n64; do ncu -i /tmp/${R}_$k.ncu-rep --page raw --csv > $O/${R}_${k}_raw.csv 2>/dev/null; done
ls -la $O | tail -30
