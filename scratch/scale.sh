#!/bin/bash
# 1/2/4/8-GPU scaling on ONE 8-GPU box: bench lines into gpurun_out/r02_scale_N.json (C4, rank-update tier),
# r02_scale_c5_lu_N.json (brute-force tier), and the NCCL fallback at 8 for comparison.
O=gpurun_out
P=29600
run() {  # N workload-args out
  N=$1; shift; OUT=$1; shift
  if [ "$N" = 1 ]; then python bench.py --gpus 1 "$@" > $O/$OUT.json 2> $O/$OUT.err
  else P=$((P+1)); python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N "$@" > $O/$OUT.json 2> $O/$OUT.err; fi
}
for N in 1 2 4 8; do run $N r02_scale_$N --steps 20 --warmup 3 --cpu-budget 0 --t0-sample 0 --no-t2; done
run 8 r02_scale_8_nccl --steps 20 --warmup 3 --cpu-budget 0 --t0-sample 0 --no-t2 --collective nccl
for N in 1 8; do run $N r02_scale_c5_lu_$N --workload c5 --tier lu --steps 5 --warmup 3 --cpu-budget 0 --t0-sample 0 --no-t2; done
python - <<'PY'
import json
base=None
for f in ["r02_scale_1","r02_scale_2","r02_scale_4","r02_scale_8","r02_scale_8_nccl","r02_scale_c5_lu_1","r02_scale_c5_lu_8"]:
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, "N", d["n_gpus"], "ms/step", round(d["ms_per_step"],4), "e2e ms", round(d["e2e"]["ms_per_call"],4), "parity", d["parity"]["max_rel_err"] if d.get("parity") else None, d["config"].get("collective"))
    except Exception as e:
        print(f, "ERR", e)
PY
