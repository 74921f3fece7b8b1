#!/bin/bash
# 2/4/8-GPU scaling on ONE 8-GPU box (short form: C4 only, rank-update tier): gpurun_out/r02_scale_N.json
O=gpurun_out
P=29700
for N in 8 4 2; do
  P=$((P+1))
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P bench.py --gpus $N --steps 20 --warmup 3 --cpu-budget 0 --t0-sample 0 --no-t2 > $O/r02_scale_$N.json 2> $O/r02_scale_$N.err
done
python bench.py --gpus 1 --steps 20 --warmup 3 --cpu-budget 0 --t0-sample 0 --no-t2 > $O/r02_scale_1.json 2> $O/r02_scale_1.err
python - <<'PY'
import json
for f in ["r02_scale_1","r02_scale_2","r02_scale_4","r02_scale_8"]:
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, "N", d["n_gpus"], "ms/step", round(d["ms_per_step"],4), "e2e ms", round(d["e2e"]["ms_per_call"],4), "parity", d["parity"]["max_rel_err"] if d.get("parity") else None, d["config"].get("collective"))
    except Exception as e:
        print(f, "ERR", e)
PY
