#!/bin/bash
# N-GPU session: NCCL parity of the band-sharded path, then bench.py at N ranks (frame sharding + per-config extras)
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/tiled_parity_check.py > gpurun_out/r2_tiled_parity_n$N.log 2>&1; echo "parity exit $?"; grep "band-sharded" gpurun_out/r2_tiled_parity_n$N.log | head -20
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err; echo "bench exit $?"
python - $N <<'PY'
import json, sys
n=sys.argv[1]
line=[l for l in open(f'gpurun_out/r2_bench_n{n}.json') if l.startswith('{')][-1]
d=json.loads(line)
print("N", d["n_gpus"], "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "host", d["host"])
for k,v in d["configs"].items(): print("  ", k, {a:(round(b,3) if isinstance(b,float) else b) for a,b in v.items() if not isinstance(b,dict)})
PY
tail -3 gpurun_out/r2_bench_n$N.err
