#!/bin/bash
# 2-GPU visit: NCCL view-parallel tests + bench at N=2 with phase tables
TAG=${1:-r02_v26}; OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_view_parallel.py -m gpu -q -x > $OUT/${TAG}_pytest_n2.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_pytest_n2.log
tail -5 $OUT/${TAG}_pytest_n2.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 --no-cpu-baseline --no-parity > $OUT/${TAG}_bench_n2.json 2> $OUT/${TAG}_bench_n2.err
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-parity > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench_n1.err
python - <<PY
import json
for f in ("bench_n2", "bench_n1"):
    try:
        d = json.loads(open("$OUT/${TAG}_%s.json" % f).read().strip().splitlines()[-1])
        print(f, d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], {k: round(v, 4) for k, v in (d.get("stage_ms") or {}).items()})
        ex = d.get("exchange") or {}
        for k in ("phase_ms", "phase_serial_ms"):
            if ex.get(k): print("  ", k, {a: round(b, 4) for a, b in ex[k].items()}, "sum", round(sum(ex[k].values()), 4))
        print("  ", {k: v for k, v in ex.items() if k not in ("phase_ms", "phase_serial_ms")})
    except Exception as e:
        print(f, "FAILED", e); print(open("$OUT/${TAG}_%s.err" % f).read()[-2500:])
PY
timeout 300 python bench.py --workload cfg5 --steps 20 --warmup 5 --no-cpu-baseline --no-parity > $OUT/${TAG}_cfg5_n1.json 2> $OUT/${TAG}_cfg5_n1.err
python - <<PY
import json
d = json.loads(open("$OUT/${TAG}_cfg5_n1.json").read().strip().splitlines()[-1])
print("cfg5 n1", d["value"], d["ms_per_step"], {k: round(v, 4) for k, v in ((d.get("exchange") or {}).get("phase_ms") or {}).items()})
PY
