#!/bin/bash
# tile-level culling: parity tests, headline bench, stats
TAG=${1:-r02_v22}; OUT=gpurun_out; mkdir -p $OUT
timeout 1200 python -m pytest tests -m gpu -q -x > $OUT/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_pytest.log
tail -12 $OUT/${TAG}_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $OUT/${TAG}_bench_ours.json 2> $OUT/${TAG}_bench_ours.err
FDGS_TILE_CULL=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-parity > $OUT/${TAG}_bench_ours_reflists.json 2> $OUT/${TAG}_bench_ours_reflists.err
python - <<PY
import json
for f in ("bench_ours", "bench_ours_reflists"):
    try:
        d = json.loads(open("$OUT/${TAG}_%s.json" % f).read().strip().splitlines()[-1])
        print(f, d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], {k: round(v, 4) for k, v in (d.get("stage_ms") or {}).items()})
        print("   stats", d.get("stats"))
        print("   parity", {k: v for k, v in (d.get("parity") or {}).items() if k != "gradients"})
    except Exception as e:
        print(f, "FAILED", e); print(open("$OUT/${TAG}_%s.err" % f).read()[-1500:])
PY
