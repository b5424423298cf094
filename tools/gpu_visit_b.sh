#!/bin/bash
# visit B of round 2: cull fix + AUX v2 backward: parity tests, headline bench, cfg3aux (v2 / v1 / reference), zero-fill overlap probe
TAG=${1:-r02_v21}; OUT=gpurun_out; mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -x > $OUT/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> $OUT/${TAG}_pytest.log
tail -5 $OUT/${TAG}_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/${TAG}_bench_ours.json 2> $OUT/${TAG}_bench_ours.err
timeout 600 python bench.py --workload cfg3aux --steps 20 --warmup 5 --no-cpu-baseline --no-parity > $OUT/${TAG}_cfg3aux_ours.json 2> $OUT/${TAG}_cfg3aux_ours.err
FDGS_BLEND_BWD_V1=1 timeout 600 python bench.py --workload cfg3aux --steps 20 --warmup 5 --no-cpu-baseline --no-parity > $OUT/${TAG}_cfg3aux_ours_v1.json 2> $OUT/${TAG}_cfg3aux_ours_v1.err
timeout 600 python bench.py --workload cfg3aux --impl reference --steps 10 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_cfg3aux_reference.json 2> $OUT/${TAG}_cfg3aux_reference.err
python - <<PY
import json
for f in ("bench_ours", "cfg3aux_ours", "cfg3aux_ours_v1", "cfg3aux_reference"):
    try:
        d = json.loads(open("$OUT/${TAG}_%s.json" % f).read().strip().splitlines()[-1])
        print(f, d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], {k: round(v, 4) for k, v in (d.get("stage_ms") or {}).items()})
    except Exception as e:
        print(f, "FAILED", e); print(open("$OUT/${TAG}_%s.err" % f).read()[-1500:])
PY
timeout 300 python tools/overlap_zero_probe.py > $OUT/${TAG}_overlap_probe.json 2> $OUT/${TAG}_overlap_probe.err
cat $OUT/${TAG}_overlap_probe.json; tail -3 $OUT/${TAG}_overlap_probe.err
