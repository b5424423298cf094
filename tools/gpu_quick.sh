#!/bin/bash
# quick GPU check of a kernel change: targeted parity tests + stage times of the bench
# usage: bash tools/gpu_quick.sh <tag> ['pytest -k expression'] ; extra env (e.g. FDGS_BWD2_REGS=1) is inherited
TAG=${1:-q}; KEXPR=${2:-"colour_only or golden or compiled_reference"}
mkdir -p gpurun_out
if [ "$KEXPR" != "none" ]; then
timeout 900 python -m pytest tests -m gpu -q -k "$KEXPR" > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
grep -E "^E  |passed|failed|Error" gpurun_out/${TAG}_pytest.log | head -30
fi
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_bench.json").read().strip().splitlines()[-1])
print("${TAG} value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"])
print("stage_ms", {k: round(v,4) for k,v in d["stage_ms"].items()})
PY
