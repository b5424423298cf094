#!/bin/bash
# One GPU visit.  Usage (from the repo root, under gpurun):  bash tools/gpu_round.sh <tag> "<what ...>"
#   what = any of: golden tests bench ref workloads parity launches full small sanitize racecheck
TAG=${1:-r02_x}
WHAT=${2:-"tests bench ref"}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1
has() { [[ " $WHAT " == *" $1 "* ]]; }

if has golden; then
  # golden vectors of the configurations added in round 2, from the UNMODIFIED reference (oracle/_ref)
  timeout 600 python tools/first_light.py --configs dur10,smod05,smod2,rotcam,n3v,deg1m4 --golden $OUT/golden --no-timing \
      > $OUT/${TAG}_golden.log 2>&1
  cp $OUT/golden/*.npz tests/golden/ 2>/dev/null
  grep -E "golden written|diff=|CONFIG" $OUT/${TAG}_golden.log | tail -40
fi
if has tests; then
  timeout 900 python -m pytest tests -m gpu -q --durations=15 > $OUT/${TAG}_pytest.log 2>&1
  echo "pytest exit $?" >> $OUT/${TAG}_pytest.log
  tail -15 $OUT/${TAG}_pytest.log
fi
if has bench; then
  timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/${TAG}_bench_ours.json 2> $OUT/${TAG}_bench_ours.err
  tail -c 4000 $OUT/${TAG}_bench_ours.json; tail -3 $OUT/${TAG}_bench_ours.err
fi
if has ref; then
  timeout 600 python bench.py --impl reference --steps 10 --warmup 3 > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err
  tail -c 1200 $OUT/${TAG}_bench_ref.json; tail -3 $OUT/${TAG}_bench_ref.err
fi
if has workloads; then
  for wl in cfg1 cfg2 cfg5 cfg4; do
    for impl in ours reference; do
      timeout 600 python bench.py --workload $wl --impl $impl --steps 20 --warmup 5 --no-cpu-baseline \
          > $OUT/${TAG}_${wl}_${impl}.json 2> $OUT/${TAG}_${wl}_${impl}.err
      python - <<PY
import json
try:
    d = json.load(open("$OUT/${TAG}_${wl}_${impl}.json"))
    print("$wl $impl", d["value"], d["unit"], "ms/step", d["ms_per_step"], "e2e", d["e2e"]["value"])
except Exception as e:
    print("$wl $impl FAILED", e); print(open("$OUT/${TAG}_${wl}_${impl}.err").read()[-1500:])
PY
    done
  done
  for impl in ours reference; do
    timeout 600 python bench.py --workload cfg4 --rigid --impl $impl --steps 20 --warmup 5 --no-cpu-baseline \
        > $OUT/${TAG}_cfg4rigid_${impl}.json 2> $OUT/${TAG}_cfg4rigid_${impl}.err
    timeout 600 python bench.py --workload train3 --impl $impl --steps 10 --warmup 3 --no-cpu-baseline \
        > $OUT/${TAG}_train3_${impl}.json 2> $OUT/${TAG}_train3_${impl}.err
    python - <<PY
import json
for wl in ("cfg4rigid", "train3"):
    try:
        d = json.load(open("$OUT/${TAG}_%s_${impl}.json" % wl))
        print(wl, "$impl", d["value"], d["unit"], "ms/step", d["ms_per_step"], d.get("phase_ms"))
    except Exception as e:
        print(wl, "$impl FAILED", e); print(open("$OUT/${TAG}_%s_${impl}.err" % wl).read()[-1500:])
PY
  done
  timeout 600 python tools/neighbours_bench.py > $OUT/${TAG}_neighbours.md 2> $OUT/${TAG}_neighbours.err
  tail -40 $OUT/${TAG}_neighbours.md; tail -3 $OUT/${TAG}_neighbours.err
  # cfg1 once more with the CPU baselines (exact cfg1 on the host cores)
  timeout 600 python bench.py --workload cfg1 --steps 20 --warmup 5 > $OUT/${TAG}_cfg1_ours_cpu.json 2> $OUT/${TAG}_cfg1_ours_cpu.err
fi
if has parity; then
  timeout 1200 python tools/parity_table.py > $OUT/${TAG}_parity_table.md 2> $OUT/${TAG}_parity_table.err
  tail -30 $OUT/${TAG}_parity_table.md; tail -3 $OUT/${TAG}_parity_table.err
fi
if has launches; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
      --log-file $OUT/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity > $OUT/${TAG}_launches.log 2>&1
  python tools/launch_summary.py $OUT/${TAG}_launches.csv | tee $OUT/${TAG}_launch_summary.txt
fi
if has full; then
  # skip the warm-up launches of each kernel (-s counts matching launches only), capture one of each
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'blend_(fwd|bwd|bwd2)_kernel' -s 6 -c 2 \
      -f -o $OUT/${TAG}_blend python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity > $OUT/${TAG}_full.log 2>&1
  ls -la $OUT/${TAG}_blend.ncu-rep
  ncu -i $OUT/${TAG}_blend.ncu-rep --page raw --csv > $OUT/${TAG}_blend_raw.csv 2>/dev/null
fi
if has small; then
  timeout 900 ncu --set full --clock-control none --import-source on \
      -k regex:'(preprocess_fwd|bin_pass|tile_sort|sh_bwd|geom_bwd|tile_scan|column_scan)' -s 20 -c 8 \
      -f -o $OUT/${TAG}_small python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity > $OUT/${TAG}_small.log 2>&1
  ncu -i $OUT/${TAG}_small.ncu-rep --page raw --csv > $OUT/${TAG}_small_raw.csv 2>/dev/null
fi
if has sanitize; then
  # memcheck of the small golden cases through the whole fwd+bwd path (both backward kernels)
  timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py tests/test_gpu_view_parallel.py -m gpu -x -q \
      -k "golden and (tiny or ragged or negfov or rotcam or deg1m4) or (colour_only_backward_equals_zero_aux_gradients and small) or (factor_mode and small) or sh_outer_sum" > $OUT/${TAG}_memcheck.log 2>&1
  echo "memcheck exit $?" >> $OUT/${TAG}_memcheck.log
  grep -E "ERROR SUMMARY|passed|failed|memcheck exit" $OUT/${TAG}_memcheck.log | tail -5
fi
if has racecheck; then
  # shared-memory hazards of the forward / both backward kernels (the v2 backward shares per-warp slots across lanes)
  timeout 420 compute-sanitizer --tool racecheck --racecheck-report all --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -x -q \
      -k "golden and (tiny or ragged) or (colour_only_backward_equals_zero_aux_gradients and small)" > $OUT/${TAG}_racecheck.log 2>&1
  echo "racecheck exit $?" >> $OUT/${TAG}_racecheck.log
  grep -E "RACECHECK SUMMARY|hazard|passed|failed|racecheck exit" $OUT/${TAG}_racecheck.log | tail -8
fi
