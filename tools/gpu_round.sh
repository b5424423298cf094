#!/bin/bash
# One GPU visit: parity tests, both bench arms, the ncu launch list and one `--set full` capture of
# the blend kernels.  Usage (from the repo root, under gpurun):  bash tools/gpu_round.sh <tag> [what]
#   what = any of: tests bench ref launches full   (default: all)
TAG=${1:-r01_x}
WHAT=${2:-"tests bench ref launches full"}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_smi.txt 2>&1
has() { [[ " $WHAT " == *" $1 "* ]]; }

if has tests; then
  timeout 900 python -m pytest tests -m gpu -x -q > $OUT/${TAG}_pytest.log 2>&1
  echo "pytest exit $?" >> $OUT/${TAG}_pytest.log
  tail -5 $OUT/${TAG}_pytest.log
fi
if has bench; then
  timeout 600 python bench.py --steps 20 --warmup 5 > $OUT/${TAG}_bench_ours.json 2> $OUT/${TAG}_bench_ours.err
  tail -c 3000 $OUT/${TAG}_bench_ours.json
fi
if has ref; then
  timeout 600 python bench.py --impl reference --steps 10 --warmup 3 > $OUT/${TAG}_bench_ref.json 2> $OUT/${TAG}_bench_ref.err
  tail -c 1200 $OUT/${TAG}_bench_ref.json
fi
if has launches; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
      --log-file $OUT/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_launches.log 2>&1
  python tools/launch_summary.py $OUT/${TAG}_launches.csv | tee $OUT/${TAG}_launch_summary.txt
fi
if has full; then
  # skip the warm-up launches of each kernel (-s counts matching launches only), capture one of each
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'blend_(fwd|bwd|bwd2)_kernel' -s 6 -c 2 \
      -f -o $OUT/${TAG}_blend python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $OUT/${TAG}_full.log 2>&1
  ls -la $OUT/${TAG}_blend.ncu-rep
fi
if has sanitize; then
  # memcheck of the small golden cases through the whole fwd+bwd path (both backward kernels)
  timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests -m gpu -x -q \
      -k "golden and (tiny or ragged or negfov) or colour_only_backward_equals_zero_aux_gradients and small" > $OUT/${TAG}_memcheck.log 2>&1
  echo "memcheck exit $?" >> $OUT/${TAG}_memcheck.log
  grep -E "ERROR SUMMARY|passed|failed|memcheck exit" $OUT/${TAG}_memcheck.log | tail -5
fi
