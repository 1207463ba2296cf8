#!/usr/bin/env bash
# Standard single-GPU measurement pass of a round (run on the GPU box, e.g. through gpurun):
#   tools/measure_round.sh r02
# writes gpurun_out/<tag>_*: GPU test log, bench lines (256^3 default workload, 512^3, reference arm),
# the ncu launch list of the bench command and one `ncu --set full` capture of a step's kernels.
# Copy what should be judged into profiles/ (tools/agg_launches.py, tools/ncu_summary.py summarise).
set -u
tag=${1:-rXX}
out=gpurun_out
mkdir -p "$out"
python -m pytest tests -m gpu -x -q > "$out/${tag}_gpu_tests.log" 2>&1
tail -2 "$out/${tag}_gpu_tests.log"
python bench.py > "$out/${tag}_bench_256.json" 2> "$out/${tag}_bench_256.err"
python bench.py --impl reference --steps 3 --warmup 1 > "$out/${tag}_bench_reference.json" 2> /dev/null
python bench.py --workload vortex_ring_512_f32 --steps 10 --no-cpu-baseline > "$out/${tag}_bench_512.json" 2> /dev/null
# (numbers printed under ncu are never bench values)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file "$out/${tag}_launches_step256.csv" \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --import-source on --clock-control none -k regex:sb_ -s 60 -c 9 -o "$out/${tag}_prof_step256" -f \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
python tools/agg_launches.py "$out/${tag}_launches_step256.csv" 12
cut -c1-240 "$out/${tag}_bench_256.json"
