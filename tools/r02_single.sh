#!/usr/bin/env bash
# one single-GPU measurement pass of round 2 (run on the GPU box through gpurun): tools/r02_single.sh TAG
set -u
tag=${1:-r02x}
out=gpurun_out
mkdir -p "$out"
true
python -m pytest tests -m gpu -q > "$out/${tag}_gpu_tests.log" 2>&1
tail -8 "$out/${tag}_gpu_tests.log"
python bench.py --steps 10 --no-cpu-baseline > "$out/${tag}_bench_fsi512.json" 2> "$out/${tag}_bench_fsi512.err"
python -c "
import json; d=json.load(open('$out/${tag}_bench_fsi512.json')); print(d['value'], d['ms_per_step'], d['config']['stage_ms'], d['e2e'], d['roofline']['frac'], d['config']['step_hbm_frac_of_measured'])"
for w in rod_fsi_512x256x256_f32 sphere_vbf_512x256x256_f32 vortex_ring_256_f32; do
  python bench.py --workload $w --no-cpu-baseline 2>/dev/null > "$out/${tag}_bench_$w.json"
  python -c "
import json; d=json.load(open('$out/${tag}_bench_$w.json')); print('$w', d['value'], d['ms_per_step'], d['config']['step_hbm_frac_of_measured'], d['roofline']['frac'])"
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file "$out/${tag}_launches_fsi512.csv" \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline > /dev/null 2>&1
python tools/agg_launches.py "$out/${tag}_launches_fsi512.csv" 18 | grep -v Greens
ncu --set full --import-source on --clock-control none -k regex:"zconvw|fused_v2|strided32_kernel|velocity_vec4|x_r2c|x_c2r|ib_" \
  -s 40 -c 14 -o "$out/${tag}_prof_fsi512" -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ls -la "$out/${tag}_prof_fsi512.ncu-rep"
