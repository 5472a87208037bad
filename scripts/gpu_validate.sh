#!/bin/bash
# The round's validation pass on a B200 box (run through gpurun): GPU parity tests, smoke, the bench line with its CPU
# baseline, the reference arm, the ncu launch list of the bench command, the per-config table and the ncu --set full
# captures the profiles/ summaries come from.
#   /usr/local/graft/bin/gpurun --timeout 3000 -- 'bash scripts/gpu_validate.sh [tag]'
set -x
tag=${1:-validate}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8 | tee gpurun_out/${tag}_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/${tag}_smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 2>gpurun_out/${tag}_bench.err | tee gpurun_out/${tag}_bench.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 2>>gpurun_out/${tag}_bench.err | tee gpurun_out/${tag}_bench_ref.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
  python bench.py --steps 2 --warmup 3 --simple-data --no-cpu-baseline > gpurun_out/${tag}_ncu_bench.log 2>&1
timeout 1500 python scripts/bench_configs.py c1 c2 c3 c4 c5 2>&1 | tee gpurun_out/${tag}_configs.jsonl
timeout 600 python scripts/strong_probe.py 2>&1 | tee gpurun_out/${tag}_strong_probe.jsonl
# ncu --set full captures (one launch each)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ekf_small_lw -s 3 -c 1 -o gpurun_out/prof_${tag}_ekf_small_lw \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --simple-data > gpurun_out/prof_${tag}_ekf_small_lw.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:kf_warp_filter -s 1 -c 1 -o gpurun_out/prof_${tag}_kf_warp_filter \
  python scripts/profile_generic.py kf > gpurun_out/prof_${tag}_kf_warp_filter.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:generic_filter -s 1 -c 1 -o gpurun_out/prof_${tag}_ukf_closed \
  python scripts/profile_generic.py ukf > gpurun_out/prof_${tag}_ukf_closed.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:enkf_kernel -s 1 -c 1 -o gpurun_out/prof_${tag}_enkf \
  python scripts/profile_generic.py enkf > gpurun_out/prof_${tag}_enkf.log 2>&1
ls -la gpurun_out/*${tag}*.ncu-rep
