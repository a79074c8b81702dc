# Round-2 profiling pass (run on the B200 box through gpurun; artefacts come back in gpurun_out/r2, then
#   python profiles/summarize.py gpurun_out/r2 r2   turns them into profiles/r2_*.{json,md,csv}).
set -x
mkdir -p gpurun_out/r2
rm -f gpurun_out/r2/*
python bench.py > gpurun_out/r2/bench_ours_n1.json 2> gpurun_out/r2/bench_ours_n1.err
python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/r2/bench_reference_n1.json 2> gpurun_out/r2/bench_reference_n1.err
python bench.py --workload C2 --no-cpu-baseline > gpurun_out/r2/bench_ours_C2.json 2>/dev/null
python bench.py --workload C2 --impl reference --steps 10 --warmup 3 > gpurun_out/r2/bench_reference_C2.json 2>/dev/null
python bench.py --workload C5 --no-cpu-baseline --steps 50 > gpurun_out/r2/bench_ours_C5.json 2>/dev/null
python bench.py --colour 27 --no-cpu-baseline --steps 30 > gpurun_out/r2/bench_ours_C3_sh27.json 2>/dev/null
python bench.py --colour 48 --no-cpu-baseline --steps 30 > gpurun_out/r2/bench_ours_C3_sh48.json 2>/dev/null
python bench.py --colour 27 --impl reference --steps 5 --warmup 2 > gpurun_out/r2/bench_reference_C3_sh27.json 2>/dev/null
# every launch of two frames with its device time (cold-cache, serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extra-legs > gpurun_out/r2/ncu_launches.log 2>&1
# full sets: all six kernels of OUR frame (second frame), RGB and SH-27
ncu --set full --clock-control none --import-source on -k regex:"blend_|fused_project|emit_keys|tile_ranges" -s 6 -c 6 -o gpurun_out/r2/prof_C3_D3 -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extra-legs > gpurun_out/r2/ncu_full_d3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"blend_sh|fused_project_bwd" -s 3 -c 3 -o gpurun_out/r2/prof_C3_D27 -f python bench.py --colour 27 --steps 2 --warmup 1 --no-cpu-baseline --no-extra-legs > gpurun_out/r2/ncu_full_d27.log 2>&1
ls -la gpurun_out/r2
