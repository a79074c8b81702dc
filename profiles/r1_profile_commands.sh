set -x
mkdir -p gpurun_out/r1
rm -f gpurun_out/r1/*
python bench.py > gpurun_out/r1/bench_ours_n1.json 2> gpurun_out/r1/bench_ours_n1.err
python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/r1/bench_reference_n1.json 2> gpurun_out/r1/bench_reference_n1.err
python bench.py --workload C2 --no-cpu-baseline > gpurun_out/r1/bench_ours_C2.json 2>/dev/null
python bench.py --workload C2 --impl reference --steps 10 --warmup 3 > gpurun_out/r1/bench_reference_C2.json 2>/dev/null
python bench.py --workload C5 --no-cpu-baseline --steps 50 > gpurun_out/r1/bench_ours_C5.json 2>/dev/null
python bench.py --colour 27 --no-cpu-baseline --steps 30 > gpurun_out/r1/bench_ours_C3_sh27.json 2>/dev/null
python bench.py --colour 48 --no-cpu-baseline --steps 30 > gpurun_out/r1/bench_ours_C3_sh48.json 2>/dev/null
python bench.py --colour 27 --impl reference --steps 5 --warmup 2 > gpurun_out/r1/bench_reference_C3_sh27.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r1/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"blend_fwd|blend_bwd" -s 4 -c 2 -o gpurun_out/r1/prof_blend -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r1/ncu_full.log 2>&1
ls -la gpurun_out/r1
