#!/bin/bash
# refresh the SH artefacts of the round-2 profile pass (same commands as profiles/r2_profile_commands.sh)
mkdir -p gpurun_out/r2
python bench.py --colour 27 --no-cpu-baseline --steps 30 > gpurun_out/r2/bench_ours_C3_sh27.json 2>/dev/null; echo "bench27 rc=$?"
python bench.py --colour 48 --no-cpu-baseline --steps 30 > gpurun_out/r2/bench_ours_C3_sh48.json 2>/dev/null; echo "bench48 rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"blend_sh|fused_project_bwd" -s 3 -c 3 -o gpurun_out/r2/prof_C3_D27 -f python bench.py --colour 27 --steps 2 --warmup 1 --no-cpu-baseline --no-extra-legs > gpurun_out/r2/ncu_full_d27.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import json
for d in (27,48):
    b=json.loads(open(f'gpurun_out/r2/bench_ours_C3_sh{d}.json').read().strip().splitlines()[-1])
    print(d, b['value'], b['ms_per_step'], 'e2e', b['e2e']['value'], b.get('stage_ms'))
PY
