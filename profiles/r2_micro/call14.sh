#!/bin/bash
# tcgen05 probe: descriptor variants and batch timing (every test in its own process, bounded)
mkdir -p gpurun_out/c14
cd profiles/r2_micro
for t in 1 2 3 4; do for v in 0 1; do
  timeout 60 ./umma_probe $t $v > ../../gpurun_out/c14/probe_${t}_${v}.txt 2>&1; echo "test $t variant $v rc=$?"
  tail -2 ../../gpurun_out/c14/probe_${t}_${v}.txt
done; done
timeout 120 ./umma_probe 5 > ../../gpurun_out/c14/probe_5.txt 2>&1; echo "test 5 rc=$?"
cat ../../gpurun_out/c14/probe_5.txt
