#!/bin/bash
mkdir -p gpurun_out/c16
cd profiles/r2_micro
for b in umma_probe umma_probe_elect; do for t in 6 1; do
  timeout 60 ./$b $t 0 > ../../gpurun_out/c16/${b}_${t}.txt 2>&1; echo "$b test $t rc=$?"; tail -6 ../../gpurun_out/c16/${b}_${t}.txt
done; done
for t in 2 3 4; do for v in 0 1; do
  timeout 60 ./umma_probe_elect $t $v > ../../gpurun_out/c16/elect_${t}_${v}.txt 2>&1; echo "elect test $t variant $v rc=$?"; tail -4 ../../gpurun_out/c16/elect_${t}_${v}.txt
done; done
