#!/bin/bash
mkdir -p gpurun_out/c15
cd profiles/r2_micro
for t in 6 7 1; do
  timeout 60 ./umma_probe $t 0 > ../../gpurun_out/c15/probe_${t}.txt 2>&1; echo "test $t rc=$?"; tail -3 ../../gpurun_out/c15/probe_${t}.txt
done
timeout 120 compute-sanitizer --tool memcheck ./umma_probe 1 0 > ../../gpurun_out/c15/sanitizer_1.txt 2>&1; echo "sanitizer rc=$?"
head -40 ../../gpurun_out/c15/sanitizer_1.txt
