#!/bin/bash
# timings of the pipelined window SpMM prototype on the GPU box (built here, binary travels)
cd scripts/proto
for cfg in "16 4 2 2 2 1" "16 4 2 3 2 1" "16 4 2 2 3 1" "8 4 4 2 2 1" "8 4 4 3 2 1" "32 2 2 2 2 1" "16 4 2 2 2 4" "16 2 2 4 2 1" "16 2 2 4 3 1" "16 2 2 5 3 1" "16 4 4 1 2 1"; do
  timeout 120 ./spmm_window2 $cfg
done
