#!/bin/bash
# Full step01..step04 runs with the GPU engine; tables are copied to gpurun_out/sweeps/.
set -u
out=$PWD/gpurun_out/sweeps; mkdir -p $out
t0=$(date +%s.%N)
(mkdir -p /tmp/s1 && cd /tmp/s1 && rm -rf * && python $OLDPWD/drivers/step01_box/test_step01_baseline.py) > $out/step01.log 2>&1
t1=$(date +%s.%N)
(cd drivers/step02_electrodes && python run_sweep.py) > $out/step02.log 2>&1; cp drivers/step02_electrodes/results/summary.csv $out/step02_summary.csv
t2=$(date +%s.%N)
(cd drivers/step03_ankle_layers && python run_layered_sweep.py) > $out/step03.log 2>&1; cp drivers/step03_ankle_layers/results/summary.csv $out/step03_summary.csv; cp drivers/step03_ankle_layers/results/summary.json $out/step03_summary.json
t3=$(date +%s.%N)
(cd drivers/step04_pressure && python run_pressure_sweep.py) > $out/step04.log 2>&1; cp drivers/step04_pressure/results/summary.csv $out/step04_summary.csv; cp drivers/step04_pressure/results/summary.json $out/step04_summary.json
t4=$(date +%s.%N)
(cd drivers/step04_pressure && python run_pressure_sweep.py --sequential) > $out/step04_seq.log 2>&1
t5=$(date +%s.%N)
python - <<PY > $out/timing.txt
print("wall seconds incl. python start, meshing, file I/O: step01 %.1f  step02(8 cases) %.1f  step03(9 cases) %.1f  step04(15 levels, batched) %.1f  step04(sequential) %.1f" % ($t1-$t0, $t2-$t1, $t3-$t2, $t4-$t3, $t5-$t4))
PY
cat $out/timing.txt; tail -3 $out/step01.log; tail -4 $out/step04.log
