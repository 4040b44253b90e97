#!/bin/bash
# Profiling pass of a round on one B200 (run through gpurun): knob check, ncu --set full of the PCG kernels, launch list of one
# bench step, driver sweeps + table comparison.  Outputs under gpurun_out/ with the given prefix.
p=${1:-r02}
o=gpurun_out
python scripts/gpu_spmm_sweep.py L "SPMM_WINDOW=1" > $o/${p}_spmm_check.txt 2>&1
ncu --set full --clock-control none --import-source on --kernel-name regex:"spmm_window_kernel|cg_pupdate_coarse|restrict_cell|cg_update_kernel|coarse_node_kernel" --launch-skip 100 --launch-count 10 -o $o/${p}_pcg8_window_L -f python scripts/gpu_spmm_probe.py L 40 > $o/${p}_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name regex:"spmv_stream_kernel" --launch-skip 10 --launch-count 2 -o $o/${p}_spmv1_L -f python scripts/gpu_spmv.py L 2 3 > $o/${p}_ncu2.log 2>&1
python scripts/ncu_summary.py $o/${p}_pcg8_window_L.ncu-rep > $o/${p}_ncu_full_pcg8_window_L.txt 2>&1
python scripts/ncu_summary.py $o/${p}_spmv1_L.ncu-rep > $o/${p}_ncu_full_spmv1_L.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $o/${p}_launches_bench_L.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e > $o/${p}_launches.log 2>&1
python scripts/launch_agg.py $o/${p}_launches_bench_L.csv --timed-step > $o/${p}_launches_bench_L_agg.txt 2>&1
bash scripts/run_all_sweeps.sh > $o/${p}_sweeps.log 2>&1
python scripts/compare_tables.py $o/sweeps > $o/sweeps/compare_vs_reference.txt 2>&1
cat $o/${p}_spmm_check.txt $o/${p}_ncu_full_pcg8_window_L.txt $o/${p}_ncu_full_spmv1_L.txt; head -12 $o/${p}_launches_bench_L_agg.txt; cat $o/sweeps/compare_vs_reference.txt
