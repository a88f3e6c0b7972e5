# GPU box: ncu launch list + ncu --set full captures of tools/profile_step.py (each after a plain run of the same command),
# exported as CSV (the .ncu-rep files stay in /tmp: gpurun_out is limited to 64 MiB); then, here:
#   python tools/ncu_summary.py gpurun_out/r02b_ncu_full_tile_kernels.csv gpurun_out/r02b_ncu_full_rowwise_kernels.csv --out profiles/r02b_ncu_summary.md
#   gpurun --timeout 1500 -- bash tools/run_ncu.sh
cd $GRAFT_REPO_ROOT
CMD="python tools/profile_step.py --steps 3"
$CMD > gpurun_out/r02b_profile_step_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02b_ncu_launch_list.csv $CMD > gpurun_out/ncu1.log 2>&1
echo "launch list rc=$?"
$CMD > gpurun_out/r02b_profile_step_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:contrastive_tile_kernel -s 6 -c 2 -o /tmp/tile $CMD > gpurun_out/ncu2.log 2>&1
echo "tile rc=$?"
ncu -i /tmp/tile.ncu-rep --page raw --csv > gpurun_out/r02b_ncu_full_tile_kernels.csv 2>/dev/null
ncu -i /tmp/tile.ncu-rep --page source --csv > /tmp/tile_source.csv 2>/dev/null; head -c 3000000 /tmp/tile_source.csv > gpurun_out/r02b_ncu_source_tile_kernels.csv
$CMD > gpurun_out/r02b_profile_step_plain.log 2>&1 &&
ncu --set full --clock-control none -k regex:"prepare_kernel|finalize_kernel" -s 9 -c 3 -o /tmp/row $CMD > gpurun_out/ncu3.log 2>&1
echo "row rc=$?"
ncu -i /tmp/row.ncu-rep --page raw --csv > gpurun_out/r02b_ncu_full_rowwise_kernels.csv 2>/dev/null
ls -la gpurun_out/r02b_ncu* ; cat gpurun_out/r02b_profile_step_plain.log
