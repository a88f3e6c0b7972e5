# GPU box: the -m gpu tests, the default bench line, the reference arm and smoke() (outputs under gpurun_out/r02b_*).
#   gpurun --timeout 2400 -- bash tools/run_final_n1.sh
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/ -x -q -m gpu > gpurun_out/r02b_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02b_pytest_gpu.log
timeout 600 python bench.py > gpurun_out/r02b_bench_n1.json 2> gpurun_out/r02b_bench_n1.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference > gpurun_out/r02b_bench_ref_n1.json 2> gpurun_out/r02b_bench_ref_n1.err; echo "ref rc=$?"; cut -c1-400 gpurun_out/r02b_bench_ref_n1.json
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02b_bench_n1.json'))
print(d['ms_per_step'], d['ms_per_step_best_of_5'], d['kernels_alone_ms'], d['roofline']['whole_step_frac'], d['roofline']['frac'], d['e2e']['ms_per_step'])
PY
