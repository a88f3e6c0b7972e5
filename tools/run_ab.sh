# GPU box: same-box A/B of two builds of the library (lib/libsimclr_b200_prev.so against the current one), twice each.
#   gpurun --timeout 900 -- bash tools/run_ab.sh
cd $GRAFT_REPO_ROOT
for i in 1 2; do
for v in prev cur; do
if [ $v = prev ]; then export SIMCLR_B200_LIB=$PWD/pytorch-simclr_b200/lib/libsimclr_b200_prev.so; else unset SIMCLR_B200_LIB; fi
timeout 300 python bench.py --no-extras > gpurun_out/ab_$v$i.json 2> gpurun_out/ab_$v$i.err; echo "$v rc=$?"; python - <<PY
import json
d=json.load(open('gpurun_out/ab_$v$i.json'))
print('$v', d['ms_per_step'], d['ms_per_step_best_of_5'], d['kernels_alone_ms'], d['roofline']['whole_step_frac'], d['e2e']['ms_per_step'], d.get('deterministic_mode',{}).get('ms_per_step'))
PY
done
done
