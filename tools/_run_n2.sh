set -x
export SIMCLR_B200_PEER_TIMEOUT_S=20
cd $GRAFT_REPO_ROOT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 300 python bench.py --no-extras --steps 30 --warmup 5 > gpurun_out/b1_win.json 2> gpurun_out/b1_win.err; echo "rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/b1_win.json'))
print(d['ms_per_step'], d['ms_per_step_best_of_5'], d['kernels_alone_ms'], d['roofline']['whole_step_frac'])
PY
timeout 400 $TR tests/distributed_check.py > gpurun_out/dc2_win.log 2>&1; echo "dc rc=$?"
grep -c "OK" gpurun_out/dc2_win.log; grep -v "OK" gpurun_out/dc2_win.log | tail -5
for i in 1 2; do
timeout 300 $TR bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/b2_win_$i.json 2> gpurun_out/b2_win_$i.err; echo "bench rc=$?"; cat gpurun_out/b2_win_$i.json | cut -c1-300
done
timeout 300 python bench.py --no-extras --steps 30 --warmup 5 > gpurun_out/b1_win2.json 2> gpurun_out/b1_win.err; echo "rc=$?"; python - <<'PY'
import json
d=json.load(open('gpurun_out/b1_win2.json'))
print(d['ms_per_step'], d['ms_per_step_best_of_5'], d['kernels_alone_ms'], d['roofline']['whole_step_frac'])
PY
