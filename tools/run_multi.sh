# GPU box, N = $1 GPUs: tests/distributed_check.py and the bench line (short cross-GPU watchdog).
#   gpurun --gpus N --timeout 1200 -- bash tools/run_multi.sh N
export SIMCLR_B200_PEER_TIMEOUT_S=20
cd $GRAFT_REPO_ROOT
N=$1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 400 $TR tests/distributed_check.py > gpurun_out/r02b_dist_check_n$N.log 2>&1; echo "dc rc=$?"
grep -c "OK" gpurun_out/r02b_dist_check_n$N.log; grep -v "OK" gpurun_out/r02b_dist_check_n$N.log | tail -3
timeout 300 $TR bench.py --gpus $N > gpurun_out/r02b_bench_n$N.json 2> gpurun_out/r02b_bench_n$N.err; echo "bench rc=$?"; python - <<PY
import json
d=json.load(open('gpurun_out/r02b_bench_n$N.json'))
print(d['ms_per_step'], d.get('ms_per_step_back_to_back'), d.get('parity',{}).get('ok'), {k:v for k,v in d.get('strong_scaling',{}).items() if k!='base'}, d.get('e2e'))
PY
