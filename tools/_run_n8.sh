export SIMCLR_B200_PEER_TIMEOUT_S=20
cd $GRAFT_REPO_ROOT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511"
timeout 400 $TR tests/distributed_check.py > gpurun_out/r02b_dist_check_n8.log 2>&1; echo "dc rc=$?"
grep -c "OK" gpurun_out/r02b_dist_check_n8.log; grep -v "OK" gpurun_out/r02b_dist_check_n8.log | tail -5
for m in 2 0 1 2 0; do
SIMCLR_B200_PEER_WINDOWS=$m timeout 300 $TR bench.py --gpus 8 --steps 30 --warmup 5 > gpurun_out/b8_m${m}.json 2> gpurun_out/b8_m${m}.err; echo "mode $m rc=$?"; cp gpurun_out/b8_m${m}.json gpurun_out/b8_m${m}_$(date +%s).json; python - <<PY
import json
d=json.load(open('gpurun_out/b8_m${m}.json'))
print(d['ms_per_step'], d.get('ms_per_step_back_to_back'), d.get('parity',{}).get('ok'), {k:v for k,v in d.get('strong_scaling',{}).items() if k!='base'})
PY
done
