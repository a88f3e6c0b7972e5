export SIMCLR_B200_PEER_TIMEOUT_S=20
cd $GRAFT_REPO_ROOT
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 400 $TR tests/distributed_check.py > gpurun_out/dc2_win.log 2>&1; echo "dc rc=$?"
grep -c "OK" gpurun_out/dc2_win.log; grep -v "OK" gpurun_out/dc2_win.log | tail -5
for i in 1 2; do
for m in 2 1 0; do
SIMCLR_B200_PEER_WINDOWS=$m timeout 300 $TR bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/b2_m${m}_$i.json 2> gpurun_out/b2_m${m}_$i.err; echo "mode $m rc=$?"; python - <<PY
import json
d=json.load(open('gpurun_out/b2_m${m}_$i.json'))
print(d['ms_per_step'], d.get('parity'), d.get('strong_scaling'))
PY
done
done
