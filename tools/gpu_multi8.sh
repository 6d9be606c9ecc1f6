mkdir -p gpurun_out
N=8
nvidia-smi --query-gpu=index,name,memory.total --format=csv,noheader | head -8
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 tests/multi_gpu/run_sharded.py > gpurun_out/sharded_parity_n$N.log 2>&1; echo "sharded parity rc=$?"; grep -E "rank 0|Error|error" gpurun_out/sharded_parity_n$N.log | tail -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29556 bench.py --gpus $N --workload config4 --steps 10 --warmup 3 --sharded-mode peer > gpurun_out/bench_config4_peer_n$N.json 2> gpurun_out/bench_config4_peer_n$N.err; echo "config4 peer full rc=$?"; grep metric gpurun_out/bench_config4_peer_n$N.json | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print(d['config']['sharded_mode'], 'f_grams', d['config']['f_grams'], 'Mtok/s', round(d['value']/1e6,1), 'per GPU', round(d['value']/1e6/d['n_gpus'],1), 'ms/step', round(d['ms_per_step'],3), 'nvlink GB/s/GPU', round(d['nvlink']['achieved_in_GBps_per_gpu'],1), 'frac', round(d['nvlink']['frac'],3), 'hbm frac', round(d['roofline']['frac'],3))"; tail -2 gpurun_out/bench_config4_peer_n$N.err | cut -c1-300
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29557 bench.py --gpus $N --workload config4 --steps 10 --warmup 3 --sharded-mode nccl --rows-per-gpu 2500000 > gpurun_out/bench_config4_nccl_n$N.json 2> gpurun_out/bench_config4_nccl_n$N.err; echo "config4 nccl rc=$?"; grep metric gpurun_out/bench_config4_nccl_n$N.json | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print(d['config']['sharded_mode'], 'f_grams', d['config']['f_grams'], 'Mtok/s', round(d['value']/1e6,1), 'per GPU', round(d['value']/1e6/d['n_gpus'],1), 'ms/step', round(d['ms_per_step'],3), 'nvlink GB/s/GPU', round(d['nvlink']['achieved_in_GBps_per_gpu'],1))"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29558 bench.py --gpus $N --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_config2_n$N.json 2> gpurun_out/bench_config2_n$N.err; echo "config2 x8 rc=$?"; grep metric gpurun_out/bench_config2_n$N.json | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print('config2 x8', round(d['value']/1e6,1), d['ms_per_step'], d['roofline']['frac'], round(d['e2e']['value']/1e6,1))"
