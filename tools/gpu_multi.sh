mkdir -p gpurun_out
N=${NGPU:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29555 tests/multi_gpu/run_sharded.py > gpurun_out/sharded_parity_n$N.log 2>&1; echo "sharded parity rc=$?"; grep -E "rank 0|Error|error|Traceback" gpurun_out/sharded_parity_n$N.log | tail -12
for MODE in peer; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29556 bench.py --gpus $N --workload config4 --steps 10 --warmup 3 --rows-per-gpu ${ROWS:-2500000} --sharded-mode $MODE > gpurun_out/bench_config4_${MODE}_n$N.json 2> gpurun_out/bench_config4_${MODE}_n$N.err; echo "config4 $MODE rc=$?"; grep metric gpurun_out/bench_config4_${MODE}_n$N.json | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print(d['config']['sharded_mode'], 'Mtok/s', round(d['value']/1e6,1), 'ms/step', round(d['ms_per_step'],3), 'nvlink GB/s/GPU', round(d['nvlink']['achieved_in_GBps_per_gpu'],1), 'hbm frac', round(d['roofline']['frac'],3))"; tail -2 gpurun_out/bench_config4_${MODE}_n$N.err | cut -c1-300
done
