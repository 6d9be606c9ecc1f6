mkdir -p gpurun_out
N=8
for MODE in peer nccl; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29556 bench.py --gpus $N --workload config3s --steps 20 --warmup 3 --sharded-mode $MODE > gpurun_out/bench_config3s_${MODE}_n$N.json 2> gpurun_out/bench_config3s_${MODE}_n$N.err; echo "config3s $MODE rc=$?"; grep metric gpurun_out/bench_config3s_${MODE}_n$N.json | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print(d['config']['sharded_mode'], 'f_grams', d['config']['f_grams'], 'Mtok/s', round(d['value']/1e6,1), 'per GPU', round(d['value']/1e6/d['n_gpus'],1), 'ms/step', round(d['ms_per_step'],3), 'nvlink GB/s/GPU', round(d['nvlink']['achieved_in_GBps_per_gpu'],1), 'hbm frac', round(d['roofline']['frac'],3))"; tail -2 gpurun_out/bench_config3s_${MODE}_n$N.err | cut -c1-200
done
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29558 bench.py --gpus $N --workload config3 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_config3_n$N.json 2> gpurun_out/bench_config3_n$N.err; echo "config3 x8 rc=$?"; grep metric gpurun_out/bench_config3_n$N.json | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print('config3 replicas x8', round(d['value']/1e6,1), d['ms_per_step'], d['roofline']['frac'], round(d['e2e']['value']/1e6,1))"
