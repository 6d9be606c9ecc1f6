mkdir -p gpurun_out
N=${NGPU:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus $N --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference_n$N.json 2> gpurun_out/bench_reference_n$N.err; echo "reference arm under torchrun rc=$?"; grep -c impl gpurun_out/bench_reference_n$N.json; grep impl gpurun_out/bench_reference_n$N.json | cut -c1-200
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29562 bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/bench_config2_n$N.json 2> gpurun_out/bench_config2_n$N.err; echo "config2 x$N rc=$?"; grep metric gpurun_out/bench_config2_n$N.json | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print('config2 x', d['n_gpus'], round(d['value']/1e6,1), d['ms_per_step'], round(d['roofline']['frac'],3), round(d['e2e']['value']/1e6,1), 'cpu_baseline' in d)"
