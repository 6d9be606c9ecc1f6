# round-2 call F (8 GPUs): the driver's scaling command at N = 8 and N = 4 -- config 2 replicas + config 4 row-sharded (819 GB at N = 8)
mkdir -p gpurun_out
N=${1:-8}
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
for n in $N; do
timeout 860 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2961$n bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/bench_suite_n$n.json 2> gpurun_out/bench_suite_n$n.err; echo "bench n$n rc=$?"; tail -2 gpurun_out/bench_suite_n$n.err | cut -c1-300
python - $n <<'PY'
import json, sys
n = sys.argv[1]
f = f'gpurun_out/bench_suite_n{n}.json'
try:
    d = json.loads([l for l in open(f) if l.startswith('{')][-1])
    print('config2 x' + n, round(d['value']/1e6,1), 'Mtok/s', round(d['ms_per_step']*1e3,2), 'us', round(d['roofline']['frac'],3), d['clocks'], 'e2e', round(d['e2e']['value']/1e6,1))
    s = d.get('sharded', {})
    print({k: v for k, v in s.items() if k not in ('peer', 'nccl', 'parity_detail', 'workload', 'partitioning', 'timing')})
    for m in ('peer', 'nccl'):
        if m in s: print(m, round(s[m]['value']/1e6,1), 'Mtok/s', round(s[m]['tokens_per_s_per_gpu']/1e6,1), 'per GPU', round(s[m]['ms_per_step'],3), 'ms nvlink', round(s[m]['nvlink']['frac'],3), 'hbm', round(s[m]['roofline']['frac'],3), s[m]['clocks'])
    print(s.get('parity'), (s.get('parity_detail') or {}).get('seconds'))
except Exception as e:
    print('parse failed', f, repr(e))
PY
done
