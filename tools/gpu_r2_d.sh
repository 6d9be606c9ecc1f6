# round-2 call D (2 GPUs): full GPU tests incl. the 2-GPU sharded test, bench suite at N = 2 (reduced, then full config-4 shard size)
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name,memory.total --format=csv,noheader
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 20 --warmup 5 --rows-per-gpu 2000000 > gpurun_out/bench_suite_n2_small.json 2> gpurun_out/bench_suite_n2_small.err; echo "bench n2 small rc=$?"; tail -3 gpurun_out/bench_suite_n2_small.err | cut -c1-300
python - <<'PY'
import json
for f in ('gpurun_out/bench_suite_n2_small.json',):
    try:
        d = json.loads([l for l in open(f) if l.startswith('{')][-1])
        print('config2 x2', round(d['value']/1e6,1), 'Mtok/s', round(d['ms_per_step']*1e3,2), 'us', round(d['roofline']['frac'],3), d['clocks'])
        s = d.get('sharded', {})
        print({k: v for k, v in s.items() if k not in ('peer', 'nccl', 'parity_detail')})
        for m in ('peer', 'nccl'):
            if m in s: print(m, round(s[m]['value']/1e6,1), 'Mtok/s', round(s[m]['ms_per_step'],3), 'ms nvlink', round(s[m]['nvlink']['frac'],3), 'hbm', round(s[m]['roofline']['frac'],3), s[m]['clocks'])
        print(s.get('parity'), s.get('parity_detail'))
    except Exception as e:
        print('parse failed', f, repr(e))
PY
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_suite_n2.json 2> gpurun_out/bench_suite_n2.err; echo "bench n2 full rc=$?"; tail -3 gpurun_out/bench_suite_n2.err | cut -c1-300
python - <<'PY'
import json
for f in ('gpurun_out/bench_suite_n2.json',):
    try:
        d = json.loads([l for l in open(f) if l.startswith('{')][-1])
        print('config2 x2', round(d['value']/1e6,1), 'Mtok/s', round(d['ms_per_step']*1e3,2), 'us', round(d['roofline']['frac'],3), d['clocks'])
        s = d.get('sharded', {})
        print({k: v for k, v in s.items() if k not in ('peer', 'nccl', 'parity_detail')})
        for m in ('peer', 'nccl'):
            if m in s: print(m, round(s[m]['value']/1e6,1), 'Mtok/s', round(s[m]['ms_per_step'],3), 'ms nvlink', round(s[m]['nvlink']['frac'],3), 'hbm', round(s[m]['roofline']['frac'],3), s[m]['clocks'])
        print(s.get('parity'), s.get('parity_detail'))
    except Exception as e:
        print('parse failed', f, repr(e))
PY
