# round-2 call AC (8 GPUs): config 3's table row-sharded instead of replicated ("replicas and sharded variant for comparison")
mkdir -p gpurun_out
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29618 bench.py --gpus 8 --workload config3s --steps 20 --warmup 5 > gpurun_out/bench_config3s_n8.json 2> gpurun_out/bench_config3s_n8.err; echo "rc=$?"; tail -2 gpurun_out/bench_config3s_n8.err | cut -c1-300
python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/bench_config3s_n8.json') if l.startswith('{')][-1])
s = d.get('sharded', d)
print(round(d['value']/1e6, 1), 'Mtok/s', round(d['ms_per_step'], 3), 'ms', d.get('clocks'))
for m in ('peer', 'nccl'):
    if m in s: print(m, round(s[m]['value']/1e6, 1), round(s[m]['ms_per_step'], 3), 'ms nvlink', round(s[m]['nvlink']['frac'], 3), 'hbm', round(s[m]['roofline']['frac'], 3), s[m]['clocks'])
print(s.get('parity'))
PY
