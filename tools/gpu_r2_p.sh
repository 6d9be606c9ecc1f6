# round-2 call P (8 GPUs): NCCL variant of config 4 in full, micro-batch sweep
mkdir -p gpurun_out
N=8
for M in 1 4 8; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2968$M bench.py --gpus $N --workload config4 --sharded-mode nccl --nccl-micro $M --steps 10 --warmup 3 > gpurun_out/bench_config4_nccl_n8_m$M.json 2> gpurun_out/bench_config4_nccl_n8_m$M.err; python - $M <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(f'gpurun_out/bench_config4_nccl_n8_m{sys.argv[1]}.json') if l.startswith('{')][-1])
    s = d['sharded']
    print('micro', sys.argv[1], round(s['nccl']['value']/1e6,1), 'Mtok/s', round(s['nccl']['ms_per_step'],3), 'ms nvlink', round(s['nccl']['nvlink']['frac'],3), s.get('parity'))
except Exception as e:
    print('micro', sys.argv[1], 'failed', repr(e))
PY
done
