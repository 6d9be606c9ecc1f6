# round-2 call N (2 or 8 GPUs): the NCCL all-to-all variant of the sharded tier, software-pipelined over micro-batches
mkdir -p gpurun_out
N=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29641 tests/multi_gpu/run_sharded.py > gpurun_out/sharded_parity_n2_r02.log 2>&1; echo "run_sharded rc=$?"; grep -c OK gpurun_out/sharded_parity_n2_r02.log; grep -E "MISMATCH|Error" gpurun_out/sharded_parity_n2_r02.log | head -5
for M in 1 2 4 8 16; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2965$M bench.py --gpus $N --workload config4 --sharded-mode nccl --nccl-micro $M --steps 10 --warmup 3 > gpurun_out/bench_config4_nccl_m$M.json 2> gpurun_out/bench_config4_nccl_m$M.err; python - $M <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(f'gpurun_out/bench_config4_nccl_m{sys.argv[1]}.json') if l.startswith('{')][-1])
    s = d['sharded']
    print('micro', sys.argv[1], round(s['nccl']['value']/1e6,1), 'Mtok/s', round(s['nccl']['ms_per_step'],3), 'ms nvlink', round(s['nccl']['nvlink']['frac'],3), s.get('parity'))
except Exception as e:
    print('micro', sys.argv[1], 'failed', repr(e))
PY
done
