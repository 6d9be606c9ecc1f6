# round-2 call AA (2 GPUs): NCCL variant of the sharded tier against NCCL's point-to-point channel count
mkdir -p gpurun_out
run() {  # tag, env...
  tag=$1; shift
  env "$@" timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29641 bench.py --gpus 2 --workload config4 --sharded-mode nccl --steps 10 --warmup 3 > gpurun_out/nccl_$tag.json 2> gpurun_out/nccl_$tag.err
  python - $tag <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(f'gpurun_out/nccl_{sys.argv[1]}.json') if l.startswith('{')][-1])
    s = d.get('sharded', d)
    n = s.get('nccl', s)
    print(sys.argv[1], round(d['value']/1e6, 1), 'Mtok/s', round(d['ms_per_step'], 3), 'ms', json.dumps(n.get('nvlink', {}))[:120])
except Exception as e:
    print(sys.argv[1], 'parse failed', repr(e))
PY
}
run default A=1
run p2p32 NCCL_MIN_P2P_NCHANNELS=32 NCCL_MAX_P2P_NCHANNELS=32
run p2p64 NCCL_MIN_P2P_NCHANNELS=64 NCCL_MAX_P2P_NCHANNELS=64 NCCL_MAX_NCHANNELS=64
run ctas NCCL_MIN_CTAS=32 NCCL_MAX_CTAS=64
