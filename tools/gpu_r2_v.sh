# round-2 call V (1 GPU): position rows staged as one block per tile -- GPU tests, then every mode on every shape
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
M="replace pos add addpos"
for spec in config2 config3 config1 custom:384:fp16:3:100000:64:1024:50257 custom:768:fp16:3:100000:64:1024:50257 custom:768:fp32:3:100000:64:1024:50257 custom:1024:fp32:3:1000000:64:1024:50257 custom:1280:fp32:3:1000000:64:1024:50257; do
  args=""
  for m in $M; do args="$args $m; $m;STABLE=1"; done
  f=gpurun_out/modes11_$(echo $spec | tr ':' '_').log
  timeout 600 python tools/tune_modes.py $spec $args > $f 2>&1; echo "$spec rc=$?"
  python - $f <<'PY'
import json, sys
rows = [json.loads(l) for l in open(sys.argv[1]) if l.startswith('{"workload')]
print("  ".join(f"{x['mode']}{'*' if x['env'] else ''} {x['us']:.1f}" for x in rows))
PY
done
