# round-2 call U (1 GPU): does the matcher start-up stagger help or hurt the pipeline kernel?
mkdir -p gpurun_out
export SCONE_B200_LIB=$PWD/scone_b200/lib/libscone_b200_tune.so
for spec in custom:1024:fp32:3:1000000:64:1024:50257 custom:768:fp32:3:100000:64:1024:50257 config2; do
  args=""
  for m in replace pos add addpos; do args="$args $m;STABLE=1 $m;STABLE=1,SCONE_STAGGER_NS=-1 $m;STABLE=1,SCONE_STAGGER_NS=200 $m;STABLE=1,SCONE_EMBED_VARIANT=2:2:5:4:3:70 $m;STABLE=1,SCONE_EMBED_VARIANT=2:2:5:4:3:70,SCONE_STAGGER_NS=500"; done
  f=gpurun_out/modes10_$(echo $spec | tr ':' '_').log
  timeout 600 python tools/tune_modes.py $spec $args > $f 2>&1; echo "$spec rc=$?"
  python - $f <<'PY'
import json, sys
rows = [json.loads(l) for l in open(sys.argv[1]) if l.startswith('{"workload')]
for x in rows: print(x['mode'], x['env'], round(x['us'], 1), x['same_bits_as_first'])
PY
done
