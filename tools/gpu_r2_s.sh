# round-2 call S (1 GPU): the reference's own preset shapes in every mode with the final kernel choice, both kernels forced, and shape sweeps
mkdir -p gpurun_out
export SCONE_B200_LIB=$PWD/scone_b200/lib/libscone_b200_tune.so
M="replace pos add addpos"
for spec in custom:384:fp16:3:100000:64:1024:50257 custom:768:fp16:3:100000:64:1024:50257 custom:768:fp32:3:100000:64:1024:50257 custom:1024:fp32:3:1000000:64:1024:50257; do
  args=""
  for m in $M; do args="$args $m; $m;STABLE=1 $m;STABLE=1,SCONE_EMBED_PIPE=0 $m;STABLE=1,SCONE_EMBED_PIPE=1"; done
  for m in $M; do
    for v in 2:1:6:4:3:70 2:2:4:4:3:70 2:2:5:4:3:70 2:1:5:4:3:70 2:2:4:8:2:110 2:4:4:8:2:110 2:2:6:8:2:110 2:2:8:8:2:110 2:4:6:12:1:200 1:0:6:4:3:70 1:0:4:4:3:70 1:0:4:6:3:70 1:0:6:6:3:70 1:0:4:8:2:100 1:0:6:6:2:100 1:0:5:3:4:54 1:0:6:2:4:54; do
      args="$args $m;STABLE=1,SCONE_EMBED_VARIANT=$v"
    done
  done
  f=gpurun_out/modes8_$(echo $spec | tr ':' '_').log
  timeout 600 python tools/tune_modes.py $spec $args > $f 2>&1; echo "$spec rc=$?"
  python - $f <<'PY'
import json, sys
rows = [json.loads(l) for l in open(sys.argv[1]) if l.startswith('{"workload')]
for m in ("replace", "pos", "add", "addpos"):
    r = [x for x in rows if x["mode"] == m]
    base = r[:4]
    best = sorted(r[4:], key=lambda x: x["us"])[:4]
    print(m, " | ".join(f"{x['env'] or 'default'}: {x['us']:.1f}" for x in base))
    print("   best:", " | ".join(f"{x['env'].split('VARIANT=')[-1]}: {x['us']:.1f}{'' if x['same_bits_as_first'] else ' BITS!'}" for x in best))
PY
done
