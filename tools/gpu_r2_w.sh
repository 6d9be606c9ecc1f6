# round-2 call W (1 GPU): plain path, single-ring vs pipeline kernel over row formats and widths (where does the pipeline win?)
mkdir -p gpurun_out
export SCONE_B200_LIB=$PWD/scone_b200/lib/libscone_b200_tune.so
for qd in fp32:768 fp32:1024 fp32:1280 fp32:1536 fp32:2048 fp16:1024 fp16:1536 fp16:2048 fp16:3072 fp16:4096 int8:1024 int8:1536 int8:2048 int8:3072 int8:4096 int4:2048 int4:4096; do
  q=${qd%%:*}; d=${qd##*:}
  spec=custom:$d:$q:4:1000000:64:1024:50257
  f=gpurun_out/modes12_${q}_$d.log
  timeout 300 python tools/tune_modes.py $spec "replace;STABLE=1,SCONE_EMBED_PIPE=0" "replace;STABLE=1,SCONE_EMBED_PIPE=1" "replace;STABLE=1,SCONE_EMBED_VARIANT=2:2:4:8:2:110" "replace;STABLE=1,SCONE_EMBED_VARIANT=2:4:6:12:1:200" "replace;STABLE=1,SCONE_EMBED_VARIANT=2:2:5:4:3:70" "replace;STABLE=1" > $f 2>&1
  python - $f $q $d <<'PY'
import json, sys
rows = [json.loads(l) for l in open(sys.argv[1]) if l.startswith('{"workload')]
print(sys.argv[2], sys.argv[3], "  ".join(f"{(x['env'].split(',')[-1]).replace('SCONE_EMBED_','')} {x['us']:.1f}" for x in rows))
PY
done
