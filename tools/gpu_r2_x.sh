# round-2 call X (1 GPU): final evidence on the final build -- smoke, GPU tests, sanitizers, ncu captures, launch list, the driver's bench command
mkdir -p gpurun_out
python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 900 compute-sanitizer --tool memcheck python tools/sanitize.py > gpurun_out/sanitizer_memcheck_r02.log 2>&1; echo "memcheck rc=$?"
SANITIZE_BUILDERS=0 timeout 1200 compute-sanitizer --tool racecheck python tools/sanitize.py > gpurun_out/sanitizer_racecheck_r02.log 2>&1; echo "racecheck rc=$?"
SANITIZE_BUILDERS=0 timeout 900 compute-sanitizer --tool synccheck python tools/sanitize.py > gpurun_out/sanitizer_synccheck_r02.log 2>&1; echo "synccheck rc=$?"
timeout 600 compute-sanitizer --tool racecheck python -c "
import torch, sys
sys.path.insert(0, '.')
import scone_b200 as sb
for quant, H in (('fp16', 512), ('int8', 512), ('int8', 1024), ('int8', 2048), ('int4', 512), ('fp32', 320)):
    t = sb.CacheTable(700, H, quant)
    t.store_projected(torch.randn(700, 128, device='cuda'), torch.randn(H, 128, device='cuda'))
torch.cuda.synchronize(); print('fold under racecheck done')
" > gpurun_out/sanitizer_racecheck_fold_r02.log 2>&1; echo "racecheck fold rc=$?"
for f in gpurun_out/sanitizer_*_r02.log; do echo "== $f"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY" $f | tail -2; done
prof() {  # name workload mode
  timeout 600 ncu --set full --clock-control none --cache-control none --import-source on -k regex:embed -s 4 -c 1 -o /tmp/prof_$1 -f python tools/prof_embed.py $2 8 $3 stable > gpurun_out/ncu_full_$1.log 2>&1
  python tools/ncu_summary.py /tmp/prof_$1.ncu-rep $1 r02 > /dev/null 2>&1; cp profiles/ncu_embed_$1_r02.md profiles/traffic_$1.json gpurun_out/ 2>/dev/null; tail -1 gpurun_out/ncu_full_$1.log | cut -c1-120
}
prof config2 config2 replace
prof config3 config3 replace
prof config2_addpos config2 addpos
prof config2_pos config2 pos
prof config3_pos config3 pos
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fold -s 14 -c 8 -o /tmp/prof_fold -f python tools/bench_fold.py 262144 > gpurun_out/ncu_full_fold.log 2>&1; tail -1 gpurun_out/ncu_full_fold.log | cut -c1-120
ncu -i /tmp/prof_fold.ncu-rep --page raw --csv > gpurun_out/ncu_fold_raw.csv 2>/dev/null
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_bench_config2_r02.csv python bench.py --workload config2 --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1; tail -1 gpurun_out/bench_under_ncu.log | cut -c1-100
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "reference rc=$?"; cut -c1-400 gpurun_out/bench_reference.json
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_suite_n1.json 2> gpurun_out/bench_suite_n1.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_suite_n1.err
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/bench_suite_n1.json').read().strip().splitlines()[-1])
    r = d['roofline']
    print('config2', round(d['value']/1e6,1), 'Mtok/s', round(d['ms_per_step']*1e3,2), 'us frac', round(r['frac'],4), 'isolated', round(r['isolated_kernel_ms']*1e3,2), 'us', round(r['isolated_frac'],4),
          'e2e', round(d['e2e']['value']/1e6,1), 'sync', round(d['e2e']['synchronous']/1e6,1), 'to_host', round(d['e2e']['embeds_to_host']['value']/1e6,1), d['clocks'], 'parity', d.get('parity', {}).get('result'), 'suite_s', d.get('suite_seconds'))
    for k, c in d.get('configs', {}).items():
        if 'error' in c:
            print(k, c)
        else:
            print(k, round(c['value']/1e6,2), 'Mtok/s', round(c['ms_per_step']*1e3,1), 'us frac', round(c['roofline']['frac'],4), c['clocks'], 'wall', round(c['wall_seconds'],1), c['config'].get('f_grams'), (c.get('parity') or {}).get('result'), c.get('staged'))
except Exception as e:
    print('parse failed', e)
PY
