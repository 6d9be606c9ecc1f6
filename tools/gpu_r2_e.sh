# round-2 call E (1 GPU): full GPU tests, fit at scale, host tier with the huge-page allocator, ncu evidence, final suite
mkdir -p gpurun_out
python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu.log | cut -c1-300
timeout 900 python tools/bench_fit.py 1e8 4 1e7 > gpurun_out/bench_fit.log 2>&1; echo "bench_fit rc=$?"; tail -2 gpurun_out/bench_fit.log | cut -c1-600
cat /sys/kernel/mm/transparent_hugepage/enabled; grep -E "MemTotal|MemAvailable|AnonHugePages" /proc/meminfo
timeout 900 python bench.py --workload config5 --steps 10 --warmup 3 > gpurun_out/bench_config5.json 2> gpurun_out/bench_config5.err; echo "config5 rc=$?"; tail -2 gpurun_out/bench_config5.err | cut -c1-300
grep -E "AnonHugePages" /proc/meminfo
python - <<'PY'
import json
try:
    d = json.loads([l for l in open('gpurun_out/bench_config5.json') if l.startswith('{')][-1])
    print('config5', d['config']['f_grams'], 'rows', round(d['config']['table_bytes_host']/1e9,1), 'GB pin_s', round(d['config']['pin_seconds'],1), round(d['value']/1e6,2), 'Mtok/s link', round(d['roofline']['achieved'],1), 'GB/s', d['clocks'], 'staged', d['staged'])
except Exception as e:
    print('parse failed', repr(e))
PY
# ncu: --set full captures summarised here (the reports are too large to bring back)
prof() {  # name workload mode
  timeout 600 ncu --set full --clock-control none --cache-control none --import-source on -k regex:embed -s 4 -c 1 -o /tmp/prof_$1 -f python tools/prof_embed.py $2 8 $3 stable > gpurun_out/ncu_full_$1.log 2>&1
  python tools/ncu_summary.py /tmp/prof_$1.ncu-rep $1 r02 > /dev/null 2>&1; cp profiles/ncu_embed_$1_r02.md profiles/traffic_$1.json gpurun_out/ 2>/dev/null; tail -1 gpurun_out/ncu_full_$1.log | cut -c1-120
}
prof config2 config2 replace
prof config3 config3 replace
prof config2_addpos config2 addpos
prof config2_pos config2 pos
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fold -s 2 -c 1 -o /tmp/prof_fold -f python tools/bench_fold.py 262144 > gpurun_out/ncu_full_fold.log 2>&1; tail -1 gpurun_out/ncu_full_fold.log | cut -c1-120
ncu -i /tmp/prof_fold.ncu-rep --page raw --csv > gpurun_out/ncu_fold_raw.csv 2>/dev/null
timeout 600 ncu --metrics gpu__time_duration.sum,pcie__read_bytes.sum,pcie__write_bytes.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:embed -c 6 --csv --log-file gpurun_out/ncu_config5_pcie.csv python bench.py --workload config5 --rows-per-gpu 4000000 --no-staged --steps 3 --warmup 3 > gpurun_out/ncu_config5.log 2>&1; tail -3 gpurun_out/ncu_config5_pcie.csv | cut -c1-300
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_bench_config2_r02.csv python bench.py --workload config2 --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1; tail -1 gpurun_out/bench_under_ncu.log | cut -c1-100
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_suite_n1.json 2> gpurun_out/bench_suite_n1.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_suite_n1.err
python - <<'PY'
import json
try:
    d = json.loads(open('gpurun_out/bench_suite_n1.json').read().strip().splitlines()[-1])
    r = d['roofline']
    print('config2', round(d['value']/1e6,1), 'Mtok/s', round(d['ms_per_step']*1e3,2), 'us frac', round(r['frac'],4), 'isolated', round(r['isolated_kernel_ms']*1e3,2), 'us', round(r['isolated_frac'],4),
          'e2e', round(d['e2e']['value']/1e6,1), 'sync', round(d['e2e']['synchronous']/1e6,1), 'to_host', round(d['e2e']['embeds_to_host']['value']/1e6,1), d['clocks'], 'suite_s', d.get('suite_seconds'))
    for k, c in d.get('configs', {}).items():
        if 'error' in c:
            print(k, c)
        else:
            print(k, round(c['value']/1e6,2), 'Mtok/s', round(c['ms_per_step']*1e3,1), 'us frac', round(c['roofline']['frac'],4), c['clocks'], 'wall', round(c['wall_seconds'],1), c['config'].get('f_grams'), c.get('staged'))
except Exception as e:
    print('parse failed', e)
PY
