# full round-end style check on the GPU box: tests, smoke, bench (+ reference arm), ncu evidence
mkdir -p gpurun_out
python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout 300 compute-sanitizer --tool memcheck python tools/sanitize.py > gpurun_out/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -1 gpurun_out/sanitizer_memcheck.log
if [ "$QUICK" != "1" ]; then
timeout 600 python bench.py > gpurun_out/bench_config2.json 2> gpurun_out/bench_config2.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_config2.json')); print(d['value']/1e6, d['ms_per_step'], d['roofline']['frac'], d['e2e']['value']/1e6, d['cpu_baseline']['value'], d['gpu_launches'], d['clocks'])"; tail -3 gpurun_out/bench_config2.err
timeout 600 python bench.py --impl reference > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/bench_reference.json
timeout 900 python bench.py --workload config3 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_config3.json 2> gpurun_out/bench_config3.err; python -c "
import json; d=json.load(open('gpurun_out/bench_config3.json')); print('config3', d['value']/1e6, d['ms_per_step'], d['roofline']['frac'], d['e2e']['value']/1e6)"
timeout 600 python bench.py --workload config1 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_config1.json 2> gpurun_out/bench_config1.err;  python -c "
import json; d=json.load(open('gpurun_out/bench_config1.json')); print('config1', d['value']/1e6, d['ms_per_step'], d['roofline']['frac'], d['e2e']['value']/1e6)"
# one --set full report is ~40 MB and gpurun copies back at most 64 MiB: NCU=config2 (default) | config3 | none per call
NCU=${NCU:-config2}
if [ "$NCU" != "none" ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:embed -s 3 -c 1 -o gpurun_out/prof_embed_${NCU}_r01 -f python tools/prof_embed.py $NCU 6 > gpurun_out/ncu_full_${NCU}.log 2>&1; tail -1 gpurun_out/ncu_full_${NCU}.log
fi
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_bench_config2_r01.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1; tail -1 gpurun_out/bench_under_ncu.log | cut -c1-120
fi
