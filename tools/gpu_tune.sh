# scratch: bench line + ncu evidence of config 2 for the final build
mkdir -p gpurun_out
timeout 70 python bench.py > gpurun_out/bench_config2_final.json 2> gpurun_out/bench_config2_final.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_config2_final.json')); print(d['value']/1e6, d['ms_per_step'], d['roofline']['frac'], d['e2e']['value']/1e6, d['cpu_baseline']['value'], d['gpu_launches'], d['clocks'])"
timeout 45 ncu --set full --clock-control none --import-source on -k regex:embed -s 3 -c 1 -o gpurun_out/prof_embed_config2_r01 -f python tools/prof_embed.py config2 6 > gpurun_out/ncu_full2.log 2>&1; tail -1 gpurun_out/ncu_full2.log
timeout 45 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_bench_config2_r01.csv python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1; tail -1 gpurun_out/bench_under_ncu.log | cut -c1-120
ls -la gpurun_out | head
