mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/bench_config2.json 2> gpurun_out/bench_config2.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_config2.json')); print(d['value']/1e6, d['ms_per_step'], d['ms_per_step_repeats'], d['roofline']['frac'], d['e2e']['value']/1e6, d['config']['index_slot_bytes'])"; tail -3 gpurun_out/bench_config2.err
timeout 600 python bench.py --id-dist zipf --no-cpu-baseline > gpurun_out/bench_config2_zipf.json 2> gpurun_out/bench_config2_zipf.err; python -c "
import json; d=json.load(open('gpurun_out/bench_config2_zipf.json')); print('zipf', d['value']/1e6, d['ms_per_step'], d['roofline']['frac'], d['e2e']['value']/1e6)"
timeout 600 python bench.py --workload config3 --id-dist zipf --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_config3_zipf.json 2> gpurun_out/bench_config3_zipf.err; python -c "
import json; d=json.load(open('gpurun_out/bench_config3_zipf.json')); print('config3 zipf', d['value']/1e6, d['ms_per_step'], d['roofline']['frac'])"
timeout 600 python bench.py --workload config1 --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_config1.json 2> gpurun_out/bench_config1.err;  python -c "
import json; d=json.load(open('gpurun_out/bench_config1.json')); print('config1', d['value']/1e6, d['ms_per_step'], d['roofline']['frac'], d['e2e']['value']/1e6)"
