mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/bench_config2.json 2> gpurun_out/bench_config2.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_config2.json')); print(d['value']/1e6, d['ms_per_step'], d['roofline']['frac'], d['e2e'])"; tail -3 gpurun_out/bench_config2.err
timeout 600 python bench.py --workload config3 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/bench_config3.json 2> gpurun_out/bench_config3.err; python -c "
import json; d=json.load(open('gpurun_out/bench_config3.json')); print(d['value']/1e6, d['ms_per_step'], d['roofline']['frac'], d['e2e'])"
cat > /tmp/san.py <<'PY'
import sys; sys.path.insert(0, '.')
import numpy as np, torch
import scone_b200 as sb
from scone_b200.utils import synthetic as S
for quant, D, max_n in (("int8", 1024, 4), ("int4", 512, 5), ("fp16", 256, 2)):
    toks, lens = S.make_vocab_numpy(2000, max_n, 300, seed=1, min_n=1 if max_n < 3 else 2)
    ix = sb.FGramIndex(torch.from_numpy(toks).cuda(), torch.from_numpy(lens).cuda())
    t = sb.CacheTable(2000, D, quant); t.store(torch.from_numpy(S.make_rows_numpy(2000, D)).cuda())
    base = torch.randn(300, D, device="cuda").to(torch.bfloat16)
    q = torch.from_numpy(S.make_stream_numpy(toks, lens, 3, 257, 300)).cuda()
    out, fid, ml = sb.embed_forward(ix, t, base, q)
    g = sb.embed_gather(t, base, q, fid)
    m = sb.embed_mean_forward(ix, t, q)
    torch.cuda.synchronize()
    assert torch.equal(out, g)
print("sanitizer workload done")
PY
timeout 900 compute-sanitizer --tool memcheck --print-limit 5 python /tmp/san.py > gpurun_out/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --print-limit 5 python /tmp/san.py > gpurun_out/sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -4 gpurun_out/sanitizer_racecheck.log
