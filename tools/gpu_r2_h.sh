# round-2 call H (1 GPU): compute-sanitizer over every kernel shape (both fused kernels, early start, builders)
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck python tools/sanitize.py > gpurun_out/sanitizer_memcheck_r02.log 2>&1; echo "memcheck rc=$?"; tail -2 gpurun_out/sanitizer_memcheck_r02.log
SANITIZE_BUILDERS=0 timeout 1200 compute-sanitizer --tool racecheck python tools/sanitize.py > gpurun_out/sanitizer_racecheck_r02.log 2>&1; echo "racecheck rc=$?"; tail -2 gpurun_out/sanitizer_racecheck_r02.log
SANITIZE_BUILDERS=0 timeout 900 compute-sanitizer --tool synccheck python tools/sanitize.py > gpurun_out/sanitizer_synccheck_r02.log 2>&1; echo "synccheck rc=$?"; tail -2 gpurun_out/sanitizer_synccheck_r02.log
timeout 600 compute-sanitizer --tool racecheck python -c "
import torch, sys
sys.path.insert(0, '.')
import scone_b200 as sb
for quant in ('fp16', 'int8', 'int4'):
    t = sb.CacheTable(300, 512, quant)
    t.store_projected(torch.randn(300, 128, device='cuda'), torch.randn(512, 128, device='cuda'))
torch.cuda.synchronize(); print('fold under racecheck done')
" > gpurun_out/sanitizer_racecheck_fold_r02.log 2>&1; echo "racecheck fold rc=$?"; tail -3 gpurun_out/sanitizer_racecheck_fold_r02.log
for f in gpurun_out/sanitizer_*_r02.log; do echo "== $f"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard" $f | tail -3; done
