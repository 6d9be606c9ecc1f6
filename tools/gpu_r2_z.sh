# round-2 call Z (1 GPU): how many clusters of 2..8 CTAs does the fold get?
mkdir -p gpurun_out
SCONE_FOLD_DEBUG=1 timeout 300 python tools/bench_fold.py 262144 2>&1 | grep -E "scone fold|^\{" | sort | uniq -c | sort -rn | cut -c1-200 | head -30
