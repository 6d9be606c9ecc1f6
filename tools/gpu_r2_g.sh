# round-2 call G (1 GPU): experiment -- fallback rows bypass the ring (SCONE_MISS_DIRECT=1), with shape sweeps under early start
mkdir -p gpurun_out
SCONE_MISS_DIRECT=1 timeout 900 python -m pytest tests -m gpu -x -q -k "embed_forward_matches_oracle or config2_full or all_hits_and_all_misses or status_and_fallback or config3_full or inputs_stable" > gpurun_out/pytest_missdirect.log 2>&1; echo "pytest (miss direct) rc=$?"; tail -3 gpurun_out/pytest_missdirect.log | cut -c1-200
export SCONE_B200_LIB=$PWD/scone_b200/lib/libscone_b200_tune.so
V() { echo "replace;STABLE=1,SCONE_EMBED_VARIANT=1:0:$1"; }
VM() { echo "replace;STABLE=1,SCONE_MISS_DIRECT=1,SCONE_EMBED_VARIANT=1:0:$1"; }
timeout 900 python tools/tune_modes.py config2 "replace;STABLE=1" "replace;STABLE=1,SCONE_MISS_DIRECT=1" \
  "$(V 4:4:3:70)" "$(V 4:6:3:70)" "$(V 3:5:3:70)" "$(V 6:6:2:100)" "$(V 4:8:2:100)" "$(V 6:10:2:100)" \
  "$(VM 4:4:3:70)" "$(VM 6:4:3:70)" "$(VM 5:5:3:70)" "$(VM 6:6:3:70)" "$(VM 4:6:3:70)" "$(VM 5:3:4:52)" "$(VM 6:2:4:52)" "$(VM 8:4:2:100)" "$(VM 6:6:2:100)" "$(VM 8:6:2:100)" "$(VM 6:10:2:100)" \
  "replace;STABLE=1" "replace;STABLE=1,SCONE_MISS_DIRECT=1" > gpurun_out/modes4_config2.log 2>&1; cut -c1-210 gpurun_out/modes4_config2.log
timeout 900 python tools/tune_modes.py config3 "replace;STABLE=1" "replace;STABLE=1,SCONE_MISS_DIRECT=1" "$(VM 6:12:1:200)" "$(VM 8:12:1:200)" "$(VM 8:16:1:200)" "$(VM 12:12:1:200)" "$(VM 6:18:1:200)" "$(VM 6:10:2:100)" "$(V 8:12:1:200)" "$(V 8:16:1:200)" > gpurun_out/modes4_config3.log 2>&1; cut -c1-210 gpurun_out/modes4_config3.log
timeout 300 python tools/tune_modes.py config1 "replace;STABLE=1" > gpurun_out/modes4_config1.log 2>&1; cut -c1-210 gpurun_out/modes4_config1.log
