# scratch: stagger sweep with the pre-filter on; new kMid / kWide2 shapes (SCONE_TUNE build)
mkdir -p gpurun_out
timeout 300 python tools/tune_modes.py config2 \
  "replace;SCONE_STAGGER_NS=-1" "replace;SCONE_STAGGER_NS=200" "replace;SCONE_STAGGER_NS=300" "replace;SCONE_STAGGER_NS=400" "replace;SCONE_STAGGER_NS=500" "replace;SCONE_STAGGER_NS=600" "replace;SCONE_STAGGER_NS=800" "replace;" \
  "pos;SCONE_STAGGER_NS=-1" "pos;" "add;SCONE_STAGGER_NS=-1" "add;" "addpos;" "addpos;SCONE_EMBED_VARIANT=1:0:2:6:3:70" \
  > gpurun_out/tune_final_config2.log 2>&1; echo "rc=$?"; cut -c1-200 gpurun_out/tune_final_config2.log | tail -15
timeout 300 python tools/tune_modes.py config1 "replace;" > gpurun_out/tune_final_config1.log 2>&1; echo "rc=$?"; cut -c1-200 gpurun_out/tune_final_config1.log | tail -1
timeout 300 python tools/tune_modes.py config3 "replace;" "pos;" "add;" "addpos;" "addpos;SCONE_EMBED_P=8,SCONE_EMBED_VARIANT=1:0:3:12:1:200" > gpurun_out/tune_final_config3.log 2>&1; echo "rc=$?"; cut -c1-200 gpurun_out/tune_final_config3.log | tail -5
