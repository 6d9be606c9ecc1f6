# round-2 call Q (1 GPU): "both" mode (base row + wpe) shapes of the pipeline kernel with a 110-112 KB ring, 2 CTAs/SM
mkdir -p gpurun_out
export SCONE_B200_LIB=$PWD/scone_b200/lib/libscone_b200_tune.so
V() { echo "addpos;STABLE=1,SCONE_EMBED_PIPE=1,SCONE_EMBED_VARIANT=2:$1"; }
timeout 600 python tools/tune_modes.py config2 "addpos;STABLE=1" "addpos;STABLE=1,SCONE_EMBED_PIPE=1" "$(V 2:4:8:2:110)" "$(V 4:4:8:2:110)" "$(V 2:6:8:2:110)" "$(V 4:6:8:2:110)" "$(V 2:8:8:2:110)" "$(V 4:6:12:1:200)" "$(V 4:8:16:1:200)" \
  "addpos;STABLE=1,SCONE_EMBED_PIPE=1,SCONE_EMBED_P=8,SCONE_EMBED_VARIANT=2:2:4:8:2:100" "addpos;STABLE=1,SCONE_EMBED_PIPE=1,SCONE_EMBED_P=8,SCONE_EMBED_VARIANT=2:2:6:8:2:100" "addpos;STABLE=1,SCONE_EMBED_PIPE=1,SCONE_EMBED_P=8,SCONE_EMBED_VARIANT=2:4:6:8:2:100" \
  "add;STABLE=1" "$(echo 'add;STABLE=1,SCONE_EMBED_VARIANT=2:1:6:4:3:70')" "$(echo 'add;STABLE=1,SCONE_EMBED_VARIANT=2:2:6:8:2:100')" "pos;STABLE=1" > gpurun_out/modes6_config2.log 2>&1; cut -c1-220 gpurun_out/modes6_config2.log
