# round-2 call I (1 GPU): L2 policy of the fallback / base rows (development build), config 1-3 and the modes
mkdir -p gpurun_out
export SCONE_B200_LIB=$PWD/scone_b200/lib/libscone_b200_tune.so
for c in config2 config3 config1; do
timeout 600 python tools/tune_modes.py $c "replace;STABLE=1" "replace;STABLE=1,SCONE_BASE_POLICY=1" "replace;STABLE=1,SCONE_BASE_POLICY=2" "replace;STABLE=1" "replace;STABLE=1,SCONE_BASE_POLICY=2" "add;STABLE=1" "add;STABLE=1,SCONE_BASE_POLICY=1" "add;STABLE=1,SCONE_BASE_POLICY=2" "addpos;STABLE=1" "addpos;STABLE=1,SCONE_BASE_POLICY=2" "pos;STABLE=1" "pos;STABLE=1,SCONE_BASE_POLICY=2" > gpurun_out/modes5_$c.log 2>&1; cut -c1-200 gpurun_out/modes5_$c.log
done
