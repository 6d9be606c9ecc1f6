# scratch: same-box A/B -- which of this session's additions costs the plain path what (libs built with SCONE_AB_* switches)
mkdir -p gpurun_out
for rep in 1 2; do
for lib in "" scone_b200/lib/libscone_old.so scone_b200/lib/libscone_noadd.so scone_b200/lib/libscone_nostag.so scone_b200/lib/libscone_noboth.so; do
AB_LIB=$lib timeout 300 python tools/tune_modes.py config1 "replace;" "replace;" 2>&1 | grep workload | cut -c1-100 | sed "s|^|lib=$lib |"
done; done
for lib in "" scone_b200/lib/libscone_old.so scone_b200/lib/libscone_noadd.so scone_b200/lib/libscone_noboth.so; do
AB_LIB=$lib timeout 300 python tools/tune_modes.py config2 "replace;" "pos;" 2>&1 | grep workload | cut -c1-100 | sed "s|^|lib=$lib |"
AB_LIB=$lib timeout 300 python tools/tune_modes.py config3 "replace;" "pos;" 2>&1 | grep workload | cut -c1-100 | sed "s|^|lib=$lib |"
done
