set -x
cd $GRAFT_REPO_ROOT
python tools/profile_step.py > gpurun_out/r4m_plain.log 2>&1 || exit 1
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:'field_fwd_kernel|field_geom_raygrad_kernel' -o gpurun_out/r4m_prof_enc -f python tools/profile_step.py > gpurun_out/r4m_ncu1.log 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:'field_scatter_slice_kernel|adam_slice_kernel' --launch-skip 10 --launch-count 2 -o gpurun_out/r4m_prof_mid -f python tools/profile_step.py > gpurun_out/r4m_ncu2.log 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:'field_scatter_slice_kernel|adam_slice_kernel' --launch-skip 44 --launch-count 2 -o gpurun_out/r4m_prof_fine -f python tools/profile_step.py > gpurun_out/r4m_ncu3.log 2>&1
ls -la gpurun_out | grep r4m
