set -x
cd $GRAFT_REPO_ROOT
python tools/profile_step.py > gpurun_out/r3j_plain.log 2>&1 || exit 1
ncu --profile-from-start off --cache-control none --clock-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv --log-file gpurun_out/r3j_step_kernels_warm.csv python tools/profile_step.py > gpurun_out/r3j_ncu1.log 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:'decoder_bwd_fold_kernel|decoder_fwd4_kernel' -o gpurun_out/r3j_prof_dec -f python tools/profile_step.py > gpurun_out/r3j_ncu2.log 2>&1
ls -la gpurun_out/ | tail -5
