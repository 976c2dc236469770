set -x
cd $GRAFT_REPO_ROOT
M=gpu__time_duration.sum,lts__t_sector_hit_rate.pct,dram__bytes_read.sum,dram__bytes_write.sum
for T in 23 22; do
SNRF_PROFILE_LOG2T=$T SNRF_FWD_SPLIT=1 ncu --profile-from-start off --clock-control none --metrics $M -k regex:field_fwd --csv --log-file gpurun_out/r3r_fwd_levels_T$T.csv python tools/profile_step.py > gpurun_out/r3r_ncu_$T.log 2>&1
done
ls gpurun_out | grep r3r
