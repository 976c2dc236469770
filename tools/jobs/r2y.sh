set -x
cd $GRAFT_REPO_ROOT
python tools/sweep_fwd.py --out gpurun_out/r2y_fwd_l2_sweep.json > gpurun_out/r2y_sweep.log 2>&1
M=gpu__time_duration.sum,lts__t_sector_hit_rate.pct,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_read_lookup_hit.sum
SNRF_FWD_SPLIT=1 ncu --profile-from-start off --clock-control none --metrics $M -k regex:field_fwd --csv --log-file gpurun_out/r2y_fwd_levels_default.csv python tools/profile_step.py > gpurun_out/r2y_ncu1.log 2>&1
SNRF_FWD_SPLIT=1 SNRF_FWD_L2=2,64 ncu --profile-from-start off --clock-control none --metrics $M -k regex:field_fwd --csv --log-file gpurun_out/r2y_fwd_levels_pin64.csv python tools/profile_step.py > gpurun_out/r2y_ncu2.log 2>&1
SNRF_FWD_SPLIT=1 SNRF_FWD_L2=1,0 ncu --profile-from-start off --clock-control none --metrics $M -k regex:field_fwd --csv --log-file gpurun_out/r2y_fwd_levels_all_last.csv python tools/profile_step.py > gpurun_out/r2y_ncu3.log 2>&1
tail -3 gpurun_out/r2y_sweep.log
