set -x
cd $GRAFT_REPO_ROOT
python -m pytest tests/test_field_encode_gpu.py -x -q -k "load_variants or fused_encode_matches" > gpurun_out/r3a_tests.log 2>&1; tail -3 gpurun_out/r3a_tests.log
python tools/sweep_fwd.py --out gpurun_out/r3a_fwd_sweep.json > gpurun_out/r3a_sweep.log 2>&1
M=gpu__time_duration.sum,lts__t_sector_hit_rate.pct,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sectors_srcunit_tex_op_read_lookup_hit.sum,l1tex__t_sector_hit_rate.pct,smsp__issue_active.avg.pct_of_peak_sustained_active
SNRF_FWD_SPLIT=1 SNRF_FWD_L2=1,0 SNRF_FWD_PAIR=2,0 ncu --profile-from-start off --clock-control none --metrics $M -k regex:field_fwd --csv --log-file gpurun_out/r3a_fwd_levels_pair2.csv python tools/profile_step.py > gpurun_out/r3a_ncu1.log 2>&1
SNRF_FWD_SPLIT=1 SNRF_FWD_L2=1,0 SNRF_FWD_PAIR=1,0 ncu --profile-from-start off --clock-control none --metrics $M -k regex:field_fwd --csv --log-file gpurun_out/r3a_fwd_levels_pair1.csv python tools/profile_step.py > gpurun_out/r3a_ncu2.log 2>&1
grep "^{" gpurun_out/r3a_sweep.log
