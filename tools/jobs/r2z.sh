set -x
cd $GRAFT_REPO_ROOT
python tools/sweep_l2gran.py --out gpurun_out/r2z_l2gran_sweep.json > gpurun_out/r2z_sweep.log 2>&1
tail -12 gpurun_out/r2z_sweep.log
