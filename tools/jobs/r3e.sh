set -x
cd $GRAFT_REPO_ROOT
timeout 300 python tools/sweep_fwd.py --out gpurun_out/r3e_fwd_sweep.json > gpurun_out/r3e_sweep.log 2>&1
grep "^{" gpurun_out/r3e_sweep.log
