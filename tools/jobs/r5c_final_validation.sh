cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/r5c_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/r5c_tests.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r5c_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r5c_smoke.log
S=$(date +%s); timeout 600 python bench.py > gpurun_out/r5c_bench.json 2> gpurun_out/r5c_bench.err; echo "bench rc=$? $(( $(date +%s) - S )) s"; cut -c1-300 gpurun_out/r5c_bench.json
S=$(date +%s); timeout 300 python bench.py --impl reference --steps 2 --warmup 3 > gpurun_out/r5c_bench_reference.json 2> gpurun_out/r5c_bench_reference.err; echo "ref rc=$? $(( $(date +%s) - S )) s"; cut -c1-300 gpurun_out/r5c_bench_reference.json
