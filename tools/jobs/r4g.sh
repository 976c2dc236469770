set -x
cd $GRAFT_REPO_ROOT
timeout 600 python bench.py --steps 2 --warmup 3 --no-variants --no-render --no-cpu-baseline > gpurun_out/r4g_plain.json 2> gpurun_out/r4g_plain.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r4g_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-variants --no-render --no-cpu-baseline > gpurun_out/r4g_ncu.log 2>&1
timeout 300 python tools/decoder_overheads.py --out gpurun_out/r4g_decoder_overheads.json 2>&1 | grep "^{"
