# The GPU checks a round ends with (one B200, through gpurun): all -m gpu tests, a live ncu capture of the default fused kernel (rewrites profiles/traffic.json for the
# kernel sources of the tree), bench.py (both arms), the ncu launch list of the bench command and smoke().  Outputs under gpurun_out/${TAG}_*; copy what is to be kept to profiles/.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu --tb=short > gpurun_out/${TAG:-r2y}_gputests.log 2>&1; tail -4 gpurun_out/${TAG:-r2y}_gputests.log
timeout 900 python tools/capture_traffic.py gpurun_out/${TAG:-r2y}_traffic steady 2>&1 | tail -1
cp gpurun_out/${TAG:-r2y}_traffic.json profiles/traffic.json
timeout 900 python bench.py > gpurun_out/${TAG:-r2y}_bench_n1.json 2> gpurun_out/${TAG:-r2y}_bench_n1.err; tail -c 300 gpurun_out/${TAG:-r2y}_bench_n1.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG:-r2y}_bench_reference.json 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/${TAG:-r2y}_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extra > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/${TAG:-r2y}_launches.csv > gpurun_out/${TAG:-r2y}_launches_summary.txt 2>&1; head -8 gpurun_out/${TAG:-r2y}_launches_summary.txt
python __graft_entry__.py smoke 2>&1 | tail -1
