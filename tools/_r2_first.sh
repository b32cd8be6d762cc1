set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv
timeout 1500 python -m pytest tests -x -q -m gpu --tb=short > gpurun_out/r2a_gputests.log 2>&1; tail -5 gpurun_out/r2a_gputests.log
timeout 600 python bench.py > gpurun_out/r2a_bench_n1.json 2> gpurun_out/r2a_bench_n1.err; tail -c 3000 gpurun_out/r2a_bench_n1.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2a_bench_reference.json 2>&1
timeout 900 python tools/capture_traffic.py gpurun_out/r2a_traffic 2>&1 | tail -2
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2a_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/r2a_launches_bench.log 2>&1
python tools/launch_summary.py gpurun_out/r2a_launches.csv > gpurun_out/r2a_launches_summary.txt 2>&1; head -30 gpurun_out/r2a_launches_summary.txt
