mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -30 > gpurun_out/pytest_gpu.log
tail -2 gpurun_out/pytest_gpu.log
python tools/profile_forward.py --breakdown > gpurun_out/r02_breakdown.log 2>&1; head -3 gpurun_out/r02_breakdown.log
python bench.py > gpurun_out/bench_r02_final.json 2> gpurun_out/bench_r02_final.err; tail -c 600 gpurun_out/bench_r02_final.json
