mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -40 > gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
bash tools/gpu_ab.sh "sk1:" "sk0:EALDM_TC_STREAMK=0" "sk1b:" 2>&1 | tee gpurun_out/ab3.log
for v in "" "EALDM_COLSUM_TWO_KERNELS=1" "EALDM_TC_STREAMK=0"; do
  echo "== train [$v]"; env $v python bench.py --workload train --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["value"], d["ms_per_step"])'
done | tee gpurun_out/train3.log
