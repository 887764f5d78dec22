mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py -x -q -k "stream_k" 2>&1 | tail -25 > gpurun_out/t_sk.log
cat gpurun_out/t_sk.log
if grep -q "passed" gpurun_out/t_sk.log && ! grep -q "failed" gpurun_out/t_sk.log; then
for v in "EALDM_TC_STREAMK=0" "EALDM_TC_STREAMK=1"; do
  echo "== variant [$v]"; env $v python tools/microbench.py --only gemm_o1_l2,conv3_1024_l2,conv3_512_l1 2>&1 | grep -v Warn
done | tee gpurun_out/mb2.log
bash tools/gpu_ab.sh "sk0:EALDM_TC_STREAMK=0" "sk1:EALDM_TC_STREAMK=1" 2>&1 | tee gpurun_out/ab2.log
fi
