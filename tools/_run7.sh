mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4 | tee gpurun_out/smoke.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --no-extras --no-cpu-baseline --no-decode 2>/dev/null | tail -1 > gpurun_out/sample_n2.json
python -c "import json; d=json.loads(open('gpurun_out/sample_n2.json').read()); print('N=2 sample', d['value'], d['ms_per_step'], d['n_gpus'])"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --workload train --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/train_n2.json
python -c "import json; d=json.loads(open('gpurun_out/train_n2.json').read()); print('N=2 train', d['value'], d['ms_per_step'], d['n_gpus'])"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --global-batch 512 --steps 2 --warmup 3 --no-extras --no-cpu-baseline --no-decode 2>/dev/null | tail -1 > gpurun_out/strong_n2.json
python -c "import json; d=json.loads(open('gpurun_out/strong_n2.json').read()); print('N=2 strong', d['value'], d['ms_per_step'], d['n_gpus'])"
