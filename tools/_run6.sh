mkdir -p gpurun_out
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r02_s3.json 2> gpurun_out/bench_r02_s3.err
tail -c 1500 gpurun_out/bench_r02_s3.json
M="--metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv"
ncu $M --log-file gpurun_out/r02_launches_forward.csv python tools/profile_forward.py > /dev/null 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_bytes.sum --clock-control none --profile-from-start off -k regex:conv_tc_kernel --csv --log-file gpurun_out/r02_conv_tc_traffic.csv python tools/profile_forward.py > /dev/null 2>&1
ncu $M -c 520 --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 1 --warmup 3 --ncu-window --no-cpu-baseline --no-decode --no-extras > gpurun_out/ncu_bench.log 2>&1
ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:conv_tc -c 8 -f -o gpurun_out/r02_conv_gna python tools/microbench.py --only conv3_256_l0 --ncu > gpurun_out/ncu_gna.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/r02_*.csv
