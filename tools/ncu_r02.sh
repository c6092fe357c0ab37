set -x
cd $GRAFT_REPO_ROOT
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/r02_ncu_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/r02_ncu_launch.log 2>&1
python tools/block_bench.py > gpurun_out/r02_blockbench.txt 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:basic_block -s 3 -c 1 -o gpurun_out/r02_block python tools/block_bench.py > /dev/null 2>&1
python tools/conv_bench.py c64 c128 --iters 3 > gpurun_out/r02_convbench_ncu.txt 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 1 -c 1 -o gpurun_out/r02_c64 python tools/conv_bench.py c64 --iters 3 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 1 -c 1 -o gpurun_out/r02_c128 python tools/conv_bench.py c128 --iters 3 > /dev/null 2>&1
python tools/conv_bench.py w48c48 --n 512 --iters 3 >> gpurun_out/r02_convbench_ncu.txt 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 1 -c 1 -o gpurun_out/r02_w48c48 python tools/conv_bench.py w48c48 --n 512 --iters 3 > /dev/null 2>&1
python bench.py --workload decode --no-sweep --steps 3 --no-cpu-baseline > gpurun_out/r02_decode_plain.json 2>&1 && \
ncu --set full --clock-control none -k regex:decode_kernel -s 3 -c 1 -o gpurun_out/r02_decode python bench.py --workload decode --no-sweep --steps 3 --no-cpu-baseline > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
