# round 2, second session: launch list of the final step, ncu --set full of the Bottleneck-junction kernel and of the
# two-input 1x1 convolution (each only after its command has run clean without ncu)
set -x
cd $GRAFT_REPO_ROOT
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/r02b_ncu_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/r02b_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/r02b_ncu_launch.log 2>&1
python tools/link_bench.py --iters 3 > gpurun_out/r02b_linkbench.txt 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:bottleneck_link -s 1 -c 1 -o gpurun_out/r02b_link python tools/link_bench.py --iters 2 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 12 -c 1 -o gpurun_out/r02b_cat python tools/link_bench.py --iters 2 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
