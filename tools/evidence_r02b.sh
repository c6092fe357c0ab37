cd $GRAFT_REPO_ROOT
for i in 1 2 3; do python -m pytest tests -m gpu -x -q 2>&1 | tail -1; done > gpurun_out/r02b_gputest.txt 2>&1
python bench.py --steps 20 --warmup 5 --dump-ops gpurun_out/r02b_ops.json > gpurun_out/r02b_bench_final.json 2> gpurun_out/r02b_bench_final.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02b_bench_reference.json 2> gpurun_out/r02b_bench_reference.err
python bench.py --width 48 --batch 256 --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/r02b_bench_w48.json 2> gpurun_out/r02b_bench_w48.err
python bench.py --workload train --batch 128 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02b_bench_train128.json 2> gpurun_out/r02b_bench_train128.err
python bench.py --workload train --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02b_bench_train32.json 2> gpurun_out/r02b_bench_train32.err
python tools/link_bench.py > gpurun_out/r02b_linkbench_full.txt 2>&1
cat gpurun_out/r02b_gputest.txt; cut -c1-400 gpurun_out/r02b_bench_final.json; cat gpurun_out/r02b_linkbench_full.txt
