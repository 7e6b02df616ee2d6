#!/bin/bash
# last confirmation of round 2 on one GPU: the driver's own commands + the ncu launch list of the final build
O=gpurun_out; T=r02h
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee $O/${T}_pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee $O/${T}_smoke.txt
timeout 600 python bench.py > $O/${T}_bench_n1.json 2> $O/${T}_bench_n1.err; tail -c 200 $O/${T}_bench_n1.err
timeout 300 python bench.py --impl reference --steps 4 --warmup 1 > $O/${T}_bench_ref.json 2>/dev/null
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file $O/${T}_ncu_launches_bench_step.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-e2e --no-cuda-graph > $O/${T}_ncu_launch.log 2>&1
python tools/ncu_times.py $O/${T}_ncu_launches_bench_step.csv > $O/${T}_ncu_launch_summary.txt 2>&1; head -8 $O/${T}_ncu_launch_summary.txt
python tools/show_bench.py $O/${T}_bench_n1.json
