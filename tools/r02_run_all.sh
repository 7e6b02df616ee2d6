#!/bin/bash
# One GPU-box session for round 2: tests, smoke, bench (C3 / C5), secondary configs, ncu launch list + full capture.
# usage (on the box): bash tools/r02_run_all.sh [tag]
T=${1:-r02}
O=gpurun_out
mkdir -p $O
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15 | tee $O/${T}_pytest_gpu.txt
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4 | tee $O/${T}_smoke.txt
echo "== bench N=1 (C3)"; timeout 600 python bench.py --steps 10 --warmup 4 > $O/${T}_bench_n1.json 2> $O/${T}_bench_n1.err; tail -c 300 $O/${T}_bench_n1.err
echo "== bench N=1 (C5)"; timeout 600 python bench.py --steps 10 --warmup 4 --workload c5 --no-cpu-baseline --no-gpu-baseline > $O/${T}_bench_c5_n1.json 2> $O/${T}_bench_c5_n1.err; tail -c 300 $O/${T}_bench_c5_n1.err
echo "== reference arm"; timeout 300 python bench.py --impl reference --steps 4 --warmup 1 > $O/${T}_bench_ref.json 2>/dev/null
echo "== secondary C2 / C4"; (timeout 300 python tools/bench_vae.py; timeout 300 python tools/bench_stn.py) > $O/${T}_secondary_c2_c4.jsonl 2>&1; tail -2 $O/${T}_secondary_c2_c4.jsonl | cut -c1-300
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file $O/${T}_ncu_launches_bench_step.csv \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-e2e --no-cuda-graph > $O/${T}_ncu_launch.log 2>&1
python tools/ncu_times.py $O/${T}_ncu_launches_bench_step.csv > $O/${T}_ncu_launch_summary.txt 2>&1; head -12 $O/${T}_ncu_launch_summary.txt
echo "== ncu full capture of the top families"
timeout 1200 ncu --set full --clock-control none --import-source on \
  -k regex:'conv1_wgrad_fold|upconv_c1|rot_sample|upsample_pad_bwd|upsample_pad_fwd|conv_tc_halo|conv1c_tc|elbo|upfold|wgrad_halo' \
  --launch-skip 170 -c 60 -o $O/${T}_top python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-gpu-baseline --no-e2e --no-cuda-graph > $O/${T}_ncu_full.log 2>&1
ncu -i $O/${T}_top.ncu-rep --page raw --csv > $O/${T}_ncu_full_top.csv 2>/dev/null
python tools/ncu_extract.py $O/${T}_ncu_full_top.csv > $O/${T}_ncu_full_top.txt 2>&1; head -50 $O/${T}_ncu_full_top.txt | cut -c1-220
rm -f $O/${T}_top.ncu-rep        # tens of MB with --import-source; gpurun copies back at most 64 MiB
python - <<PY
import csv
rows = list(csv.reader(open("$O/${T}_ncu_full_top.csv")))
keep = [i for i, h in enumerate(rows[0]) if any(s in h for s in ("Kernel Name", "Grid Size", "Block Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput", "sm__pipe_tensor", "sm__inst_issued", "sm__warps_active", "launch__registers", "lts__t_sector_hit_rate", "sm__throughput", "l1tex__data_bank_conflicts", "smsp__inst_executed.sum", "launch__occupancy"))]
csv.writer(open("$O/${T}_ncu_full_top_selected.csv", "w")).writerows([[r[i] for i in keep] for r in rows if len(r) >= len(rows[0])])
PY
rm -f $O/${T}_ncu_full_top.csv
