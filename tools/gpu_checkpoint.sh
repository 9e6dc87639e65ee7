#!/bin/bash
# One GPU checkpoint = what every profiles/rNN_* set of this repo comes from.  Run it on a B200 box from the repo root
# (here: gpurun --timeout 1500 -- 'bash tools/gpu_checkpoint.sh TAG'); everything lands in gpurun_out/TAG_*.
#   1. the GPU test suite                      2. the bench line (default path G, all configs, CPU arm)
#   3. phase cycles of the three IC paths      4. Test 4 at full size
#   5. the ncu launch list of the bench        6. ncu --set full of the three dominant kernels, exported as text/CSV
# Numbers are only ever taken from the plain runs (1-4); the ncu passes come after them and are never timed.
TAG=${1:-ckpt}
O=gpurun_out/$TAG
set -x
python -m pytest tests -m gpu -q > ${O}_pytest.log 2>&1; echo "pytest rc=$?" >> ${O}_pytest.log
python bench.py --steps 10 --warmup 3 > ${O}_b1.json 2> ${O}_b1.err; echo "bench rc=$?"
for p in G R S; do python tools/prof_run.py --walkers 1184 --ic-path $p; done > ${O}_phases.txt 2>&1
python tools/bench_test4.py > ${O}_test4.txt 2>&1
python tools/bench_sweep.py --walkers 4096 --grids 100 200 400 --ic-path G > ${O}_sweep.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file ${O}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-configs --no-cpu-baseline --no-fast-path > ${O}_ncu_launch.log 2>&1
cap() {  # name, kernel regex, launches to skip, command...
  local n=$1 k=$2 s=$3; shift 3
  ncu --set full --clock-control none --import-source on -k regex:$k -s $s -c 1 -o ${O}_$n "$@" > ${O}_ncu_$n.log 2>&1
  ncu -i ${O}_$n.ncu-rep --page raw --csv > ${O}_${n}_raw.csv 2>/dev/null
  ncu -i ${O}_$n.ncu-rep --page details > ${O}_${n}_details.txt 2>/dev/null
  ncu -i ${O}_$n.ncu-rep --page source --csv --print-source cuda,sass > ${O}_${n}_src.csv 2>/dev/null
  rm -f ${O}_$n.ncu-rep          # the reports are 20-40 MB each; gpurun_out/ carries 64 MiB back
}
cap icg '^icg_kernel' 3 python tools/prof_run.py --walkers 1184 --ic-path G --reps 1
cap stepG stepG_kernel 3 python tools/prof_run.py --walkers 1184 --ic-path G --reps 1
cap lh run_lh_kernel 1 python tools/bench_test4.py
ls -la gpurun_out/ | tail -30
