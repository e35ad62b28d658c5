#!/bin/bash
# One-GPU records of a round: default bench line, triangle mesh, configs 3-5, ncu launch list and full-set captures.
#   gpurun -- 'bash tools/gpu_records.sh r02b'      -> gpurun_out/<tag>_*.{json,jsonl,csv,log}
tag=${1:-rXX}
out=gpurun_out
mkdir -p $out
NCU=/usr/local/cuda/bin/ncu
quick="--no-cpu --no-sub --no-weak --no-parity"
(time python bench.py) > $out/${tag}_bench_n1.json 2> $out/${tag}_bench_n1.err
python bench.py --mesh tri --no-cpu --no-sub --no-weak > $out/${tag}_bench_n1_tri.json 2> $out/${tag}_bench_n1_tri.err
python tools/poisson_sweep.py --sizes 1,4,16,64 --precond amg > $out/${tag}_sweep_const_amg_n1.jsonl 2> $out/${tag}_sweep_const.err
python tools/poisson_sweep.py --sizes 1,4,16 --precond amg --variable > $out/${tag}_sweep_var_amg_n1.jsonl 2> $out/${tag}_sweep_var.err
python tools/bubble_case.py > $out/${tag}_bubble_4M_amg.json 2> $out/${tag}_bubble.err
python tools/cylinder_case.py --cells 16e6 --steps 5 --warmup 2 > $out/${tag}_cylinder_16M_n1.json 2> $out/${tag}_cylinder.err
# launch list of the default configuration (numbers printed under ncu are never bench values)
timeout 600 $NCU --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file $out/${tag}_launches_default.csv \
  python bench.py --steps 2 --warmup 1 $quick > $out/${tag}_ncu_bench.log 2>&1
# full-set captures: the fused assembly kernels; then 60 consecutive multigrid / Krylov launches inside a timed step
timeout 600 $NCU --set full --clock-control none --import-source on -k regex:'k_momentum_fused|k_pressure_fused' -c 4 \
  -o $out/${tag}_prof_assembly python bench.py --steps 1 --warmup 1 $quick > $out/${tag}_ncu_full_a.log 2>&1
timeout 900 $NCU --set full --clock-control none --import-source on -k regex:'k_amg_tail|k_amg_spmv|k_amg_scale|k_spmv|k_update' \
  --launch-skip 500 -c 60 -o $out/${tag}_prof_cycle python bench.py --steps 1 --warmup 1 $quick > $out/${tag}_ncu_full_c.log 2>&1
for n in assembly cycle; do
  $NCU -i $out/${tag}_prof_$n.ncu-rep --page raw --csv > $out/${tag}_prof_${n}_raw.csv 2>/dev/null
  rm -f $out/${tag}_prof_$n.ncu-rep
done
ls -la $out | tail -20
