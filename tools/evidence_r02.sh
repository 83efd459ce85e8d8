#!/bin/bash
# Round-2 evidence run (one GPU): launch lists, ncu full captures of the dominant kernels, the HBM-bound mixture / weight kernels
# under ncu, a sustained (> 2 s) bench line and the bench lines of the other BASELINE configurations.  Outputs -> gpurun_out/.
set -x
O=gpurun_out
NCU="ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file $O/r02_launches_cfg3.csv python bench.py --steps 2 --warmup 3 --no-cpu > $O/r02_ncu_a.log 2>&1
$NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file $O/r02_launches_cfg5.csv python bench.py --config cfg5_bsds300 --steps 2 --warmup 3 --no-cpu > $O/r02_ncu_b.log 2>&1
$NCU --set full --kernel-name regex:coupling_tc2 -c 1 -o $O/r02_tc2_full -f python bench.py --steps 1 --warmup 3 --no-cpu > $O/r02_ncu_c.log 2>&1
$NCU --set full --kernel-name regex:coupling_tc3 -c 1 -o $O/r02_tc3_full -f python bench.py --config cfg5_bsds300 --steps 1 --warmup 3 --no-cpu > $O/r02_ncu_d.log 2>&1
$NCU --set full --kernel-name 'regex:mixture_lse|softmax_stats|weight_apply|weight_renorm' -s 12 -c 8 -o $O/r02_mixture_full -f python tools/mixture_bench.py 33554432 > $O/r02_ncu_e.log 2>&1
$NCU --set full --kernel-name regex:coupling_fp32 -c 1 -o $O/r02_fp32_full -f python bench.py --mode fp32 --batch 8192 --rows 65536 --steps 1 --warmup 3 --no-cpu > $O/r02_ncu_f.log 2>&1
$NCU --set full --kernel-name 'regex:train_bwd_rows|train_wgrad' -c 2 -o $O/r02_train_full -f python tools/train_step_bench.py --batches 512 --steps 2 > $O/r02_ncu_g.log 2>&1
# the .ncu-rep files are too large to travel back (64 MiB cap on gpurun_out): keep their raw-metric pages as CSV
for r in r02_tc2_full r02_tc3_full r02_mixture_full r02_fp32_full r02_train_full; do
  ncu -i $O/$r.ncu-rep --page raw --csv > $O/$r.raw.csv 2>/dev/null
  rm -f $O/$r.ncu-rep
done
python bench.py --steps 1500 --warmup 10 --no-cpu > $O/r02_bench_cfg3_sustained.json 2>/dev/null
python bench.py --mode f16 --steps 200 --warmup 10 --no-cpu > $O/r02_bench_cfg3_f16.json 2>/dev/null
for c in cfg1_toy cfg2_power cfg4_hepmass; do python bench.py --config $c --steps 100 --warmup 10 > $O/r02_bench_$c.json 2>/dev/null; done
python bench.py --config cfg5_bsds300 --steps 60 --warmup 5 > $O/r02_bench_cfg5_bsds300.json 2>/dev/null
python tools/mixture_bench.py 2097152 > $O/r02_mixture_bench_2M.json 2>/dev/null
python tools/mixture_bench.py 33554432 > $O/r02_mixture_bench_32M.json 2>/dev/null
python tools/train_step_bench.py > $O/r02_train_step.json 2>/dev/null
ls -la $O | tail -30
