#!/bin/bash
# Round-2 evidence, final pass (round-end build): smoke, bench lines of every configuration, the launch list of the timed steps only
# (kernel-name filter: the model set-up launches torch kernels), ncu capture of the interleaved s / t kernel.
set -x
O=gpurun_out
NCU="ncu --clock-control none"
python __graft_entry__.py smoke > $O/r02c_smoke.log 2>&1; tail -3 $O/r02c_smoke.log
$NCU --metrics gpu__time_duration.sum --kernel-name 'regex:coupling_|softmax_stats|weight_apply|weight_renorm|scan_|resample_|gather_rows' -c 60 --csv --log-file $O/r02c_launches_cfg3_steps.csv python bench.py --steps 6 --warmup 3 --no-cpu > $O/r02c_ncu_a.log 2>&1
$NCU --set full --kernel-name regex:coupling_tc5 -c 1 -o $O/r02c_tc5_full -f python bench.py --config cfg2_power --steps 1 --warmup 3 --no-cpu > $O/r02c_ncu_b.log 2>&1
ncu -i $O/r02c_tc5_full.ncu-rep --page raw --csv > $O/r02c_tc5_full.raw.csv 2>/dev/null
rm -f $O/*.ncu-rep
python bench.py --steps 200 --warmup 10 > $O/r02c_bench_cfg3.json 2>/dev/null
python bench.py --impl reference --steps 3 --warmup 1 > $O/r02c_bench_cfg3_reference_arm.json 2>/dev/null
for c in cfg1_toy cfg2_power cfg4_hepmass; do python bench.py --config $c --steps 100 --warmup 10 --no-cpu > $O/r02c_bench_$c.json 2>/dev/null; done
python bench.py --config cfg5_bsds300 --steps 60 --warmup 5 --no-cpu > $O/r02c_bench_cfg5_bsds300.json 2>/dev/null
python tools/tc4_profile.py > $O/r02c_tc4_profile.txt 2>&1
python tools/train_step_bench.py > $O/r02c_train_step.json 2>/dev/null
ls -la $O | grep r02c
