#!/bin/bash
# Round-2 evidence, second pass (after the two-chain kernel and the instruction-footprint work): launch list + ncu full capture of
# coupling_tc4_kernel (cfg3) and of the slimmed coupling_tc3_kernel (cfg5), bench lines of every configuration, sustained line.
set -x
O=gpurun_out
NCU="ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file $O/r02b_launches_cfg3.csv python bench.py --steps 2 --warmup 3 --no-cpu > $O/r02b_ncu_a.log 2>&1
$NCU --set full --import-source on --kernel-name regex:coupling_tc4 -c 1 -o $O/r02b_tc4_full -f python bench.py --steps 1 --warmup 3 --no-cpu > $O/r02b_ncu_c.log 2>&1
$NCU --set full --kernel-name regex:coupling_tc3 -c 1 -o $O/r02b_tc3_full -f python bench.py --config cfg5_bsds300 --steps 1 --warmup 3 --no-cpu > $O/r02b_ncu_d.log 2>&1
for r in r02b_tc4_full r02b_tc3_full; do
  ncu -i $O/$r.ncu-rep --page raw --csv > $O/$r.raw.csv 2>/dev/null
done
ncu -i $O/r02b_tc4_full.ncu-rep --page source --csv > $O/r02b_tc4_source.csv 2>/dev/null
gzip -f $O/r02b_tc4_source.csv
rm -f $O/*.ncu-rep
python bench.py --steps 200 --warmup 10 > $O/r02b_bench_cfg3.json 2>/dev/null
python bench.py --steps 1500 --warmup 10 --no-cpu > $O/r02b_bench_cfg3_sustained.json 2>/dev/null
python bench.py --mode f16 --steps 200 --warmup 10 --no-cpu > $O/r02b_bench_cfg3_f16.json 2>/dev/null
GBNF_TC4=0 python bench.py --steps 200 --warmup 10 --no-cpu > $O/r02b_bench_cfg3_single_chain.json 2>/dev/null
for c in cfg1_toy cfg2_power cfg4_hepmass; do python bench.py --config $c --steps 100 --warmup 10 --no-cpu > $O/r02b_bench_$c.json 2>/dev/null; done
python bench.py --config cfg5_bsds300 --steps 60 --warmup 5 --no-cpu > $O/r02b_bench_cfg5_bsds300.json 2>/dev/null
python tools/tc4_profile.py > $O/r02b_tc4_profile.txt 2>&1
GBNF_GRID=74 python tools/tc4_profile.py cfg3_miniboone f16fast 32768 > $O/r02b_tc4_profile_half_grid.txt 2>&1
ls -la $O | tail -30
du -sh $O
