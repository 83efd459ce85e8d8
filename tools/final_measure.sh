#!/bin/bash
# Round-end measurement pass on one B200: bench lines of every configuration, the timeline and the ncu captures that
# profiles/ summarises.  Outputs go to gpurun_out/final/.
O=gpurun_out/final; mkdir -p $O
python bench.py --steps 200 --warmup 10 2>/dev/null | tail -1 > $O/bench_f16fast.json
python bench.py --steps 200 --warmup 10 --mode f16 --no-cpu 2>/dev/null | tail -1 > $O/bench_f16.json
python bench.py --steps 20 --warmup 3 --mode fp32 --no-cpu 2>/dev/null | tail -1 > $O/bench_fp32.json
python bench.py --steps 50 --warmup 5 --config cfg4_hepmass --no-cpu 2>/dev/null | tail -1 > $O/bench_cfg4_hepmass.json
python bench.py --steps 200 --warmup 10 --config cfg2_power --no-cpu 2>/dev/null | tail -1 > $O/bench_cfg2_power.json
python bench.py --steps 200 --warmup 10 --config cfg1_toy --no-cpu 2>/dev/null | tail -1 > $O/bench_cfg1_toy.json
python bench.py --steps 5 --warmup 3 --config cfg5_bsds300 --no-cpu 2>/dev/null | tail -1 > $O/bench_cfg5_bsds300.json
python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 > $O/bench_reference.json
python tools/tc_trace.py cfg3_miniboone f16fast > $O/timeline.txt 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"coupling|softmax_stats|weight_apply|weight_renorm|zero_double" -c 200 --csv --log-file $O/launches.csv python bench.py --steps 5 --warmup 3 --no-cpu > $O/ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:coupling_tc2 -c 1 -o $O/tc2_full -f python bench.py --steps 1 --warmup 3 --no-cpu > $O/ncu_full.log 2>&1
ncu -i $O/tc2_full.ncu-rep --page raw --csv > $O/tc2_full_raw.csv 2>/dev/null
ls -la $O
for f in $O/bench_*.json; do echo "$f: $(cut -c1-200 $f)"; done
