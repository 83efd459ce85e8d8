#!/bin/bash
# Same-box A/B of prebuilt libraries ab_tmp/lib_<variant>.so (alternating runs; noise is ~0.1 % on one box).
# usage: tools/ab_bench.sh <reps> <variant> [<variant> ...]
L=gradient-boosted-normalizing-flows_b200/libgbnf_b200.so
reps=$1; shift
for rep in $(seq $reps); do
  for v in "$@"; do
    cp ab_tmp/lib_$v.so $L; touch $L
    echo -n "$v: "
    timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu ${AB_ARGS} 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']/1e6,3), round(d['roofline']['frac'],4), d['clocks']['sm_mhz'], d['clocks']['power_w_max'])"
  done
done
