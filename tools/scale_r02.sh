#!/bin/bash
# Multi-GPU evidence on one 8 x B200 box: the torchrun parity check at 2 and 8 ranks, batch-parallel weak scaling at 1 / 2 / 4 / 8
# (library-owned peer-memory exchange; --nccl = torch.distributed collectives for the A/B), component-parallel cfg4 / cfg3 at 8.
O=gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for n in 2 8; do
  timeout 300 $TR --nproc-per-node $n --master-port 2951$n tests/multi_gpu_check.py > $O/r02b_multi_gpu_check_$n.log 2>&1; tail -2 $O/r02b_multi_gpu_check_$n.log
done
python bench.py --gpus 1 --steps 200 --warmup 10 --no-cpu > $O/r02b_scale_1gpu.json 2>/dev/null
for n in 2 4 8; do
  timeout 300 $TR --nproc-per-node $n --master-port 2952$n bench.py --gpus $n --steps 200 --warmup 10 > $O/r02b_scale_${n}gpu.json 2>$O/r02b_scale_${n}gpu.err
done
timeout 300 $TR --nproc-per-node 8 --master-port 29531 bench.py --gpus 8 --steps 200 --warmup 10 --nccl > $O/r02b_scale_8gpu_nccl.json 2>/dev/null
timeout 300 $TR --nproc-per-node 8 --master-port 29532 bench.py --gpus 8 --config cfg4_hepmass --parallel component --steps 100 --warmup 10 > $O/r02b_cfg4_component_parallel_8gpu.json 2>$O/r02b_cfg4_cp.err
timeout 300 $TR --nproc-per-node 8 --master-port 29533 bench.py --gpus 8 --config cfg3_miniboone --parallel component --steps 200 --warmup 10 > $O/r02b_cfg3_component_parallel_8gpu.json 2>/dev/null
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r02b_scale_*gpu*.json") + glob.glob("gpurun_out/r02b_cfg*_component_parallel_8gpu.json")):
    try:
        d = json.load(open(f))
        print(f.split("/")[-1], d["n_gpus"], round(d["value"] / 1e6, 2), "M/s e2e", round(d["e2e"]["value"] / 1e6, 2), "frac", round(d["roofline"]["frac"], 3),
              "weights_ms", round(d["roofline"]["weights_ms"], 4), d.get("parity_check", {}).get("passed"), d["clocks"]["sm_mhz"])
    except Exception as e:
        print(f, "ERR", e)
PY
