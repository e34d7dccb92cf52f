#!/bin/bash
# round 2, call ac (8 GPUs): the bench line at N=8: value and e2e scaling, the concurrent-H2D ceiling of the node
OUT=gpurun_out/${1:-r02ac}
mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo.txt 2>&1
nproc > $OUT/nproc.txt; free -g >> $OUT/nproc.txt; lscpu | grep -i "numa\|model name\|socket" >> $OUT/nproc.txt
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29557 bench.py --gpus 8 --steps 5 --warmup 3 ) > $OUT/bench_8gpu.json 2> $OUT/bench_8gpu.err
python - <<PY
import json
b = json.load(open("$OUT/bench_8gpu.json"))
print("value %.4e ms/step %.3f frac %.3f share %.3f" % (b["value"], b["ms_per_step"], b["roofline"]["frac"], b["roofline"]["kernel_share_of_step"]))
print(json.dumps(b["e2e"], indent=0))
print(b["clocks"])
PY
tail -3 $OUT/bench_8gpu.err; cat $OUT/nproc.txt
