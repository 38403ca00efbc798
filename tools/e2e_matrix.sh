#!/bin/bash
# dev: e2e leg of the bench at N GPUs under a few host-side settings
N=${1:-4}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544"
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
for g in 0 1 2 3 4 5 6 7; do b=$(nvidia-smi -i $g --query-gpu=pci.bus_id --format=csv,noheader 2>/dev/null | tr 'A-Z' 'a-z' | sed 's/^0000//'); [ -n "$b" ] && echo "gpu $g $b numa $(cat /sys/bus/pci/devices/$b/numa_node 2>/dev/null)"; done
lscpu | grep -i "numa\|socket\|^CPU(s)" | head -8
for e in "MRT_NUMA_BIND=1" "MRT_NUMA_BIND=1 MRT_HP_ZEROCOPY=0"; do
  env $e timeout 300 $TR bench.py --gpus $N --steps 10 --warmup 3 --no-cpu 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); e=d['e2e']; print('$e N=$N', 'e2e ms/step %.4f  value %.1f G/s  numa %s  device step %.4f' % (e['ms_per_step'], e['value']/1e9, e.get('numa_node_rank0'), d['ms_per_step']))"
done
