#!/bin/bash
# what the GPU box looks like from the host side: cores, memory, NUMA placement of the GPUs
mkdir -p gpurun_out
{
echo "nproc: $(nproc)  affinity: $(python -c 'import os;print(len(os.sched_getaffinity(0)))')"
free -g | head -3
echo "--- numa nodes"
for n in /sys/devices/system/node/node*; do echo "$n cpus=$(cat $n/cpulist) $(grep MemTotal $n/meminfo | awk '{print $4/1048576 " GB"}')"; done
echo "--- gpus"
nvidia-smi --query-gpu=index,pci.bus_id,name,memory.total --format=csv,noheader
for d in /sys/bus/pci/devices/*; do if [ "$(cat $d/vendor 2>/dev/null)" = "0x10de" ] && [ "$(cat $d/class 2>/dev/null | cut -c1-4)" = "0x03" ]; then echo "$(basename $d) numa_node=$(cat $d/numa_node) local_cpus=$(cat $d/local_cpulist)"; fi; done
nvidia-smi topo -m 2>&1 | head -30
echo "--- cgroup"
cat /sys/fs/cgroup/cpuset.cpus.effective 2>/dev/null; cat /sys/fs/cgroup/cpuset.mems.effective 2>/dev/null; cat /sys/fs/cgroup/memory.max 2>/dev/null
ulimit -l
df -h /tmp /dev/shm | tail -2
} > gpurun_out/probe.txt 2>&1
cat gpurun_out/probe.txt
