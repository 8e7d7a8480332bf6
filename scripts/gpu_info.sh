#!/bin/bash
mkdir -p gpurun_out
{
echo "== thp"; cat /sys/kernel/mm/transparent_hugepage/enabled /sys/kernel/mm/transparent_hugepage/shmem_enabled /sys/kernel/mm/transparent_hugepage/defrag 2>&1
echo "== hugepages"; grep -i huge /proc/meminfo
echo "== numa"; ls /sys/devices/system/node/ 2>&1 | head; lscpu | grep -i -E "numa|socket|model name|^cpu\(s\)|thread"
echo "== affinity"; python -c "import os;print(len(os.sched_getaffinity(0)), sorted(os.sched_getaffinity(0))[:8])"
echo "== topo"; nvidia-smi topo -m 2>&1 | head -30
echo "== pci numa"; for d in /sys/bus/pci/devices/*; do if [ -f $d/vendor ] && grep -q 0x10de $d/vendor; then echo $d $(cat $d/numa_node) $(cat $d/current_link_speed 2>/dev/null) $(cat $d/current_link_width 2>/dev/null); fi; done | head -20
echo "== shm"; df -h /dev/shm; mount | grep shm
echo "== iommu"; ls /sys/class/iommu 2>&1 | head -3; cat /proc/cmdline
echo "== mem"; free -g
} > gpurun_out/box_info.txt 2>&1
cat gpurun_out/box_info.txt
