mkdir -p gpurun_out
{
nvidia-smi topo -m
lscpu | grep -E "NUMA|Socket|^CPU\(s\)|Model name"
for g in 0 1; do bdf=$(nvidia-smi --query-gpu=pci.bus_id --format=csv,noheader -i $g | tr 'A-Z' 'a-z' | sed 's/^0000//'); echo "gpu $g $bdf numa=$(cat /sys/bus/pci/devices/$bdf/numa_node 2>/dev/null) cpus=$(cat /sys/bus/pci/devices/$bdf/local_cpulist 2>/dev/null)"; done
python -c "import os; print('affinity', sorted(os.sched_getaffinity(0)))"
cat /sys/devices/system/node/online 2>/dev/null
free -g | head -2
} > gpurun_out/numa_probe.txt 2>&1
cat gpurun_out/numa_probe.txt
