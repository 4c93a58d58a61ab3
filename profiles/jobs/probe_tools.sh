# Round 2, first job: which reference toolchains exist on the GPU box.
mkdir -p gpurun_out
{
  echo "== tool probe on the GPU box ($(date -u +%FT%TZ)) =="
  for t in octave octave-cli matlab ghdl nvc iverilog vvp verilator vivado xsim xvhdl vsim vcom go javac node; do
    p=$(command -v $t 2>/dev/null); echo "$t: ${p:-absent}"
  done
  echo "== host =="; nproc; lscpu | egrep 'Model name|Socket|Core|Thread|NUMA|MHz' ; free -g | head -2
  echo "== gpu =="; nvidia-smi --query-gpu=name,pci.bus_id,pcie.link.gen.current,pcie.link.width.current,clocks.max.sm,memory.total --format=csv
  nvidia-smi topo -m 2>/dev/null | head -20
} > gpurun_out/tool_probe_r02.txt 2>&1
cat gpurun_out/tool_probe_r02.txt
