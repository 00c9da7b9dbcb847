#!/bin/bash
# ncu evidence for the eval_forces() kernels (run under gpurun, ONE GPU):
#   launch list (times + DRAM bytes)  -> gpurun_out/evalf_launches.csv   (table: python scripts/ncu_launches.py <csv>)
#   --set full of k_mol_frame / k_make_sites / k_dipole_partial -> gpurun_out/evalf_full.ncu-rep
# usage: gpurun --timeout 400 -- 'bash scripts/evalf_ncu.sh [n=10]'
set -e
N=${1:-10}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    -k regex:"k_make_sites|k_dipole|k_mol_frame|k_eval_finish" -c 14 --csv --log-file gpurun_out/evalf_launches.csv \
    python scripts/evalf_probe.py "$N" 1 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_mol_frame|k_make_sites|k_dipole_partial" -s 4 -c 4 \
    -o gpurun_out/evalf_full -f python scripts/evalf_probe.py "$N" 1 > gpurun_out/evalf_full.log 2>&1
python scripts/ncu_launches.py gpurun_out/evalf_launches.csv | head -20
