#!/bin/bash
# round-2 ncu evidence (ONE GPU, under gpurun): launch list of a bench step + --set full of the pair instantiations that
# carry the non-TIP4P configs (Buckingham fused pass, MCY potential pass) and of the TIP4P kernels.  The reports are
# summarised on the box (gpurun_out/ is capped at 64 MiB) and deleted.
mkdir -p gpurun_out
M=gpu__time_duration.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum
ncu --metrics $M --clock-control none -s 100 -c 120 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02_launches_bench.log 2>&1
for W in quartz_48 mgcl2_7 tip4p_10; do
  ncu --set full --clock-control none --import-source on -k regex:"k_pair_tiled|k_sfac_mma|k_kforce_mma" -s 0 -c 4 \
      -o /tmp/r02_full_$W -f python scripts/perf_probe.py $W 1 > gpurun_out/r02_full_$W.log 2>&1
  python scripts/ncu_summary.py /tmp/r02_full_$W.ncu-rep > gpurun_out/r02_full_$W.summary.md 2>&1
  python scripts/ncu_hot.py /tmp/r02_full_$W.ncu-rep k_pair_tiled 0.3 --list > gpurun_out/r02_full_$W.hot_pair.txt 2>&1
  ncu -i /tmp/r02_full_$W.ncu-rep --page raw --csv > gpurun_out/r02_full_$W.raw.csv 2>/dev/null
  tail -2 gpurun_out/r02_full_$W.log
done
python scripts/ncu_hot.py /tmp/r02_full_tip4p_10.ncu-rep k_sfac_mma 0.3 > gpurun_out/r02_full_tip4p_10.hot_sfac.txt 2>&1
python scripts/ncu_hot.py /tmp/r02_full_tip4p_10.ncu-rep k_kforce_mma 0.3 > gpurun_out/r02_full_tip4p_10.hot_kforce.txt 2>&1
du -sh gpurun_out
