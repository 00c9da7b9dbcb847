# per-kernel device times of one perf_probe step (ncu, gpu__time_duration): usage  bash scripts/ktime.sh [regex] [n]
RX=${1:-"k_sfac_mma|k_kforce_mma|k_ktables|k_pair_tiled"}
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"$RX" -s ${3:-8} -c ${4:-8} --csv python scripts/perf_probe.py ${2:-10} 2 2>/dev/null | grep -E '^"[0-9]' | python -c "
import csv,sys
for r in csv.reader(sys.stdin): print('  ',r[4][:60], float(r[-1].replace(',',''))/1e6, 'ms')
"
