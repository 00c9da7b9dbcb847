python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for w in tip4p5 mgcl2_7 quartz48 tip4p13 tip4p16; do python bench.py --steps 10 --warmup 3 --workload $w --no-cpu-baseline > gpurun_out/r02_${w}_1gpu.json 2> gpurun_out/r02_${w}_1gpu.err; tail -c 120 gpurun_out/r02_${w}_1gpu.json; echo " <- $w"; done
python bench.py --steps 20 --warmup 3 > gpurun_out/r02_tip4p10_1gpu.json 2> gpurun_out/r02_tip4p10_1gpu.err; tail -c 300 gpurun_out/r02_tip4p10_1gpu.json
