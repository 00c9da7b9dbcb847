run() { n=$1; w=$2; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+RANDOM%300)) bench.py --gpus $n --steps 10 --warmup 3 --workload $w > gpurun_out/r02_${w}_${n}gpu.json 2> gpurun_out/r02_${w}_${n}gpu.err; tail -c 200 gpurun_out/r02_${w}_${n}gpu.json | head -c 100; echo " <- $w x$n"; }
run 8 tip4p13; run 8 tip4p16; run 8 quartz48; run 8 mgcl2_7; run 8 tip4p10
run 4 tip4p13; run 4 tip4p16; run 2 tip4p13; run 2 tip4p16
