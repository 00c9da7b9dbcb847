run() { tag=$1; shift; env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((29600+RANDOM%300)) bench.py --gpus 8 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02f_8gpu_$tag.json 2> gpurun_out/r02f_8gpu_$tag.err; python - <<PY
import json
d=json.loads(open('gpurun_out/r02f_8gpu_$tag.json').read().strip().splitlines()[-1])
print('$tag', round(d['value'],2), round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['phase_ms'].items()}, 'e2e', round(d['e2e']['value'],1), d['check']['force_checksum'], d['check']['ranks_identical'])
PY
}
run default A=1
