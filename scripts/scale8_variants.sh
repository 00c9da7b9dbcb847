#!/bin/bash
# 8-GPU variants of the bench step (one box, `gpurun --gpus 8`): what phase A depends on.  Results of round 2:
# profiles/r02_scaling/r02c_*, r02d_*, r02f_* (side stream, priorities, waves of slabs, sharing vs waiting: all within 0.7 %)
run() { tag=$1; shift; env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $((29600+RANDOM%300)) bench.py --gpus 8 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/scale8_$tag.json 2> gpurun_out/scale8_$tag.err; python - <<PY
import json
d=json.loads(open('gpurun_out/scale8_$tag.json').read().strip().splitlines()[-1])
print('$tag', round(d['value'],2), round(d['ms_per_step'],3), {k:round(v,3) for k,v in d['phase_ms'].items()}, 'e2e', round(d['e2e']['value'],1), d['check']['force_checksum'], d['check']['ranks_identical'])
PY
}
run default A=1
run noside MDB_PEER_NO_SIDE=1
run share MDB_PEER_SHARE=1
run prio MDB_PEER_SIDE_PRIO=1
run waves4 MDB_SFAC_WAVES=4
