#!/bin/bash
# compute-sanitizer over the kernels added in round 2 (small systems): peer layer with several ranks on one device, device group,
# NVE step, force_both with a filler grid.  usage (one GPU): bash scripts/r02_sanitize.sh > gpurun_out/r02_sanitizer.txt
SEL='tests/test_gpu_overlap.py tests/test_gpu_md.py tests/test_gpu_peer.py'
K='(test_force_both_equals and (tip4p or slab) and not tip4p_2) or test_substeps or test_h0 or test_resident_and_uploading or test_ranks_sharing or test_rdf_pass_on_a_device_group'
for tool in memcheck initcheck synccheck; do
  echo "== $tool"
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 0 python -m pytest $SEL -q -x -k "$K" 2>&1 | grep -E "passed|failed|ERROR SUMMARY|Error|error" | tail -6
done
