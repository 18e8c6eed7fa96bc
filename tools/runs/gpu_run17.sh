#!/bin/bash
# e2e (host buffers): number of chunks the block is streamed in; new tests
mkdir -p gpurun_out
for c in 8 16 32 64; do
  ZG_TUNE_HOST_CHUNKS=$c timeout 300 python bench.py --no-also --no-cpu --steps 5 --warmup 3 --e2e-steps 5 > gpurun_out/e2e_$c.log 2>&1
  python - <<PY
import json
d=json.loads(open("gpurun_out/e2e_$c.log").read().strip().splitlines()[-1])
print("chunks", $c, "e2e", round(d["e2e"]["value"]), "Msamples/s", round(d["e2e"]["ms_per_step"],2), "ms")
PY
done
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "million or sample_rate or host" 2>&1 | tail -3
