import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
import zignal_b200 as zg, flowz_oracle as fo
C, T = int(sys.argv[1]), int(sys.argv[2])
x = fo.noise(C, T, seed=1)
expr = fo.biquad_cascade(4)
plan = zg.compile(expr).plan(channels=C, lanes_per_channel=4)
y = plan.process([zg.to_block(x)])[0]
torch.cuda.synchronize()
i = plan.info()
print("geometry", i.threads_per_cta, i.stages, i.boxes, i.smem_bytes)
ref = fo.COracle(expr, C).process([x])[0]
print("equal", np.array_equal(y.cpu().numpy(), ref))
