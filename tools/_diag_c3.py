import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from depthhead_b200 import Context, HoughPrediction, IntrinsicMatrix, synth
K = IntrinsicMatrix.default_kinect_intrinsic()
ctx = Context(0)
ctx.set_stream(torch.cuda.current_stream().cuda_stream)
frames = synth.make_frames(16, seed=8)
arr = synth.make_forest(seed=9, n_trees=50, max_depth=20, stop_prob=0.3)
hp = HoughPrediction.from_arrays(arr, stepwidth=1)
dev = torch.from_numpy(frames.view(np.int16)).cuda()
n, h, w = frames.shape
step = lambda: hp.predict_batch(None, K, ctx=ctx, device_ptr=dev.data_ptr(), n=n, w=w, h=h)
step(); step()
ctx.enable_stage_timing(True)
step()
st = ctx.stage_ms(); cnt = ctx.counters()
print(os.environ.get("TAG"), {k: round(v / n, 4) for k, v in st.items()}, {k: cnt[k] for k in ("meanshift_iters", "cube_rebuilds", "gate_patches", "centre_votes", "rot_votes")})
