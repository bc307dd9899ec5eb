import sys; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
import numpy as np, oracle
from depthhead_b200 import synth
from test_gpu_configs import _ramp_forest
arr=_ramp_forest(); js = synth.forest_to_json(arr, stepwidth=5, meanshift_iterations=30)
of = oracle.OracleForest.from_json(js)
d = np.full((90, 90), 1000, np.uint16)
tr = of.predict(d, synth.KINECT_K, mode=oracle.MODE_SAT, keep=True)
print(len(tr.mid_keys), tr.mid_keys[:,0].min(), tr.mid_keys[:,0].max(), tr.gate, tr.seed_mid, tr.ms_mid[:5])
x0=int(tr.mid_keys[:,0].min()); y0,z0=int(tr.mid_keys[0,1]),int(tr.mid_keys[0,2])
for start in (x0-6,x0+2):
    t2=of.predict(d, synth.KINECT_K, [float(start),float(y0),float(z0)], [0.15,-0.2,0.05], mode=oracle.MODE_SAT, keep=True)
    print(start, t2.ms_mid[:12].tolist(), t2.mid_point)
