import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from multioptpy_b200 import ops, _lib
n = int(sys.argv[1]); B = int(sys.argv[2]); cl = int(sys.argv[3])
rng = np.random.default_rng(0)
A = rng.standard_normal((B, n, n)); A = 0.5 * (A + A.transpose(0, 2, 1))
At = torch.from_numpy(A).cuda()
_lib.load().mop_priv_large_cluster(cl)
ops.eigh(At, "large"); torch.cuda.synchronize()
