import sys; sys.path.insert(0, "/root/repo")
import torch
from multioptpy_b200 import _lib
lib = _lib.load()
out = torch.zeros(8, dtype=torch.float64, device="cuda")
lib.mop_priv_latency(out.data_ptr(), None); torch.cuda.synchronize()
print("DFMA, DADD, DMUL, LDS chase, SHFL64, fast_rcp, sqrt, div:", [round(float(v), 1) for v in out.cpu()])
