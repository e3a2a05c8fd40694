#!/usr/bin/env python3
"""Host-side overhead of pgm_nw_ml_pair_range_dev: wall time per call vs the sum of its kernels, for a 1/W shard."""
import sys, time, ctypes as C
import numpy as np
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import torch
import bench
from prographmsa_b200 import _lib, alphabet, distance, sharding

W = int(sys.argv[1]) if len(sys.argv) > 1 else 8
opts, mf, seqs, names, _ = bench.make_family()
AA = alphabet.AA
codes, offsets = distance.pack_sequences([AA.nw_codes(seqs[k]) for k in names])
tab, go, ge = distance.nw_table(AA)
keep = []
prm = distance.make_ml_params(mf, opts, keep)
N = len(names)
pb, pe = sharding.pair_shard(N, 0, W)
ctx = _lib.Context(0); ctx.enable_timing(True)
lib = _lib.lib()
d_codes = torch.from_numpy(codes).cuda()
d_out = torch.zeros(2, pe - pb, dtype=torch.float64, device="cuda")
vp = lambda t: C.c_void_p(t.data_ptr())
def call():
    _lib.check(lib.pgm_nw_ml_pair_range_dev(ctx.handle, C.byref(prm), vp(d_codes), offsets.ctypes.data_as(C.c_void_p), N,
                                            tab.ctypes.data_as(C.c_void_p), go, ge, C.c_int64(pb), C.c_int64(pe),
                                            vp(d_out[0]), vp(d_out[1]), None, None, None))
for _ in range(5): call()
torch.cuda.synchronize()
ws = []; ks = []
for _ in range(20):
    t0 = time.perf_counter(); call(); ws.append(1e3 * (time.perf_counter() - t0))
    tm = ctx.timing(); ks.append(tm["nw_forward_ms"] + tm["nw_walk_ms"] + tm["ml_ms"])
print("shard 1/%d: %d pairs; wall per call %.3f ms (min %.3f), kernels %.3f ms, host+gaps %.3f ms" %
      (W, pe - pb, np.median(ws), min(ws), np.median(ks), np.median(ws) - np.median(ks)))
