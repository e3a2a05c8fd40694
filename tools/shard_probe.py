import os, sys, time, ctypes as C
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import bench
from prographmsa_b200 import alphabet, distance, sharding, _lib
opts, mf, seqs, names, _ = bench.make_family()
AA = alphabet.AA
code_list = [AA.nw_codes(seqs[k]) for k in names]
codes, offsets = distance.pack_sequences(code_list)
tab, go, ge = distance.nw_table(AA)
keep = []
prm = distance.make_ml_params(mf, opts, keep)
N = bench.N_SEQ
ctx = _lib.Context(0); ctx.enable_timing(True)
lib = _lib.lib()
d_codes = torch.from_numpy(codes).cuda()
for world in (1, 2, 4, 8):
    per = sharding.shard_size(N, world)
    d_out = torch.zeros(2, per, dtype=torch.float64, device="cuda")
    for rank in (0, world - 1):
        pb, pe = sharding.pair_shard(N, rank, world)
        def step():
            _lib.check(lib.pgm_nw_ml_pair_range_dev(ctx.handle, C.byref(prm), C.c_void_p(d_codes.data_ptr()), offsets.ctypes.data_as(C.c_void_p), N,
                                                    tab.ctypes.data_as(C.c_void_p), go, ge, C.c_int64(pb), C.c_int64(pe),
                                                    C.c_void_p(d_out[0].data_ptr()), C.c_void_p(d_out[1].data_ptr()), None, None, None))
        for _ in range(3): step()
        torch.cuda.synchronize()
        ts = []; ks = []
        for _ in range(10):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); t0 = time.perf_counter()
            a.record(); step(); b.record(); torch.cuda.synchronize()
            ts.append((time.perf_counter() - t0) * 1e3)
            t = ctx.timing(); ks.append((t["nw_forward_ms"], t["nw_walk_ms"], t["ml_ms"]))
        k = np.mean(ks, axis=0)
        print("world %d rank %d: pairs %6d  wall %.3f ms  kernels fwd %.3f walk %.3f ml %.3f sum %.3f" % (world, rank, pe - pb, np.mean(ts), k[0], k[1], k[2], k.sum()))
