#!/usr/bin/env python3
"""Timeline of the pipelined alignGraphs batch of the bench (one emission -> fill -> traceback pipeline per size class):
   python tools/timeline_hp1.py   (prints the begin / duration of every timed section of one run; EMISSION=1: tensor)"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
from prographmsa_b200 import model, synth, graphalign, progressive, _lib
from prographmsa_b200.options import Options
opts = Options(mldist_flag=True)
mf = model.getDefault(opts)
seqs, root = synth.simulate_family(256, 400, mf, seed=2)
tree = progressive.from_synth(root)
ctx = _lib.default_context()
jobs = []
def batch(js, o):
    jobs.extend(js)
    return graphalign.alignGraphsBatch(js, o, ctx=ctx)
progressive.progressive_alignment(seqs, tree, mf, opts, batch_align_fn=batch)
ctx.enable_timing(True)
gb = graphalign.GraphBatch(jobs * 4, opts, int(os.environ.get("EMISSION", "0")), ctx)
for it in range(3):
    gb.run(); ctx.sync()
os.environ["PGM_TIMELINE"] = "1"
gb.run(); ctx.sync(); ctx.timing()
