import os, sys, subprocess, json
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from helpers import Problem
pb = Problem(700, 24, 4, 4, seed=81, s=0.9)
ch = pb.chain(L=5, c_i=None, compact_init=True, seed=82)
ch.step(1100)
s = ch.snapshot()
print(json.dumps({"K": s["K"], "ll": s["loglikelihood"], "csum": int(np.sum(s["c_i"] * np.arange(pb.n))), "sig": float(s["sigmas"].sum())}))
