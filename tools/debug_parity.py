import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np
import parity as P
import quicksand_b200 as qs
np.set_printoptions(precision=17, linewidth=200)
prog = qs.programs.indep([("gaussian", 0.1, 0.5)])
gpu, cpu = P.run_both(prog, qs.HMCKernel(numSteps=1), chains=4, iters=6, seed=20261019, leap=True)
print("init g", gpu["init"].ravel()); print("init c", cpu["init"].ravel())
print("step g", gpu["step"]); print("step c", cpu["step"])
print("dH g", gpu["dH"]); print("dH c", cpu["dH"])
print("acc g", gpu["accept"]); print("acc c", cpu["accept"])
print("state g", gpu["state"][:, :, 0]); print("state c", cpu["state"][:, :, 0])
print("stats", gpu["stats"], cpu["stats"])
