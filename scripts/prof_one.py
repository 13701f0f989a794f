"""Launch one local-steps kernel (for ncu): python scripts/prof_one.py [case] [layout_hint] [n_steps]"""
import sys
sys.path.insert(0, "/root/repo")
import numpy as np
import torch
from scripts.sweep_local import run
from flowmc_b200 import targets as T

case = sys.argv[1] if len(sys.argv) > 1 else "c2"
hint = int(sys.argv[2]) if len(sys.argv) > 2 else 0
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 128
if case == "c2":
    r = run(0, T.ar1_gaussian(0.9), 128, 8192, steps, hint, 0.1, reps=1)
elif case == "c3":
    r = run(1, T.rosenbrock(), 64, 32768, steps, hint, 0.01, 10, reps=1)
elif case == "c5":
    r = run(0, T.gaussian_mixture(np.random.RandomState(0).randn(8, 64) * 3), 64, 65536, steps, hint, 0.1, reps=1)
print(case, hint, steps, r)
