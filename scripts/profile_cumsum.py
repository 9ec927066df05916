import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import modppl_b200 as m
rng = np.random.default_rng(0)
w = np.exp(rng.normal(size=1 << 24)); p = w / w.sum()
s = m.parity.cumsum_sequential(p)
print("ok", s[-1])
