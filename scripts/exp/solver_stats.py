"""Experiment: Newton iteration statistics of the engine core (CPU build) on the bench workload."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", ".."))
from tests.hostcore import hostcore as hc
from oracle import oracle as orc
import ctypes as C
if os.path.exists(os.path.join(os.path.dirname(__file__), 'libhc_stats.so')):
    hc._lib = C.CDLL(os.path.join(os.path.dirname(__file__), 'libhc_stats.so')); hc._lib.hc_ls_evals.restype = C.c_long

def tilt(q):
    w, x, y, z = q[3:7]
    return np.degrees(np.arccos(np.clip(1 - 2 * (x * x + y * y), -1, 1)))

def run(terrain, episodes, seed=0):
    rng = np.random.default_rng(seed)
    nit, ncon, lens = [], [], []
    for ep in range(episodes):
        if terrain == "perlin":
            hf = orc.perlin_terrain(seed=int(rng.integers(0, 10000)))
        else:
            hf = np.zeros(293 * 293, np.float32)
        off = orc.lib().bbo_spawn_offset(orc._fp(hf), 2.0)
        q = np.array([0, 0, 0.24 + off, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0.26 + off, 1, 0, 0, 0], float)
        v = np.zeros(15); w = np.zeros(15)
        for t in range(4000):
            a = rng.uniform(-1, 1, 3)
            q, v, w, kin, nc, ni = hc.step(q, v, w, -10 * a, hf)
            nit.append(ni); ncon.append(nc)
            if tilt(q) > 20: break
        lens.append(t + 1)
    nit = np.array(nit); ncon = np.array(ncon)
    if hasattr(hc._lib, 'hc_ls_evals'): print('  ls evals total', hc._lib.hc_ls_evals(), 'per newton iter %.2f' % (hc._lib.hc_ls_evals() / max(nit.sum(), 1)))
    print(terrain, "episodes", episodes, "mean len", np.mean(lens))
    print("  niter/step mean %.1f  median %d  p90 %d  p99 %d  max %d" % (nit.mean(), np.median(nit), np.percentile(nit, 90), np.percentile(nit, 99), nit.max()))
    print("  ncon(max over stages) mean %.2f hist" % ncon.mean(), np.bincount(ncon)[:16])
    for k in range(0, 12):
        m = ncon == k
        if m.sum(): print("   ncon=%d: %5d steps, niter mean %.1f" % (k, m.sum(), nit[m].mean()))

if __name__ == "__main__":
    run(sys.argv[1] if len(sys.argv) > 1 else "perlin", int(sys.argv[2]) if len(sys.argv) > 2 else 10)
