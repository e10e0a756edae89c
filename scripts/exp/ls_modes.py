"""Experiment (CPU build of the engine core): solver_mode 0 (the reference's exact line search, every RK stage warm-started
from qacc_warmstart) against solver_mode 1 (strong-Wolfe line search with cone-apex candidates, warm start chained through
the stages) -- Newton iterations, line-search evaluations and the distance between the two solutions after one step, on
every state of random-action rollouts of the bench workload."""
import sys, os, subprocess, ctypes as C
import numpy as np
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
sys.path.insert(0, ROOT)
SO = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libhc_stats.so")
subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-DBB_STATS", "-x", "c++", "-o", SO, os.path.join(ROOT, "tests/hostcore/hostcore.cpp")])
from tests.hostcore import hostcore as hc
from oracle import oracle as orc
hc._lib = C.CDLL(SO); L = hc._lib
L.hc_ls_evals.restype = C.c_long; L.hc_newton_iters.restype = C.c_long


def tilt(q):
    w, x, y, z = q[3:7]
    return np.degrees(np.arccos(np.clip(1 - 2 * (x * x + y * y), -1, 1)))


def states(terrain, episodes, seed=0):
    """(qpos, qvel, warm, ctrl, hf) of every step of reference-path rollouts under random actions"""
    rng = np.random.default_rng(seed); out = []
    L.hc_set_solver(0)
    for ep in range(episodes):
        hf = orc.perlin_terrain(seed=int(rng.integers(0, 10000))) if terrain == "perlin" else np.zeros(293 * 293, np.float32)
        off = orc.lib().bbo_spawn_offset(orc._fp(hf), 2.0)
        q = np.array([0, 0, 0.24 + off, 1, 0, 0, 0, 0, 0, 0, 0, 0, 0.26 + off, 1, 0, 0, 0], float)
        v = np.zeros(15); w = np.zeros(15)
        for t in range(1500):
            a = rng.uniform(-1, 1, 3)
            out.append((q.copy(), v.copy(), w.copy(), -10 * a, hf))
            q, v, w, kin, nc, ni = hc.step(q, v, w, -10 * a, hf)
            if tilt(q) > 20: break
    return out


def run(st, mode):
    L.hc_set_solver(mode); L.hc_stats_reset()
    res = []
    for (q, v, w, c, hf) in st:
        q2, v2, w2, kin, nc, ni = hc.step(q, v, w, c, hf)
        res.append(np.concatenate([q2, v2]))
    return np.array(res), L.hc_newton_iters(), L.hc_ls_evals()


if __name__ == "__main__":
    terrain = sys.argv[1] if len(sys.argv) > 1 else "perlin"
    st = states(terrain, int(sys.argv[2]) if len(sys.argv) > 2 else 8)
    ref, n0, e0 = run(st, 0)
    r, n, e = run(st, 1)
    d = np.abs(r - ref); rel = d[:, 17:] / (np.abs(ref[:, 17:]).max(axis=1, keepdims=True) + 1e-12)
    print(terrain, len(st), "steps")
    print("  solver_mode 0: newton iterations / step %.2f, line-search evaluations / step %.1f (%.2f per iteration)" % (n0 / len(st), e0 / len(st), e0 / max(n0, 1)))
    print("  solver_mode 1: newton iterations / step %.2f, line-search evaluations / step %.1f (%.2f per iteration)" % (n / len(st), e / len(st), e / max(n, 1)))
    print("  one-step difference: max |dqpos| %.1e, max |dqvel| / max |qvel| %.1e" % (d[:, :17].max(), rel.max()))
