import numpy as np, torch, sys
sys.path.insert(0, ".")
from openballbot_rl_b200.engine import BallbotEngine
from tests.hostcore import hostcore as H
np.set_printoptions(precision=9, suppress=True, linewidth=200)
eng = BallbotEngine(num_envs=2, precision=64, terrain="flat", cameras=False, auto_reset=False)
eng.reset()
q0, v0, w0 = [x.cpu().numpy() for x in eng.get_state()]
dev = eng.device
c = torch.zeros(3, dtype=torch.float64, device=dev); out = torch.zeros(64, dtype=torch.float64, device=dev)
cd = torch.zeros(53, dtype=torch.float64, device=dev); cp = torch.zeros(53, 3, dtype=torch.float64, device=dev); cf = torch.zeros(53, 9, dtype=torch.float64, device=dev)
import ctypes as C
eng._L.bb_probe_forward(eng._h, 0, C.c_void_p(c.data_ptr()), C.c_void_p(out.data_ptr()), C.c_void_p(cd.data_ptr()), C.c_void_p(cp.data_ptr()), C.c_void_p(cf.data_ptr()), None)
torch.cuda.synchronize()
o = out.cpu().numpy()
s = H.step(q0[0], v0[0], w0[0], np.zeros(3), None, prec=64)
o = out.cpu().numpy()
print("probe rk4: vz", o[54], "ballvz", o[55], "warm z", o[56], "qz", o[57])
print("host  rk4: vz", s[1][2], "ballvz", s[1][11], "warm z", s[2][2], "qz", s[0][2])
eng.step(torch.zeros(2, 3, device=dev))
q1, v1, w1 = [x.cpu().numpy() for x in eng.get_state()]
print("k_step   : vz", v1[0][2], "ballvz", v1[0][11], "warm z", w1[0][2], "qz", q1[0][2])
print("device per-stage acc z:", o[58:62], "stage4 xv z", o[62], "xq z", o[63])
s_ = (q0[0].copy(), v0[0].copy(), w0[0].copy())
# host per-stage via forward
xq, xv = q0[0].copy(), v0[0].copy(); X = None
for sg in range(4):
    f = H.forward(xq, xv, np.zeros(3), w0[0], prec=64)
    print(" host stage", sg, "acc z", f["qacc"][2], "ncon", f["ncon"], "niter", f["niter"])
    if sg == 0: X = (xq.copy(), xv.copy())
    if sg < 3:
        a = 1.0 if sg == 2 else 0.5
        xq = X[0].copy(); xq[:3] += 0.002 * a * xv[:3]; xq[10:13] += 0.002 * a * xv[9:12]   # (rotation ignored: zero)
        xv = X[1] + 0.002 * a * f["qacc"]
