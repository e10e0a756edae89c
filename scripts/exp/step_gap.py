import sys, time, torch
sys.path.insert(0, ".")
from openballbot_rl_b200.engine import BallbotEngine
N = 65536
eng = BallbotEngine(num_envs=N, precision=64, terrain="perlin", cameras=True, seed=0)
eng.reset()
g = torch.Generator(device="cuda"); g.manual_seed(0)
act = torch.rand(16, N, 3, device="cuda", generator=g) * 2 - 1
for t in range(300): eng.step(act[t % 16])
def run(label, prof=False, summ=False, K=40):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    dc = torch.zeros((), dtype=torch.int64, device="cuda")
    if prof: eng.profile_begin(K)
    torch.cuda.synchronize(); e0.record()
    for t in range(K):
        eng.step(act[t % 16])
        if summ: dc += eng.terminated.sum()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / K
    extra = ""
    if prof:
        p = eng.profile_end(); extra = " kernels: step %.2f terrain %.2f reset %.2f depth %.2f sum %.2f" % tuple([p[k] / K for k in ("step_ms", "terrain_ms", "reset_ms", "depth_ms")] + [sum(p[k] for k in ("step_ms", "terrain_ms", "reset_ms", "depth_ms")) / K])
    print(f"{label}: {ms:.2f} ms/step{extra}")
run("plain"); run("plain"); run("profile events", prof=True); run("with sum", summ=True); run("plain")
