"""C4 smoke (BASELINE.json configs[3]): device-resident PPO iteration(s) with the env shards on N GPUs and the gradient
all-reduce over NCCL.  Launch: torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/ppo_smoke.py [--envs E]"""
import argparse, os, sys, time
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from openballbot_rl_b200.training.policy import BallbotPolicy
from openballbot_rl_b200.training.ppo import PPOConfig, PPOLearner
from openballbot_rl_b200.training.utils import make_ballbot_vec_env

ap = argparse.ArgumentParser()
ap.add_argument("--envs", type=int, default=4096, help="total envs over all ranks")
ap.add_argument("--n-steps", type=int, default=64)
ap.add_argument("--iters", type=int, default=2)
a = ap.parse_args()
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
torch.manual_seed(rank)
venv = make_ballbot_vec_env(a.envs, terrain_config={"type": "perlin", "config": {}}, seed=0, device=local, rank=rank, world_size=world)
pol = BallbotPolicy().to(f"cuda:{local}")
L = PPOLearner(venv, pol, PPOConfig(n_steps=a.n_steps, batch_size=8192, n_epochs=2), total_timesteps=a.envs * a.n_steps * a.iters)
t0 = time.perf_counter()
L.learn(callback=lambda d: rank == 0 and print({k: (round(v, 4) if isinstance(v, float) else v) for k, v in d.items()}, flush=True))
torch.cuda.synchronize()
dt = time.perf_counter() - t0
flat = torch.cat([p.detach().reshape(-1) for p in L.params])
if world > 1:
    ref = flat.clone(); dist.broadcast(ref, src=0)
    same = torch.tensor([float(torch.equal(ref, flat))], device=flat.device); dist.all_reduce(same, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("replicas identical on all ranks:", bool(same.item()))
if rank == 0:
    print(f"{L.num_timesteps} env-steps (rollout + update) in {dt:.1f} s on {world} GPU(s): {L.num_timesteps / dt / 1e6:.2f} M env-steps/s incl. learner")
venv.close()
if world > 1:
    dist.destroy_process_group()
