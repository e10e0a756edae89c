"""Depth-encoder forward cost against batch-shape churn: a new batch size on every call (what the embedding cache produces)
against sizes padded to a bucket, eval mode, fp32."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from openballbot_rl_b200.training.policy import make_depth_encoder
dev = torch.device("cuda", 0)
enc = make_depth_encoder().to(dev).eval()
x = torch.rand(8192, 1, 64, 64, device=dev)


def run(sizes, label, **kw):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    with torch.no_grad():
        for b in sizes:
            enc(x[:b].contiguous(**kw) if kw else x[:b])
    torch.cuda.synchronize()
    print(f"{label:50s} {(time.perf_counter() - t0) / len(sizes) * 1e3:7.3f} ms / call")


run([2731] * 20, "warm-up")
run([2731] * 50, "fixed batch 2731")
run(list(range(2600, 2900, 6)), "new batch size every call (first visit)")
run(list(range(2600, 2900, 6)), "same sizes, second visit")
run([3072] * 50, "fixed batch 3072 (bucket of 512)")
enc_cl = enc.to(memory_format=torch.channels_last)
run([3072] * 50, "fixed 3072, channels_last weights")
torch.backends.cudnn.benchmark = True
run([3072] * 50, "fixed 3072, cudnn.benchmark")
run([3072] * 50, "fixed 3072, cudnn.benchmark (2)")
with torch.autocast("cuda", dtype=torch.bfloat16):
    run([3072] * 50, "fixed 3072, bf16 autocast")
