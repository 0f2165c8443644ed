"""One eager training cycle between cudaProfilerStart/Stop (for `ncu --profile-from-start off`)."""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "melo-gan_b200")]
import torch
import bench
from melogan.trainer import GanTrainer

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=1024)
ap.add_argument("--precision", default="bf16")
ap.add_argument("--cycles", type=int, default=1)
a = ap.parse_args()
cfg, ed_cfg = bench.load_cfgs()
dev = torch.device("cuda:0")
tr = GanTrainer(cfg, ed_cfg, batch=a.batch, precision=a.precision, device=dev)
g = torch.Generator(device=dev).manual_seed(7)
reals = torch.rand((5, a.batch, 512, 4), generator=g, device=dev) * 2 - 1
nums = torch.randn((5, a.batch, 6), generator=g, device=dev)
labels = (torch.arange(a.batch, device=dev) % 4).to(torch.int64)
for _ in range(2):
    tr.train_cycle(reals, nums, labels)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
for _ in range(a.cycles):
    tr.train_cycle(reals, nums, labels)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("done", tr.epoch_means())
