"""cProfile of the host side of the training step (which Python/torch calls eat the launch budget)."""
import argparse, cProfile, os, pstats, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import torch
ap = argparse.ArgumentParser(); ap.add_argument("--mode", default="bf16"); a = ap.parse_args()
step, host, resident, _ = bench.setup_training(dict(bench.CFG), a.mode, 8)
for i in range(3):
    step(*resident[i])
torch.cuda.synchronize()
t0 = time.time()
for i in range(3, 6):
    step(*resident[i])
t1 = time.time()
torch.cuda.synchronize()
t2 = time.time()
print(f"host issue {1e3*(t1-t0)/3:.2f} ms/step, wall {1e3*(t2-t0)/3:.2f} ms/step")
pr = cProfile.Profile()
pr.enable()
for i in range(5, 8):
    step(*resident[i])
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(28)
