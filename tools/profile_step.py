"""One training step of the bench workload inside a cudaProfilerStart/Stop window (for ncu --profile-from-start off).

    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
        --log-file gpurun_out/launches.csv python tools/profile_step.py [--mode tf32] [--steps 1]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--mode", default="tf32")
ap.add_argument("--steps", type=int, default=1)
ap.add_argument("--warmup", type=int, default=3)
a = ap.parse_args()
import torch  # noqa: E402

step, host, resident, _ = bench.setup_training(dict(bench.CFG), a.mode, a.warmup + a.steps)
for i in range(a.warmup):
    step(*resident[i])
torch.cuda.synchronize()
torch.cuda.profiler.start()
for i in range(a.steps):
    step(*resident[a.warmup + i])
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("profiled", a.steps, "step(s)")
