"""Host-side cProfile of the GD step at a tiny batch (GPU time negligible -> the profile is the enqueue cost)."""
import cProfile, pstats, sys, os, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from efficientvlm_b200 import ops
from efficientvlm_b200.distill import XVLM, gd_loss
from efficientvlm_b200.optim import LinearWarmupDecay, create_optimizer

dev = torch.device("cuda", 0)
torch.manual_seed(42)
student = XVLM(bench.make_cfg("student", 224)).to(dev).train()
teacher = XVLM(bench.make_cfg("teacher", 224)).to(dev).eval()
for p in teacher.parameters():
    p.requires_grad_(False)
opt = create_optimizer(dict(lr=1e-4, weight_decay=0.01, lr_mult=2), student, clip_grad_norm=1.0)
sched = LinearWarmupDecay(opt, 100000, 2)
ops.manual_seed(1)
batch = [t.to(dev) for t in bench.make_batch(8, 224, 1)]

def step():
    so = student(*batch, output_attentions=True, output_hidden_states=True)
    with torch.no_grad():
        to = teacher(*batch, output_attentions=True, output_hidden_states=True)
    total, _ = gd_loss(so, to, 1.0)
    total.backward()
    opt.step(); sched.step(); opt.zero_grad()

import time, gc
def timed_sections(tag, n=5):
    acc = [0.0] * 6
    for _ in range(n):
        torch.cuda.synchronize()
        t = [time.perf_counter()]
        so = student(*batch, output_attentions=True, output_hidden_states=True); t.append(time.perf_counter())
        with torch.no_grad():
            to = teacher(*batch, output_attentions=True, output_hidden_states=True)
        t.append(time.perf_counter())
        total, _ = gd_loss(so, to, 1.0); t.append(time.perf_counter())
        total.backward(); t.append(time.perf_counter())
        opt.step(); sched.step(); opt.zero_grad(); t.append(time.perf_counter())
        del so, to, total
        torch.cuda.synchronize(); t.append(time.perf_counter())
        for i in range(6):
            acc[i] += (t[i + 1] - t[i]) * 1e3 / n
    print(tag, "student fwd %.2f | teacher fwd %.2f | loss %.2f | backward %.2f | opt %.2f | drain %.2f | total %.2f ms" % (*acc, sum(acc)))
for _ in range(3):
    step()
timed_sections("gc on ")
gc.disable()
timed_sections("gc off")
gc.enable()
gc.freeze()
timed_sections("gc frozen")
if len(sys.argv) > 1:
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(3):
        step()
    pr.disable()
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(45)
    print(s.getvalue()[:9000])
