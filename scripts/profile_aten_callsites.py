"""Attribute the framework-launched (at::) kernels of one eager step to Python call sites: torch.profiler with stacks.
usage: python scripts/profile_aten_callsites.py [workload] (default itr_step)"""
import argparse
import collections
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402


def main():
    wlname = sys.argv[1] if len(sys.argv) > 1 else "itr_step"
    sys.argv = [sys.argv[0], "--workload", wlname]
    ap_defaults = dict(itr_step=(128, 384), gd=(128, 224), vqa_step=(16, 480))
    args = argparse.Namespace(workload=wlname, batch=ap_defaults[wlname][0], image_res=ap_defaults[wlname][1], sparsity=0.25, loga_shift=0.0,
                              eager=True, gate_loga=None, no_zero_skip=False, no_overlap=True, materialize=False, beams=1)
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    build = {"gd": bench.build_gd, "itr_step": bench.build_itr_step, "vqa_step": bench.build_vqa_step}[wlname]
    wl = build(args, dev, 0, 1)
    batch = [t.to(dev) for t in wl["host"]]
    for _ in range(3):
        wl["device_step"](*batch)
    torch.cuda.synchronize()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=True, record_shapes=True) as prof:
        wl["device_step"](*batch)
        torch.cuda.synchronize()
    rows = collections.defaultdict(lambda: [0, 0.0])
    for ev in prof.key_averages(group_by_input_shape=True, group_by_stack_n=12):
        t = getattr(ev, "self_device_time_total", 0) or getattr(ev, "self_cuda_time_total", 0)
        if t <= 0 or not ev.key.startswith("aten::"):
            continue
        stack = [f for f in (ev.stack or []) if "efficientvlm_b200" in f or "bench.py" in f][:3]
        key = (ev.key, str(ev.input_shapes)[:90], " <- ".join(s.split("/")[-1] for s in stack))
        rows[key][0] += ev.count
        rows[key][1] += t
    tot = sum(v[1] for v in rows.values())
    print("aten ops with device time: %.1f us total" % tot)
    for k, v in sorted(rows.items(), key=lambda kv: -kv[1][1])[:28]:
        print("%8.1f us n=%3d %-22s %s\n             %s" % (v[1], v[0], k[0], k[1], k[2]))

main()
