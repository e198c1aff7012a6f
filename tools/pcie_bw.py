#!/usr/bin/env python
"""PCIe floor of the end-to-end leg: pinned H2D, D2H and both at once for the 2.1 GB vectors of the headline workload."""
import json, torch
n = 262144000
h_in = torch.empty(n, dtype=torch.float64, pin_memory=True); h_in.zero_()
h_out = torch.empty(n, dtype=torch.float64, pin_memory=True); h_out.zero_()
d_in = torch.empty(n, dtype=torch.float64, device="cuda"); d_out = torch.zeros(n, dtype=torch.float64, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def timed(f, reps=3):
    f(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f()
    torch.cuda.synchronize(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
def h2d():
    with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
    torch.cuda.current_stream().wait_stream(s1)
def d2h():
    with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
    torch.cuda.current_stream().wait_stream(s2)
def both():
    with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
    with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
    torch.cuda.current_stream().wait_stream(s1); torch.cuda.current_stream().wait_stream(s2)
gb = n * 8 / 1e9
a, b, c = timed(h2d), timed(d2h), timed(both)
print(json.dumps({"bytes_each_way": n * 8, "h2d_ms": a, "h2d_GBs": gb / a * 1e3, "d2h_ms": b, "d2h_GBs": gb / b * 1e3, "both_ms": c, "both_GBs_each": gb / c * 1e3}))
