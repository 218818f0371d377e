"""Probe: what does a tiny kernel cost between two CUDA events when the stream is kept busy? (context for nq_scan's fixed cost)"""
import torch
x = torch.zeros(1024, device="cuda")
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    for _ in range(20):
        x.add_(1)
    torch.cuda.synchronize()
    res = []
    for rep in range(5):
        evs = []
        for i in range(8):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(s); x.add_(1); b.record(s)
            evs.append((a, b))
        torch.cuda.synchronize()
        res += [a.elapsed_time(b) * 1e3 for a, b in evs]
    res.sort()
    print("tiny torch kernel between events, queue busy: median %.1f us  min %.1f us" % (res[len(res) // 2], res[0]))
