"""Idle gaps between the kernels of one device-resident step (torch profiler timeline)."""
import sys, os, json, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from torch.profiler import profile, ProfilerActivity
exec(open(os.path.join(os.path.dirname(__file__), "graph_probe.py")).read().split("def timed")[0])
for _ in range(5): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(4): step()
    torch.cuda.synchronize()
prof.export_chrome_trace("/tmp/trace.json")
ev = [e for e in json.load(open("/tmp/trace.json"))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")]
ev.sort(key=lambda e: e["ts"])
n = len(ev) // 4
one = ev[2 * n:3 * n]
t0 = one[0]["ts"]; busy = 0.0; prev_end = None
for e in one:
    gap = 0.0 if prev_end is None else e["ts"] - prev_end
    busy += e["dur"]; prev_end = e["ts"] + e["dur"]
    print(f"{e['ts']-t0:9.1f} us  dur {e['dur']:8.1f}  gap {gap:6.1f}  {e['name'][:60]}")
print(f"step span {prev_end - t0:.1f} us, busy {busy:.1f} us, gaps {prev_end - t0 - busy:.1f} us over {len(one)} kernels")
