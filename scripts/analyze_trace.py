"""Summarise a torch.profiler chrome trace: device busy time (union over streams), idle gaps and what the host was
doing during them (development aid for the multi-rank step)."""
import gzip, json, sys

path = sys.argv[1]
raw = (gzip.open(path, "rt") if path.endswith(".gz") else open(path)).read()
ev = json.loads(raw)["traceEvents"]
k = [e for e in ev if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "dur" in e]
cpu = [e for e in ev if e.get("cat") in ("cpu_op", "user_annotation", "cuda_runtime", "cuda_driver") and "dur" in e]
k.sort(key=lambda e: e["ts"])
import os
if os.environ.get("TRACE_FROM"):          # keep only the tail of the trace (fraction of the span), e.g. the steady-state steps
    a0, a1 = k[0]["ts"], max(e["ts"] + e["dur"] for e in k)
    cut = a0 + float(os.environ["TRACE_FROM"]) * (a1 - a0)
    k = [e for e in k if e["ts"] >= cut]
t0, t1 = k[0]["ts"], max(e["ts"] + e["dur"] for e in k)
print(f"device span {t1 - t0:.0f} us, {len(k)} device activities, streams: {sorted({e['args'].get('stream') for e in k})}")
busy, cur_s, cur_e, gaps = 0.0, None, None, []
for e in k:
    s, e_ = e["ts"], e["ts"] + e["dur"]
    if cur_e is None:
        cur_s, cur_e = s, e_
    elif s <= cur_e:
        cur_e = max(cur_e, e_)
    else:
        busy += cur_e - cur_s
        gaps.append((s - cur_e, cur_e, s, e["name"][:60]))
        cur_s, cur_e = s, e_
busy += cur_e - cur_s
print(f"device busy (union) {busy:.0f} us, idle {t1 - t0 - busy:.0f} us")
per_stream = {}
for e in k:
    per_stream.setdefault(e["args"].get("stream"), 0.0)
    per_stream[e["args"].get("stream")] += e["dur"]
print("per-stream kernel time:", {s: round(v) for s, v in per_stream.items()})
gaps.sort(reverse=True)
print("largest idle gaps (us, then the kernel that ended the gap, and the longest host ops overlapping the gap):")
for g, a, b, name in gaps[:25]:
    ops = [(min(c["ts"] + c["dur"], b) - max(c["ts"], a), c["name"][:50]) for c in cpu if c["ts"] < b and c["ts"] + c["dur"] > a]
    ops.sort(reverse=True)
    print(f"  {g:7.1f} at +{a - t0:8.0f}  -> {name:60s} | host: " + "; ".join(f"{n} {d:.0f}" for d, n in ops[:4]))
# main-stream order of kernels with gaps before each (first 200)
if len(sys.argv) > 2:
    prev = None
    for e in k:
        gap = 0 if prev is None else e["ts"] - prev
        print(f"{e['ts'] - t0:9.1f} {e['dur']:8.1f} s{e['args'].get('stream')} gap {gap:7.1f} {e['name'][:90]}")
        prev = max(prev or 0, e["ts"] + e["dur"])
