"""GPU timeline of the bench's training step (torch.profiler / CUPTI; nsys is not in the image).

    python scripts/timeline.py [--config wanjinyou] [--steps 6] [--no-pipeline-march] > gpurun_out/timeline.json

Per stream: busy time, kernel list with start offsets inside one steady-state step; overall: the step's critical path
(union of busy intervals over all streams), idle gaps, and for every kernel whether it overlaps kernels on ANOTHER stream
(that is how "the march runs behind the backward" and "the scatter co-runs with the dense chain" are shown).
Times under the profiler are inflated by its own overhead; the bench line is the number, this is the picture.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    from torch.profiler import ProfilerActivity, profile
    import bench
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="wanjinyou")
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--rays", type=int, default=0)
    ap.add_argument("--no-pipeline-march", dest="pipeline_march", action="store_false")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    prob = bench.build_problem(0, 1, args, dev)
    o, d, cam, gt = (torch.from_numpy(x).to(dev) for x in prob["host"])
    nxt = (o, d) if args.pipeline_march else None
    for _ in range(5):
        bench.train_step(prob, o, d, cam, gt, None, nxt)
    torch.cuda.synchronize()
    marks = []
    with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
        for i in range(args.steps):
            torch.cuda.nvtx.range_push(f"step{i}")
            bench.train_step(prob, o, d, cam, gt, None, nxt)
            torch.cuda.nvtx.range_pop()
        torch.cuda.synchronize()
    path = "/tmp/f2b_trace.json"
    prof.export_chrome_trace(path)
    ev = json.load(open(path))["traceEvents"]
    kern = [e for e in ev if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "dur" in e]
    kern.sort(key=lambda e: e["ts"])
    # steady-state window: from the start of the march (or field pass) of step 2 to the same point of step 3
    anchors = [e["ts"] for e in kern if "field_fwd_kernel" in e["name"]]
    if len(anchors) < 4:
        print(json.dumps({"error": "too few steps captured", "n_kernels": len(kern)}))
        return
    t0, t1 = anchors[2], anchors[3]
    win = [e for e in kern if t0 <= e["ts"] < t1]
    streams = {}
    for e in win:
        streams.setdefault(e["args"].get("stream", -1), []).append(e)

    def union(iv):
        iv = sorted(iv)
        tot, cur_s, cur_e = 0.0, None, None
        for s, e_ in iv:
            if cur_e is None or s > cur_e:
                if cur_e is not None:
                    tot += cur_e - cur_s
                cur_s, cur_e = s, e_
            else:
                cur_e = max(cur_e, e_)
        return tot + ((cur_e - cur_s) if cur_e is not None else 0.0)

    out = {"config": args.config, "pipeline_march": args.pipeline_march, "step_us_under_profiler": t1 - t0,
           "busy_union_us": union([(e["ts"], e["ts"] + e["dur"]) for e in win]), "streams": {}}
    out["idle_us"] = out["step_us_under_profiler"] - out["busy_union_us"]
    short = lambda n: n.split("(")[0].replace("void ", "").replace("f2b::", "")[:70]
    for sid, es in streams.items():
        others = [(x["ts"], x["ts"] + x["dur"]) for s2, e2 in streams.items() if s2 != sid for x in e2]
        rows = []
        for e in es:
            a, b = e["ts"], e["ts"] + e["dur"]
            ov = sum(max(0.0, min(b, y) - max(a, x)) for x, y in others)
            rows.append({"name": short(e["name"]), "start_us": round(a - t0, 1), "dur_us": round(e["dur"], 1),
                         "overlap_with_other_streams_us": round(min(ov, e["dur"]), 1)})
        agg = {}
        for r in rows:
            a = agg.setdefault(r["name"], [0, 0.0, 0.0])
            a[0] += 1; a[1] += r["dur_us"]; a[2] += r["overlap_with_other_streams_us"]
        out["streams"][str(sid)] = {"busy_us": round(union([(e["ts"], e["ts"] + e["dur"]) for e in es]), 1), "n_kernels": len(es),
                                    "by_kernel": sorted(({"name": k, "calls": v[0], "dur_us": round(v[1], 1), "overlapped_us": round(v[2], 1)}
                                                         for k, v in agg.items()), key=lambda r: -r["dur_us"])[:40],
                                    "sequence": rows if len(rows) <= 12 else None}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
