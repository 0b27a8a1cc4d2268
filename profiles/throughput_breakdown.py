#!/usr/bin/env python
"""Where does the GPU time of an alignment go when many are in flight?  Batched throughput
(sm_align_pairs, 16 pipelines, 2 host threads, device-resident clouds) at several iteration counts:
the slope is the cost of one ICP iteration under concurrency, the intercept the prologue
(centre + k-d tree + Morton sort of the source)."""
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")
import torch  # noqa: E402
import bench  # noqa: E402
import staticmapping_b200 as smb  # noqa: E402

dev = torch.device("cuda", 0)
P, T, B, STEPS = 16, 2, 64, 6
data = [bench.PairData(smb, torch, dev, 0, w) for w in range(P)]
out = {}
for iters in (1, 8, 15, 30):
    for qpc in (0, 1024):
        ms_ = []
        for _ in range(P):
            m = smb.IcpFast(0)
            m.InitWithXml({"max_iteration": iters, "disable_convergence_check": 1, "knn_queries_per_cta": qpc})
            ms_.append(m)
        pairs = [data[k % P].pair(False) for k in range(B)]

        def body(j, steps):
            for _ in range(steps):
                smb.AlignPairs(ms_[j::T], pairs[j::T])

        def run(steps):
            ths = [threading.Thread(target=body, args=(j, steps)) for j in range(T)]
            torch.cuda.synchronize(); t0 = time.perf_counter()
            for t in ths: t.start()
            for t in ths: t.join()
            torch.cuda.synchronize()
            return time.perf_counter() - t0
        run(2)
        secs = run(STEPS)
        out[f"iters{iters}_qpc{qpc}"] = {"alignments_per_s": STEPS * B / secs, "ms_per_alignment": 1e3 * secs / (STEPS * B)}
        print(iters, qpc, out[f"iters{iters}_qpc{qpc}"], flush=True)
        for m in ms_:
            m.__del__()
print(json.dumps(out))
