"""Where the e2e leg of bench.py differs from the device-resident leg: wall clock of the Python call, device span reported by
the library for the same call, host vs device inputs.  python tools/e2e_probe.py [steps]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import load_pairs  # noqa: E402

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
pairs = load_pairs(16)
import torch  # noqa: E402
import aicp_mapping_b200 as ab  # noqa: E402
from aicp_mapping_b200 import capi  # noqa: E402

ovl = ab.B200Overlap()
reg = ab.B200Registration()
reg.setProfiling(1)
dev, host, ratios = [], [], []
for p in pairs:
    ovl.computeOverlap(p["ref"], p["read"], p["ref_origin"], p["read_origin"])
    ratios.append(ab.autotune_ratio(float(ovl.getOverlap())))
    r4, q4 = capi.to_xyzw(p["ref"]), capi.to_xyzw(p["read"])
    dev.append((torch.from_numpy(r4).cuda(), torch.from_numpy(q4).cuda()))
    hr, hq = torch.from_numpy(r4).pin_memory(), torch.from_numpy(q4).pin_memory()
    host.append((hr, hq, hr.numpy(), hq.numpy()))
P = 64
order = [j % 16 for j in range(P)]
br = [ratios[k] for k in order]
dev_batch = [(dev[k][0], dev[k][1]) for k in order]
host_batch = [(host[k][2], host[k][3]) for k in order]
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for name, batch in (("device inputs", dev_batch), ("pinned host inputs", host_batch), ("device inputs", dev_batch), ("pinned host inputs", host_batch)):
    walls, spans = [], []
    for s in range(steps + 2):
        flush.zero_(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        T, stats, status, ms = reg.registerBatch(batch, ratios=br, streams=8)
        w = (time.perf_counter() - t0) * 1e3
        if s >= 2:
            walls.append(w); spans.append(ms)
    print("%-20s wall %.2f ms  device span %.2f ms  (python + call overhead %.2f ms)  -> %.1f / %.1f registrations/s" % (
        name, np.mean(walls), np.mean(spans), np.mean(walls) - np.mean(spans), P / np.mean(walls) * 1e3, P / np.mean(spans) * 1e3))

# device inputs while an unrelated stream uploads the same 268 MB per step from pinned memory: is the longer device span of
# the host-input leg exposed copy time (no: upload-ahead changes nothing) or interference of the PCIe traffic itself?
import threading  # noqa: E402
big_h = torch.empty(64 << 20, dtype=torch.uint8).pin_memory()
big_d = torch.empty(64 << 20, dtype=torch.uint8, device="cuda")
cs = torch.cuda.Stream()
stop = False


def pump():
    while not stop:
        with torch.cuda.stream(cs):
            for k in range(4):
                big_d.copy_(big_h, non_blocking=True)
        cs.synchronize()


th = threading.Thread(target=pump)
th.start()
walls, spans = [], []
for s in range(steps + 2):
    flush.zero_(); torch.cuda.current_stream().synchronize()
    t0 = time.perf_counter()
    T, stats, status, ms = reg.registerBatch(dev_batch, ratios=br, streams=8)
    w = (time.perf_counter() - t0) * 1e3
    if s >= 2:
        walls.append(w); spans.append(ms)
stop = True
th.join()
print("%-20s wall %.2f ms  device span %.2f ms  -> %.1f registrations/s" % ("device inputs + background H2D", np.mean(walls), np.mean(spans), P / np.mean(spans) * 1e3))
