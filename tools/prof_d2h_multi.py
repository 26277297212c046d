"""Plain pinned device-to-host copy rate of this box: one device alone, then all visible devices at
once (one thread and one stream per device, each into its own part of ONE pinned host buffer).  The
yardstick for the host-output all-to-all call, whose row strips travel exactly like this."""
import ctypes as C
import json
import os
import sys
import threading
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from loos_b200 import capi  # noqa: E402

GB = float(os.environ.get("D2H_GB", "4"))
n = int(GB * (1 << 30))
nd = torch.cuda.device_count()
L = capi.lib()
host = L.loosb200_host_alloc(n * nd)
assert host
rt = C.CDLL("libcudart.so.12")
srcs = []
for d in range(nd):
    torch.cuda.set_device(d)
    srcs.append(torch.empty(n, dtype=torch.uint8, device="cuda:%d" % d))
    torch.cuda.synchronize()


def copy(d, reps, out):
    torch.cuda.set_device(d)
    s = torch.cuda.Stream(device=d)
    dst = host + d * n
    with torch.cuda.stream(s):
        rt.cudaMemcpyAsync(C.c_void_p(dst), C.c_void_p(srcs[d].data_ptr()), C.c_size_t(n), 2, C.c_void_p(s.cuda_stream))
        s.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            rt.cudaMemcpyAsync(C.c_void_p(dst), C.c_void_p(srcs[d].data_ptr()), C.c_size_t(n), 2, C.c_void_p(s.cuda_stream))
        s.synchronize()
    out[d] = reps * n / (time.perf_counter() - t0) / 1e9


for group in [[0]] + ([list(range(2))] if nd >= 2 else []) + ([list(range(nd))] if nd > 2 else []):
    out = {}
    th = [threading.Thread(target=copy, args=(d, 3, out)) for d in group]
    t0 = time.perf_counter()
    [t.start() for t in th]
    [t.join() for t in th]
    print(json.dumps({"measure": "pinned_d2h", "devices": group, "per_device_GBps": [round(out[d], 1) for d in group],
                      "aggregate_GBps": round(sum(out.values()), 1)}))
L.loosb200_host_free(host)
