"""Upload / download throughput between PAGEABLE host memory and the device (the normalizer path's staging pipeline)."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from gcmapexplorer_b200 import device as gx
n = 24926
A = np.zeros((n, n), dtype=np.float32); A[:] = 1.0
out = np.empty_like(A); out[:] = 0
with gx.DeviceMap(n) as dm:
    dm.upload(A); dm.download(0, out=out)
    t0 = time.perf_counter(); dm.upload(A); tu = time.perf_counter() - t0
    t0 = time.perf_counter(); dm.download(0, out=out); td = time.perf_counter() - t0
print(json.dumps({"threads": os.environ.get("GCX_COPY_THREADS", "8"), "upload_GBps": round(A.nbytes / tu / 1e9, 1), "download_GBps": round(A.nbytes / td / 1e9, 1)}))
