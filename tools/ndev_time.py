import sys, time, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np, mdz_b200
from views import config4m
v = config4m(3840, 2160, 100000)
n = mdz_b200.device_count()
devs = tuple(range(n))
for label, env in (("dynamic", None), ("static", "static")):
    if env: os.environ["MDZCUDA_SCHED"] = env
    mdz_b200.render(v, devs)
    for rep in range(2):
        t = time.perf_counter(); raw = mdz_b200.render(v, devs); dt = time.perf_counter() - t
        it = int(np.where(raw > 0, raw, v.depth).astype(np.int64).sum())
        print("%s ndev=%d rep %d: %.1f ms, %.2f G it/s" % (label, n, rep, dt * 1e3, it / dt / 1e9))
