"""Where does the host-path (e2e) time go?  1M fp32, single query."""
import os, sys, time, ctypes as C
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multimodal_audio_search_b200 import SegmentIndex, synth, _native as N
n = 1_000_000
idx = SegmentIndex("fp32", capacity=n); idx.append_synth(1, n, 0, n, n_queries=8, plants=30)
q = synth.raw_queries(1, 0, 64); wa = np.array([0.5]); wb = np.array([0.5]); k = 10
def timeit(f, reps=300):
    for _ in range(20): f()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): f()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps * 1e3
i = [0]
def py_host():
    i[0] += 1; idx.search(q[i[0] % 64], 0.5, 0.5, k=k)
print("python wrapper, host in/out      ms:", round(timeit(py_host), 4))
L = N.lib()
oi = np.empty((1, k), np.int64); of = np.empty((1, k)); oa = np.empty((1, k), np.float32); ob = np.empty((1, k), np.float32)
fl = np.empty((1, k), np.uint8); oc = np.empty(1, np.int32)
p = lambda a: a.ctypes.data_as(C.c_void_p)
args = [p(wa), p(wb), 1, k, 0.1, 0, p(oi), p(of), p(oa), p(ob), p(fl), p(oc), 0, None]
qq = np.ascontiguousarray(q[:1])
def c_host():
    L.cab_search(idx._h, p(qq), 0, *args)
print("direct C-ABI call, host in/out   ms:", round(timeit(c_host), 4))
qd = torch.from_numpy(q).cuda()
def py_dev():
    i[0] += 1; idx.search(qd[i[0] % 64: i[0] % 64 + 1], wa, wb, k=k)
print("torch.ops path, device in/out    ms:", round(timeit(py_dev), 4))
idx.set_option("time_kernels", 1); idx.search(q[0], 0.5, 0.5); print("scan kernel ms:", round(idx.last_scan_ms(), 4))
