import faulthandler, sys, os, ctypes as C
faulthandler.enable()
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from multimodal_audio_search_b200 import SegmentIndex, synth, _native as N
def say(*a): print(*a, flush=True)
idx = SegmentIndex("fp32"); idx.append_synth(1, 5000, 0, 5000, n_queries=2, plants=10)
q = synth.raw_queries(1, 0, 2); wa = np.array([0.5, 0.3]); wb = 1 - wa
say("host", idx.search(q, wa, wb).indices[0][:3])
qd = torch.from_numpy(q).cuda()
L = N.lib()
def dev_search(stream):
    k = 10
    oi = torch.empty((2, k), dtype=torch.int64, device="cuda"); of = torch.empty((2, k), dtype=torch.float64, device="cuda")
    oa = torch.empty((2, k), dtype=torch.float32, device="cuda"); ob = torch.empty((2, k), dtype=torch.float32, device="cuda")
    fl = torch.empty((2, k), dtype=torch.uint8, device="cuda"); oc = torch.empty((2,), dtype=torch.int32, device="cuda")
    p = lambda t: C.c_void_p(t.data_ptr())
    rc = L.cab_search(idx._h, p(qd), 1, wa.ctypes.data_as(C.c_void_p), wb.ctypes.data_as(C.c_void_p), 2, k, 0.1, 0,
                      p(oi), p(of), p(oa), p(ob), p(fl), p(oc), 1, stream)
    torch.cuda.synchronize()
    return rc, oi[0][:3].tolist()
say("ctypes own stream", dev_search(None))
s2 = torch.cuda.Stream()
say("ctypes side stream", dev_search(C.c_void_p(s2.cuda_stream)))
say("ctypes per-thread (0x2)", dev_search(C.c_void_p(2)))
say("ctypes legacy (0x1)", dev_search(C.c_void_p(1)))
say("torch op default stream", idx.search(qd, wa, wb).indices[0][:3].tolist())
