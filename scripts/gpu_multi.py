"""One process, one handle, several GPUs (kb_finalize_multi): end-to-end C2 throughput through kb_feasible_batch with pinned host
buffers (H2D + D2H inside the timer), per number of devices.  usage: python scripts/gpu_multi.py [max_devices] [configs_per_device]"""
import sys, time, ctypes as C
sys.path.insert(0, ".")
import numpy as np
import torch
from klampt_b200 import synth
from klampt_b200.engine import Engine
from klampt_b200._capi import check

maxd = int(sys.argv[1]) if len(sys.argv) > 1 else torch.cuda.device_count()
per = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
only = [int(c) for c in sys.argv[3]] if len(sys.argv) > 3 else [1, 2, 4, 8]          # e.g. "18": one and eight devices
w = synth.world_c2()
base = None
for nd in [n for n in only if n <= maxd]:
    t0 = time.time()
    eng = Engine(w, device=list(range(nd)))
    tf = time.time() - t0
    N = nd * per
    Q = torch.from_numpy(synth.sample_configs(w.robot, N, 7)).pin_memory()
    Qf = Q.to(torch.float32).pin_memory()
    out = torch.empty(N, dtype=torch.uint8).pin_memory()
    bits = torch.zeros((N + 7) // 8, dtype=torch.uint8).pin_memory()
    res = {}
    for name, fn in (("f64 rows, bytes out", lambda: check(eng.lib.kb_feasible_batch(eng.h, C.c_void_p(Q.data_ptr()), N, C.c_void_p(out.data_ptr()), None))),
                     ("f64 rows, bitmask out", lambda: check(eng.lib.kb_feasible_batch_bits(eng.h, C.c_void_p(Q.data_ptr()), N, C.c_void_p(bits.data_ptr())))),
                     ("f32 rows, bytes out", lambda: check(eng.lib.kb_feasible_batch_f32(eng.h, C.c_void_p(Qf.data_ptr()), N, C.c_void_p(out.data_ptr()), None)))):
        for _ in range(3):
            fn()
        t = time.perf_counter()
        for _ in range(8):
            fn()
        dt = (time.perf_counter() - t) / 8
        res[name] = N / dt
    if base is None:
        base = dict(res)
    print("devices %d (finalize %.1f s): " % (nd, tf) + "; ".join("%s %.4g cfg/s (x%.2f, eff %.3f)" % (k, v, v / base[k], v / base[k] / nd) for k, v in res.items()), flush=True)
    eng.close()
