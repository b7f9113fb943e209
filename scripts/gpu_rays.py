"""ray kernel variants on the C6 workload (640 x 480 images of the C2 world): launch shape x camera tiling"""
import os, sys, time, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np
import torch
from klampt_b200 import synth, sensing
from klampt_b200.engine import Engine
from klampt_b200._capi import check, KbCamera

spec = synth.world_c2()
eng = Engine(spec)
q = np.ascontiguousarray(synth.sample_configs(spec.robot, 1, 77)[0])
W, H = 640, 480
cams, dev = [], []
for k in range(8):
    a = 2.0 * math.pi * k / 8
    eye = np.array([3.2 * math.cos(a), 3.2 * math.sin(a), 1.3])
    fwd = np.array([0.0, 0.0, 0.5]) - eye; fwd /= np.linalg.norm(fwd)
    right = np.cross(fwd, [0.0, 0.0, 1.0]); right /= np.linalg.norm(right)
    down = np.cross(fwd, right)
    cam = sensing.CameraSensor(W, H, zmin=0.1, zmax=8.0, Tsensor=synth.make_T(np.stack([right, down, fwd], axis=1), eye))
    kc = KbCamera(); kc.pose[:] = list(cam.Tsensor); kc.fx, kc.fy, kc.cx, kc.cy = cam.viewport(); kc.zmin, kc.zmax, kc.xres, kc.yres = cam.zmin, cam.zmax, W, H
    cams.append(kc)
    dev.append(torch.from_numpy(np.ascontiguousarray(cam.rays()[0])).cuda())
n = W * H
d_id, d_dist = torch.empty(n, dtype=torch.int32, device="cuda"), torch.empty(n, dtype=torch.float64, device="cuda")
h_depth, h_id = torch.empty(n, dtype=torch.float32).pin_memory(), torch.empty(n, dtype=torch.int32).pin_memory()
qp = C.c_void_p(q.ctypes.data)
ref = None
for variant in (0, 1):
    for tile in (0, 1):
        eng.set_option("ray_variant", variant); eng.set_option("ray_tile", tile)
        def dev_step(k):
            check(eng.lib.kb_raycast_batch_device(eng.h, qp, C.c_void_p(dev[k % 8].data_ptr()), n, None, C.c_void_p(d_id.data_ptr()), C.c_void_p(d_dist.data_ptr()), None))
        def cam_step(k):
            check(eng.lib.kb_camera_depth(eng.h, qp, C.byref(cams[k % 8]), None, C.c_void_p(h_depth.data_ptr()), C.c_void_p(h_id.data_ptr())))
        res = []
        for fn, sync in ((dev_step, True), (cam_step, False)):
            for k in range(8): fn(k)
            eng.synchronize(); torch.cuda.synchronize()
            t0 = time.perf_counter()
            for k in range(40): fn(k)
            eng.synchronize()
            res.append((time.perf_counter() - t0) / 40 * 1e3)
        cam_step(0)
        img = h_depth.clone()
        if ref is None: ref = img
        print("variant %d tile %d: rays device-resident %.3f ms / image, camera e2e %.3f ms / image, same image: %s" % (variant, tile, res[0], res[1], bool(torch.equal(img, ref))), flush=True)
