"""TRT-like stand-in of the reference front-end: what Omni-swarm's swarm_loop does per keyframe, with the TensorRT fp16
engines replaced by the same networks in PyTorch/cuDNN fp16 (TensorRT 7 engines cannot be built or run here).

It keeps the reference's STRUCTURE, which is what the product path removes (paths relative to /root/reference):
  * one image at a time, batch 1 (swarm_loop/src/loop_cam.cpp:541-556: SuperPoint, then NetVLAD, per image);
  * TensorRTInferenceGeneric::doInference (swarm_loop/src/tensorrt_generic.cpp:58-75): H2D of the fp32 image, enqueue,
    D2H of EVERY output binding (semi H*W*4 B + desc 256*H/8*W/8*4 B), cudaStreamSynchronize -- per call;
  * CPU pre-processing u8 -> f32 * 1/255 (superpoint_tensorrt.cpp:127) and CPU post-processing: getKeyPoints + NMS2
    (:164-189,237-310; here the plain-C restatement oracle/c/nms2_ref.c), computeDescriptors with libtorch on one thread
    (:98,192-230), PCA as a GEMM (:221);
  * cv::BFMatcher(NORM_L2, crossCheck) for the stereo and loop matches (loop_cam.cpp:147-150, loop_detector.cpp:564-567);
  * faiss::IndexFlatIP scan on the CPU (loop_detector.cpp:213) = one sgemv over the database.
This is BASELINE code (bench.py's `trt_like_baseline` leg); it uses cuDNN and the oracle's CPU stages on purpose and is
never imported by the product.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F


class TrtLikeFrontend:
    def __init__(self, sp_w: dict, nv_w: dict, pca_comp: np.ndarray, pca_mean: np.ndarray, db: np.ndarray, H: int, W: int,
                 thres: float = 0.015, max_num: int = 200, device: str = "cuda"):
        from omniswarm_b200 import synth
        self.H, self.W, self.thres, self.max_num, self.dev = H, W, thres, max_num, torch.device(device)
        torch.backends.cudnn.benchmark = True
        half = lambda v: torch.from_numpy(np.ascontiguousarray(v)).to(self.dev).half()
        self.sp = {k: (half(v).contiguous(memory_format=torch.channels_last) if v.ndim == 4 else half(v)) for k, v in sp_w.items()}
        self.nv = {k: (half(v).contiguous(memory_format=torch.channels_last) if v.ndim == 4 else half(v)) for k, v in nv_w.items()}
        self.nv_blocks, self.nv_scale = synth.NV_BLOCKS, float(synth.NV_INPUT_SCALE)
        self.comp_t = torch.from_numpy(np.ascontiguousarray(pca_comp.T.astype(np.float32)))     # [256,64]
        self.mean = torch.from_numpy(pca_mean.astype(np.float32))[None]
        self.db = np.ascontiguousarray(db, np.float32)
        self.stream = torch.cuda.Stream(self.dev)
        # device input + pinned host output buffers, allocated once (tensorrt_generic.cpp:99-120)
        self.d_in = torch.empty(1, 1, H, W, dtype=torch.float32, device=self.dev)
        self.h_semi = torch.empty(H, W, dtype=torch.float32).pin_memory()
        self.h_desc = torch.empty(256, H // 8, W // 8, dtype=torch.float32).pin_memory()
        self.h_nv = torch.empty(4096, dtype=torch.float32).pin_memory()
        import cv2
        self.bf = cv2.BFMatcher(cv2.NORM_L2, True)
        self.bytes_h2d = self.bytes_d2h = 0

    # ---- the two "engines" (fp16 inside, fp32 bindings) ----
    def _sp_engine(self, x):
        t, relu = self.sp, F.relu
        conv = lambda x, n, p: F.conv2d(x, t[n + ".weight"], t[n + ".bias"], padding=p)
        x = x.half().contiguous(memory_format=torch.channels_last)
        x = relu(conv(x, "conv1a", 1)); x = relu(conv(x, "conv1b", 1)); x = F.max_pool2d(x, 2, 2)
        x = relu(conv(x, "conv2a", 1)); x = relu(conv(x, "conv2b", 1)); x = F.max_pool2d(x, 2, 2)
        x = relu(conv(x, "conv3a", 1)); x = relu(conv(x, "conv3b", 1)); x = F.max_pool2d(x, 2, 2)
        x = relu(conv(x, "conv4a", 1)); x = relu(conv(x, "conv4b", 1))
        semi = conv(relu(conv(x, "convPa", 1)), "convPb", 0)
        desc = conv(relu(conv(x, "convDa", 1)), "convDb", 0)
        desc = desc / torch.norm(desc, p=2, dim=1, keepdim=True)
        semi = torch.softmax(semi.float(), 1)[:, :64].permute(0, 2, 3, 1)
        Hc, Wc = semi.shape[1], semi.shape[2]
        semi = semi.reshape(-1, Hc, Wc, 8, 8).permute(0, 1, 3, 2, 4).reshape(-1, Hc * 8, Wc * 8)
        return semi[0].float(), desc[0].float()

    def _nv_engine(self, x):
        t = self.nv
        relu6 = lambda v: torch.clamp(v, 0.0, 6.0)
        x = (x * self.nv_scale).half().contiguous(memory_format=torch.channels_last)
        x = relu6(F.conv2d(x, t["conv0.weight"], t["conv0.bias"], stride=2, padding=1))
        for i, (ci, co, s) in enumerate(self.nv_blocks):
            x = relu6(F.conv2d(x, t[f"b{i}.dw.weight"], t[f"b{i}.dw.bias"], stride=s, padding=1, groups=ci))
            x = relu6(F.conv2d(x, t[f"b{i}.pw.weight"], t[f"b{i}.pw.bias"]))
        x = F.conv2d(x, t["proj.weight"], t["proj.bias"]).float()
        x = x - x.mean(dim=(2, 3), keepdim=True)
        x = x / torch.clamp(torch.norm(x, dim=1, keepdim=True), min=1e-12)
        a = torch.softmax(F.conv2d(x.half(), t["assign.weight"], t["assign.bias"]).float(), 1)
        D, K = x.shape[1], a.shape[1]
        xf, af = x.reshape(D, -1), a.reshape(K, -1)
        vlad = af @ xf.t() - af.sum(1, keepdim=True) * t["centroids"].float()
        vlad = vlad / torch.clamp(torch.norm(vlad, dim=1, keepdim=True), min=1e-12)
        v = vlad.reshape(-1)
        return v / torch.clamp(torch.norm(v), min=1e-12)

    # ---- doInference: H2D, enqueue, D2H of every binding, synchronize (tensorrt_generic.cpp:58-75) ----
    def _do_inference(self, host_f32: torch.Tensor, engine, outs):
        with torch.cuda.stream(self.stream), torch.no_grad():
            self.d_in.copy_(host_f32.view(1, 1, self.H, self.W), non_blocking=True)
            res = engine(self.d_in)
            res = res if isinstance(res, tuple) else (res,)
            for o, r in zip(outs, res):
                o.copy_(r, non_blocking=True)
        self.stream.synchronize()
        self.bytes_h2d += host_f32.numel() * 4
        self.bytes_d2h += sum(o.numel() * 4 for o in outs)

    def superpoint(self, img_u8: np.ndarray):
        """SuperPointTensorRT::inference (superpoint_tensorrt.cpp:117-162)"""
        from oracle import nms2_c
        x = torch.from_numpy(img_u8.astype(np.float32) * np.float32(1.0 / 255.0))           # :127 on the CPU
        self._do_inference(x, self._sp_engine, (self.h_semi, self.h_desc))
        kpts, _ = nms2_c.get_keypoints(self.h_semi.numpy(), self.thres, self.max_num)        # :164-189,237-310
        n = len(kpts)
        if n == 0:
            return kpts, np.zeros((0, 64), np.float32)
        fk = torch.from_numpy(kpts)                                                          # :192-230 (libtorch, 1 thread)
        grid = torch.empty(1, 1, n, 2)
        grid[0, 0, :, 0] = 2.0 * fk[:, 0] / self.W - 1
        grid[0, 0, :, 1] = 2.0 * fk[:, 1] / self.H - 1
        s = F.grid_sample(self.h_desc[None], grid, mode="bilinear", padding_mode="zeros", align_corners=False)[0, :, 0]
        s = s / torch.norm(s, 2, 1, keepdim=True)
        d = (s.t() - self.mean) @ self.comp_t
        return kpts, d.numpy()

    def netvlad(self, img_u8: np.ndarray):
        """MobileNetVLADTensorRT::inference (mobilenetvlad_tensorrt.cpp:4-15): u8 -> f32 unscaled"""
        self._do_inference(torch.from_numpy(img_u8.astype(np.float32)), self._nv_engine, (self.h_nv,))
        return self.h_nv.numpy().copy()

    def keyframe(self, up: np.ndarray, down: np.ndarray):
        """One 4-view fisheye keyframe: loop_cam.cpp:341-523 per direction, then the database work of
        loop_detector.cpp:150-287 and the per-direction loop match (:539-567)."""
        H = self.H
        dirs = []
        for d in range(up.shape[0]):
            u = up[d].copy(); u[H * 3 // 4:] = 0                     # loop_cam.cpp:535-538
            dn = down[d].copy(); dn[H * 3 // 4:] = 0
            ku, du = self.superpoint(u)
            g = self.netvlad(u)
            kd, dd = self.superpoint(dn)
            m = self.bf.match(du, dd) if len(ku) > 10 and len(kd) else []      # :385-391
            dirs.append((ku, du, g, m))
        q = dirs[min(1, len(dirs) - 1)][2]
        scores = self.db @ q                                          # IndexFlatIP::search: one sgemv over the rows
        k = 10
        top = np.argpartition(-scores, k)[:k]
        top = top[np.argsort(-scores[top], kind="stable")]
        n_match = 0
        for d in range(len(dirs)):                                    # loop_detector.cpp:455-465,564-567
            a, b = dirs[d][1], dirs[(d + 1) % len(dirs)][1]
            if len(a) and len(b):
                n_match += len(self.bf.match(a, b))
        return int(top[0]), [len(x[0]) for x in dirs], n_match
