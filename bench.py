"""bench.py - headline benchmark of the 3D-box geometry hot path (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--no-extras]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...

Headline workload (BASELINE.json configs[1]): the training-loss path - IoU assignment + focal / corner / direction
losses, forward AND backward - on a batch of 32 images at 1080p (A = 389 205 anchors, C = 8), 200 GT boxes per image.
A step = FocalLoss forward + backward over one batch.  Metric: G anchor-GT pairs/s (B * A * G pairs per step).
  value : inputs resident in HBM, device time (CUDA events), max over ranks.
  e2e   : the same step through the public API (FocalLoss module + .backward()) starting from pinned HOST buffers, with
          the host->device copies of classification / regression / annotations and the device->host read of the three
          losses inside the timed region.
Multi-GPU: weak scaling - every rank owns its own 32-image shard (global batch 32 N); the only collective is the
5-scalar all-gather of dist.sharded_focal_loss.
Other workloads (configs 3-5: decode+NMS, homography, tracking frame) are reported in "other_workloads" (N = 1 only).
--impl reference: the oracle port of the reference's CPU implementation, timed on the host cores (rank 0 only).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402

H_IMG, W_IMG = 1080, 1920
B_PER_GPU, G_PER_IMG, C_CLS, R_REG = 32, 200, 8, 12
METRIC = "IoU-assign+loss G anchor-GT pairs/s"
UNIT = "G pairs/s"


def _traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of `kernel` from the newest committed ncu capture"""
    best = None
    pdir = os.path.join(ROOT, "profiles")
    for rnd in sorted(os.listdir(pdir)) if os.path.isdir(pdir) else []:
        f = os.path.join(pdir, rnd, "traffic.json")
        if os.path.exists(f):
            with open(f) as fh:
                for name, rec in json.load(fh).items():
                    if name.startswith(kernel):
                        best = rec["dram_bytes_read"] + rec["dram_bytes_write"]
    return best


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clock / throttle-reason samples during the timed region"""
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.samples, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for s in self.samples:
            try:
                sm.append(float(s[0]))
                mx = float(s[1])
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def _dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def _max_over_ranks(value, world, dev):
    if world == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=dev)
    torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    return float(t.item())


# ------------------------------------------------------------------------------------------------- reference arm (CPU)
def cpu_loss_sample(batch, repeats=2, threads=None):
    """Oracle port of the reference's CPU loss path (fwd + bwd) on `batch` 1080p images.  Returns (G pairs/s, seconds)."""
    import synth
    from oracle import losses_oracle
    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    g = synth.gen(1234)
    anc = synth.anchors(H_IMG, W_IMG)
    A = anc.shape[1]
    ann = synth.gt_annotations_3d(batch, G_PER_IMG, H_IMG, W_IMG, g)
    cls, reg = synth.head_outputs(batch, A, C_CLS, R_REG, g)
    best = float("inf")
    for it in range(repeats + 1):
        c, r = cls.clone().requires_grad_(True), reg.clone().requires_grad_(True)
        t0 = time.perf_counter()
        out = losses_oracle.focal_loss(c, r, anc, ann)
        (out[0].sum() + out[1].sum() + out[2].sum()).backward()
        dt = time.perf_counter() - t0
        if it > 0:
            best = min(best, dt)
    pairs = batch * A * G_PER_IMG
    return pairs / best / 1e9, best


def run_reference(args):
    rank, world, _ = _dist_env()
    if rank != 0:
        return
    import synth
    from oracle import losses_oracle
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    batch = 2 if (args.steps + args.warmup) <= 12 else 1   # bounded sample: ~2 s of CPU work per image
    g = synth.gen(1234)
    anc = synth.anchors(H_IMG, W_IMG)
    A = anc.shape[1]
    ann = synth.gt_annotations_3d(batch, G_PER_IMG, H_IMG, W_IMG, g)
    cls, reg = synth.head_outputs(batch, A, C_CLS, R_REG, g)

    def step():
        c, r = cls.clone().requires_grad_(True), reg.clone().requires_grad_(True)
        out = losses_oracle.focal_loss(c, r, anc, ann)
        (out[0].sum() + out[1].sum() + out[2].sum()).backward()

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    value = batch * A * G_PER_IMG / dt / 1e9
    sample = f"{batch} of the {B_PER_GPU} images per step (oracle port of the reference loss, forward+backward, torch CPU ops)"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "training-loss path (BASELINE configs[1]): 1080p, A=389205, G=200/img, C=8, fwd+bwd",
                   "batch_per_step": batch, "note": "bounded sample of the 32-image batch; the path is linear in images"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ------------------------------------------------------------------------------------------------------- GPU arm
def _event_ms(pairs):
    return sum(a.elapsed_time(b) for a, b in pairs)


def other_workloads(dev, hbm_peak):
    """configs 3-5 on one GPU; each entry: metric, value, unit, roofline of its dominant kernel"""
    import synth
    from geom3d_b200 import ops, postprocess, tracker_geometry
    out = []

    def timed(fn, iters, warm=3):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize(dev)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize(dev)
        return e0.elapsed_time(e1) / iters

    g = synth.gen(7)
    anc = synth.anchors(H_IMG, W_IMG).to(dev)
    A = anc.shape[1]
    # ---- config 3: decode + score filter + NMS, batch 64 at 1080p, ~5k pre-NMS boxes per image
    B3 = 64
    cls = torch.rand(B3, A, C_CLS, device=dev) * 0.04
    small = synth.detection_scores(1, A, C_CLS, g)           # 200 objects x 25 anchors scoring U(0.05, 1)
    hot = torch.nonzero(small[0] > 0.04)
    for b in range(B3):
        shift = (hot[:, 0] + 1237 * b) % A
        cls[b, shift.to(dev), hot[:, 1].to(dev)] = small[0][hot[:, 0], hot[:, 1]].to(dev)
    reg3 = torch.randn(B3, A, 12, device=dev) * 0.1
    reg3[..., 8:12] = torch.tensor([-0.5, -0.5, 0.5, 0.5], device=dev) + torch.randn(B3, A, 4, device=dev) * 0.05
    t_dec = timed(lambda: ops.decode3d(anc, reg3), 5)
    dec_bytes = B3 * A * (48 + 80) + 16 * A
    boxes = ops.decode3d(anc, reg3)
    t_pipe = timed(lambda: postprocess.detect_per_class(cls, boxes, box_col=16, score_threshold=0.05), 3, warm=2)
    n_det = postprocess.detect_per_class(cls, boxes, box_col=16, score_threshold=0.05)[0].numel()
    del boxes
    # the path PostProcess3D takes: filter first, decode only the candidates / kept rows (no [B,A,20] tensor)
    t_fused = timed(lambda: postprocess.detect_per_class_fused(cls, reg3, anc, score_threshold=0.05), 5, warm=2)
    n_fused = postprocess.detect_per_class_fused(cls, reg3, anc, score_threshold=0.05)[0].numel()
    out.append({"workload": "config 3: 3D decode + scores>0.05 + per-class NMS 0.5, batch 64 at 1080p, ~5k pre-NMS boxes/img",
                "metric": "decode+NMS img/s", "value": B3 / (t_fused * 1e-3), "unit": "img/s",
                "ms": {"filter->decode-on-the-fly->nms->assemble (PostProcess path)": t_fused,
                       "decode3d (full tensor, BBoxTransform alone)": t_dec,
                       "filter+nms+assemble on the decoded tensor": t_pipe},
                "unfused_img_per_s": B3 / ((t_dec + t_pipe) * 1e-3), "detections": n_fused, "detections_unfused": n_det,
                "note": "the fused tail reads the class scores once (0.80 GB) - its HBM floor is 122 us per batch; "
                        "the NMS chain is latency-bound (SURVEY.md §8d)",
                "roofline": {"kernel": "decode3d_kernel", "bound": "hbm", "achieved": dec_bytes / (t_dec * 1e-3) / 1e9,
                             "peak": hbm_peak, "unit": "GB/s", "frac": dec_bytes / (t_dec * 1e-3) / 1e9 / hbm_peak}})
    # CPU port of the reference on ONE of the 64 images (decode + scores > 0.05 + per-class nms + cat)
    from oracle import decode_oracle, homography_oracle, nms_oracle, tracker_oracle
    cls1, reg1, anc_h = cls[:1].cpu(), reg3[:1].cpu(), anc.cpu()
    try:    # the reference calls torchvision.ops.nms (C++ CPU kernel); fall back to the oracle's restatement of it
        from torchvision.ops import nms as cpu_nms
        nms_name = "torchvision.ops.nms"
    except Exception:   # noqa: BLE001
        cpu_nms, nms_name = nms_oracle.nms, "oracle nms"
    t0 = time.perf_counter()
    dec1 = decode_oracle.decode3d(anc_h, reg1)[0]
    for c in range(C_CLS):
        m = cls1[0, :, c] > 0.05
        if int(m.sum()):
            kept = cpu_nms(dec1[m][:, 16:20].contiguous(), cls1[0, m, c], 0.5)
            _ = dec1[m][kept]
    t_cpu3 = time.perf_counter() - t0
    out[-1]["cpu_baseline"] = {"value": 1.0 / t_cpu3, "unit": "img/s", "cores": os.cpu_count(), "kind": "port",
                               "sample": f"1 of the 64 images (oracle decode3d + scores>0.05 + per-class {nms_name} + gather, "
                                         f"{t_cpu3:.2f} s)"}
    del cls, reg3, cls1, reg1
    torch.cuda.empty_cache()
    # ---- config 4: homography, 10 M states x 18 cameras
    P, Hm = synth.camera_matrices(18)
    Pd, Hd = torch.from_numpy(P).to(dev), torch.from_numpy(Hm).to(dev)
    d = 10_000_000
    st, cam = synth.vehicle_states(d, g)
    st, cam = st.to(dev), cam.to(dev)
    t_s2i = timed(lambda: ops.state_to_im(st, Pd, cam, wrapper=True), 5)
    s2i_bytes = d * (24 + 1 + 128)
    im = ops.state_to_im(st, Pd, cam, wrapper=True)
    hts = st[:, 4].contiguous()
    t_i2s = timed(lambda: ops.im_to_state(im, hts, Hd, cam, wrapper=True), 5)
    # 128 B of float64 image points + height + camera in, 24 B state out per object.  The fused kernel
    # only needs the bottom face (64 B), but HBM delivers most of each 128-byte line anyway (ncu: ~137 B read / object).
    i2s_bytes = d * (128 + 8 + 1 + 24)
    i2s_touched = d * (64 + 8 + 1 + 24)
    del im
    d_all = 1_000_000
    t_all = timed(lambda: ops.state_to_im(st[:d_all], Pd, None, wrapper=True, all_cams=True), 3)
    all_bytes = d_all * (24 + 18 * 128)
    out.append({"workload": "config 4: state_to_im / im_to_state, 10M states, one of 18 cameras each (float64 out)",
                "metric": "state_to_im M states/s", "value": d / (t_s2i * 1e-3) / 1e6, "unit": "M states/s",
                "ms": {"state_to_im_10M": t_s2i, "im_to_state_10M": t_i2s, "state_to_im_all18_1M": t_all},
                "im_to_state_M_per_s": d / (t_i2s * 1e-3) / 1e6,
                "all_cameras_M_state_cams_per_s": d_all * 18 / (t_all * 1e-3) / 1e6,
                "roofline": {"kernel": "state_to_im_kernel", "bound": "hbm", "achieved": s2i_bytes / (t_s2i * 1e-3) / 1e9,
                             "peak": hbm_peak, "unit": "GB/s", "frac": s2i_bytes / (t_s2i * 1e-3) / 1e9 / hbm_peak,
                             "im_to_state_frac": i2s_bytes / (t_i2s * 1e-3) / 1e9 / hbm_peak,
                             "im_to_state_frac_bytes_touched": i2s_touched / (t_i2s * 1e-3) / 1e9 / hbm_peak,
                             "all_cameras_frac": all_bytes / (t_all * 1e-3) / 1e9 / hbm_peak}})
    # CPU port: 1 M of the 10 M states through the wrapper's state_to_im (per-object camera matrices)
    n_cpu = 1_000_000
    st_h, cam_h = st[:n_cpu].cpu(), cam[:n_cpu].cpu().long()
    P_h = torch.from_numpy(P)[cam_h]
    t0 = time.perf_counter()
    homography_oracle.wrapper_state_to_im(st_h, P_h[:, 0], P_h[:, 1])
    t_cpu4 = time.perf_counter() - t0
    out[-1]["cpu_baseline"] = {"value": n_cpu / t_cpu4 / 1e6, "unit": "M states/s", "cores": os.cpu_count(), "kind": "port",
                               "sample": f"1 M of the 10 M states (oracle wrapper_state_to_im, {t_cpu4:.2f} s)"}
    del st, cam, st_h, cam_h, P_h
    torch.cuda.empty_cache()
    # ---- config 5: tracking frame, 2000 objects: association matrix + space NMS + image NMS
    s5, c5 = synth.vehicle_states(2000, g)
    j5 = s5.clone()
    j5[:, :2] += torch.randn(2000, 2, generator=g) * torch.tensor([3.0, 0.5])
    s5, j5, c5 = s5.to(dev), j5.to(dev), c5.to(dev)
    sc5 = torch.rand(2000, device=dev)

    def frame():
        cost = tracker_geometry.association_cost(s5, j5)
        k1 = tracker_geometry.space_nms(j5, sc5, 0.1)          # NMS runs on the detections, the cost is tracked x detected
        corners = ops.state_to_im(j5, Pd, c5, wrapper=True)
        k2 = tracker_geometry.im_nms(corners, sc5, 0.3)
        return cost, k1, k2
    t_frame = timed(frame, 10)
    # the same frame as ONE CUDA graph (tracker_geometry.FrameGeometry: three forked streams, device-side lengths,
    # one 8-byte read of the two keep-list lengths per frame); inputs already on the device
    fg = tracker_geometry.FrameGeometry(Pd, 2000, 2000, phi_space=0.1, phi_im=0.3)
    ref_frame = frame()
    got_frame = fg(s5, j5, sc5, c5)
    assert torch.equal(got_frame["cost"], ref_frame[0]) and torch.equal(got_frame["space_keep"], ref_frame[1]) \
        and torch.equal(got_frame["im_keep"], ref_frame[2]), "graph frame differs from the separate calls"
    t_graph = timed(lambda: fg(s5, j5, sc5, c5), 10)
    s5h, j5h, sc5h = s5.cpu(), j5.cpu(), sc5.cpu()
    P5 = torch.from_numpy(P)[c5.cpu().long()]
    t0 = time.perf_counter()
    tracker_oracle.association_cost(s5h, j5h)
    tracker_oracle.space_nms(j5h, sc5h, 0.1)
    tracker_oracle.im_nms(homography_oracle.wrapper_state_to_im(j5h, P5[:, 0], P5[:, 1]), sc5h, 0.3)
    t_cpu5 = time.perf_counter() - t0
    out.append({"workload": "config 5: 2000 objects: footprint association matrix (f64) + space NMS 0.1 + image NMS 0.3",
                "metric": "tracking-frame geometry frames/s", "value": 1e3 / t_graph, "unit": "frames/s",
                "ms": {"frame (one CUDA graph, FrameGeometry)": t_graph, "frame (separate drop-in calls)": t_frame},
                "separate_calls_frames_per_s": 1e3 / t_frame,
                "note": "launch/latency bound (SURVEY.md §8d); graph frame: 5 small input copies + one graph launch + one "
                        "8-byte read of the keep-list lengths; separate calls: 2 host syncs for the NMS lengths",
                "cpu_baseline": {"value": 1.0 / t_cpu5, "unit": "frames/s", "cores": os.cpu_count(), "kind": "port",
                                 "sample": f"one frame (oracle association_cost + space_nms + state_to_im + im_nms, {t_cpu5:.2f} s)"}})
    # ---- batched Kalman filter (Torch_KF.predict with per-object dt + update of every object)
    nk, Sk, Mk = 1_000_000, 6, 5
    Fk = torch.eye(Sk)
    Hk = torch.zeros(Mk, Sk); Hk[:Mk, :Mk] = torch.eye(Mk)
    Qk, Rk = torch.eye(Sk) * 0.5, torch.eye(Mk) * 0.8
    Xk = torch.randn(nk, Sk, device=dev) * 20
    Ck = torch.randn(nk, Sk, Sk, device=dev)
    Pk = (Ck @ Ck.transpose(1, 2) + torch.eye(Sk, device=dev) * 3.0).contiguous()
    del Ck
    Dk = torch.ones(nk, device=dev)
    Tk = torch.zeros(nk, dtype=torch.float64, device=dev)
    dtk = torch.full((nk,), 1 / 30.0, dtype=torch.float64, device=dev)
    rowsk = torch.arange(nk, device=dev)
    zk = torch.randn(nk, Mk, dtype=torch.float64, device=dev) * 20
    t_pred = timed(lambda: ops.kf_predict_(Xk, Pk, Dk, dtk, Fk, Qk, 1 / 30.0, Tk), 5)
    t_upd = timed(lambda: ops.kf_update_(Xk, Pk, rowsk, zk, Hk, Rk, None), 5)
    pred_bytes = nk * (2 * (Sk + Sk * Sk) * 4 + 4 + 8 + 16)         # X, P in and out, D, dt, T in and out
    upd_bytes = nk * (2 * (Sk + Sk * Sk) * 4 + 8 + Mk * 8)          # X, P in and out, row index, measurement
    from oracle import kf_oracle
    n_cpu = 100_000
    Xc, Pc, Dc, dtc = Xk[:n_cpu].cpu(), Pk[:n_cpu].cpu(), Dk[:n_cpu].cpu(), dtk[:n_cpu].cpu()
    t0 = time.perf_counter()
    Xc, Pc = kf_oracle.predict(Xc, Pc, Dc, dtc, Fk, Qk)
    kf_oracle.update(Xc, Pc, torch.arange(n_cpu), zk[:n_cpu].cpu(), Hk, Rk, torch.zeros(Mk))
    t_cpuk = time.perf_counter() - t0
    out.append({"workload": "SURVEY 8(f)-4: Torch_KF predict (per-object dt) + update, 1M objects, 6 states / 5 measurements",
                "metric": "Kalman predict+update M objects/s", "value": nk / ((t_pred + t_upd) * 1e-3) / 1e6,
                "unit": "M objects/s", "ms": {"predict": t_pred, "update": t_upd},
                "roofline": {"kernel": "kf_predict_kernel", "bound": "hbm", "achieved": pred_bytes / (t_pred * 1e-3) / 1e9,
                             "peak": hbm_peak, "unit": "GB/s", "frac": pred_bytes / (t_pred * 1e-3) / 1e9 / hbm_peak,
                             "update_frac": upd_bytes / (t_upd * 1e-3) / 1e9 / hbm_peak},
                "cpu_baseline": {"value": n_cpu / t_cpuk / 1e6, "unit": "M objects/s", "cores": os.cpu_count(), "kind": "port",
                                 "sample": f"100 k of the 1 M objects (oracle predict + update, torch CPU bmm / inverse, {t_cpuk:.2f} s)"}})
    return out


def run_ours(args):
    rank, world, local = _dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (use --impl reference for the CPU arm)")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=dev)
    import synth
    from geom3d_b200 import dist as gdist
    from geom3d_b200 import losses_impl, ops

    hbm_peak, peak_src = _peaks()
    B = B_PER_GPU
    g = synth.gen(100 + rank)
    # the anchor table as the model gets it (retinanet/model.py:306: self.anchors(img_batch)): the drop-in Anchors module
    # writes it on the device and tags it as the regular pyramid, which lets the loss run its GT-centric assignment
    from geom3d_b200.anchors_impl import Anchors
    anc = Anchors()(torch.zeros(1, 3, H_IMG, W_IMG, device=dev))
    assert torch.equal(anc.cpu(), synth.anchors(H_IMG, W_IMG))
    A = anc.shape[1]
    ann_h = synth.gt_annotations_3d(B, G_PER_IMG, H_IMG, W_IMG, g).pin_memory()
    torch.manual_seed(100 + rank)
    cls_d = (torch.rand(B, A, C_CLS, device=dev) * 0.1).requires_grad_(True)
    reg_d = (torch.randn(B, A, R_REG, device=dev) * 0.1).requires_grad_(True)
    ann_d = ann_h.to(dev)
    ones = torch.ones(3, device=dev)

    def step_device(record=None, trace=None):
        cls_d.grad = None
        reg_d.grad = None
        if record is not None:
            record[0].record()
        if world > 1:
            losses = gdist.sharded_focal_loss(cls_d, reg_d, anc, ann_d, trace_events=trace)
        else:
            losses = losses_impl.focal_loss(cls_d, reg_d, anc, ann_d, trace_events=trace)[0]
        if record is not None:
            record[1].record()
        losses.backward(ones)
        if record is not None:
            record[2].record()
        return losses

    for _ in range(args.warmup):
        step_device()
    torch.cuda.synchronize(dev)
    if world > 1:
        torch.distributed.barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    kev = [[torch.cuda.Event(enable_timing=True) for _ in range(6)] for _ in range(args.steps)]   # per-kernel events
    for tr in kev:
        for e in tr:
            e.record()      # creates the CUDA events outside the timed region
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(dev)
    e0.record()
    for i in range(args.steps):
        losses = step_device(evs[i], kev[i])
    e1.record()
    torch.cuda.synchronize(dev)
    if world > 1:
        torch.distributed.barrier()
    losses = losses.detach().clone()    # drop the autograd graph (its AccumulateGrad nodes would pin the legacy stream)
    ms_total = _max_over_ranks(e0.elapsed_time(e1), world, dev)
    ms_step_eager = ms_total / args.steps
    # ---- the same K steps as ONE CUDA graph launch per step (forward + backward, all kernels and the small torch ops):
    # removes the host-side launch gaps between the six short kernels.  Falls back to the eager figure if capture fails.
    ms_step, mode = ms_step_eager, "eager (one Python call per step)"
    if not args.no_graph:
        graph, ok = None, 1
        try:
            side = torch.cuda.Stream(dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                for _ in range(3):
                    step_device()
            torch.cuda.current_stream(dev).wait_stream(side)
            torch.cuda.synchronize(dev)
            graph = torch.cuda.CUDAGraph()
            # thread_local: the NCCL watchdog thread's CUDA calls must not invalidate the capture (world > 1)
            with torch.cuda.graph(graph, capture_error_mode="thread_local"):
                g_losses = step_device().detach()
            torch.cuda.synchronize(dev)
        except Exception as e:   # noqa: BLE001 - any capture problem: keep the eager number
            ok, mode = 0, f"eager (graph capture failed: {type(e).__name__}: {str(e)[:200]})"
        if world > 1:   # every rank must take the same branch (the replays contain a collective)
            flag = torch.tensor([ok], dtype=torch.int32, device=dev)
            torch.distributed.all_reduce(flag, op=torch.distributed.ReduceOp.MIN)
            if ok and int(flag.item()) == 0:
                mode = "eager (graph capture failed on another rank)"
            ok = int(flag.item())
        if ok:
            for _ in range(args.warmup):
                graph.replay()
            torch.cuda.synchronize(dev)
            if world > 1:
                torch.distributed.barrier()
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            g0.record()
            for _ in range(args.steps):
                graph.replay()
            g1.record()
            torch.cuda.synchronize(dev)
            if world > 1:
                torch.distributed.barrier()
            ms_graph = _max_over_ranks(g0.elapsed_time(g1), world, dev) / args.steps
            same = torch.tensor([1 if torch.equal(g_losses, losses) else 0], dtype=torch.int32, device=dev)
            if world > 1:
                torch.distributed.all_reduce(same, op=torch.distributed.ReduceOp.MIN)
            if int(same.item()) and ms_graph < ms_step:
                ms_step, mode = ms_graph, "CUDA graph replay (one launch per step)"
        if graph is not None:
            graph.reset()
            del graph
    clocks = sampler.stop() if rank == 0 else None
    ms_fwd = _event_ms([(e[0], e[1]) for e in evs]) / args.steps
    ms_bwd = _event_ms([(e[1], e[2]) for e in evs]) / args.steps
    ms_k = [_event_ms([(e[i], e[i + 1]) for e in kev]) / args.steps for i in range(5)]   # prologue, pairs, resolve, stream, positives
    ms_assign = ms_k[0] + ms_k[1] + ms_k[2]
    ms_pos = ms_k[4]
    ms_stream = ms_k[3]
    pairs_per_step = world * B * A * G_PER_IMG
    value = pairs_per_step / (ms_step * 1e-3) / 1e9
    loss_vals = [float(x) for x in losses.detach().cpu()]
    with torch.no_grad():
        per_image_vals = losses_impl.focal_loss(cls_d.detach(), reg_d.detach(), anc, ann_d)[2].cpu().tolist()

    # ---- end to end through the public module, from pinned host buffers
    cls_h = torch.empty((B, A, C_CLS), dtype=torch.float32).pin_memory()
    reg_h = torch.empty((B, A, R_REG), dtype=torch.float32).pin_memory()
    cls_h.copy_(cls_d.detach())
    reg_h.copy_(reg_d.detach())
    module = losses_impl.FocalLoss()
    cls_in = torch.empty_like(cls_d).requires_grad_(True)
    reg_in = torch.empty_like(reg_d).requires_grad_(True)

    def step_e2e():
        cls_in.grad = None
        reg_in.grad = None
        with torch.no_grad():
            cls_in.copy_(cls_h, non_blocking=True)
            reg_in.copy_(reg_h, non_blocking=True)
        ann_in = ann_h.to(dev, non_blocking=True)
        if world > 1:
            l3 = gdist.sharded_focal_loss(cls_in, reg_in, anc, ann_in)
            l3.backward(ones)
            return l3.detach().cpu()
        lc, lr, lv = module(cls_in, reg_in, anc, ann_in)
        (lc + lr + lv).backward()
        return torch.cat((lc, lr, lv)).detach().cpu()      # device->host read of the step's result

    e2e_steps = max(2, min(args.steps, 5))
    step_e2e()
    torch.cuda.synchronize(dev)
    if world > 1:
        torch.distributed.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        host_losses = step_e2e()
    torch.cuda.synchronize(dev)
    t_e2e = _max_over_ranks((time.perf_counter() - t0) / e2e_steps, world, dev)
    e2e_value = pairs_per_step / t_e2e / 1e9
    h2d = cls_h.numel() * 4 + reg_h.numel() * 4 + ann_h.numel() * 4
    del cls_h, reg_h, cls_in, reg_in

    if rank == 0:
        # algorithmic bytes of the forward launches: assign_codes_kernel (anchors + GT in, codes + the
        # zero-filled dreg out; issue-bound), positives_kernel (negligible) and the HBM-bound focal_stream_kernel (cls +
        # codes in, dcls out; the dominant kernel).  G3D_LOSS_FUSED=1: the experimental single persistent kernel for
        # both.  backward (dcls already written): the positive rows only.
        fused = os.environ.get("G3D_LOSS_FUSED", "0") == "1"
        stream_bytes = B * A * (C_CLS * 4 + 4 + C_CLS * 4)          # cls + codes in, dcls out
        assign_bytes = B * A * (4 + R_REG * 4) + A * 16             # codes + the zero-filled dreg out, anchors in
        fwd_bytes = stream_bytes + assign_bytes + ann_h.numel() * 4
        bwd_bytes = int(sum(p[3] for p in per_image_vals)) * (R_REG * 4 * 2 + 4 + 21 * 4)
        if fused:
            knames = ("focal_fused_kernel", "positives_finalize_kernel", "-")
            dom, dom_bytes, dom_ms = "focal_fused_kernel", stream_bytes + assign_bytes, ms_assign
        else:
            knames = ("assign_codes_kernel", "positives_kernel", "focal_stream_kernel")
            # the two long launches are within a few percent of each other: the dominant one is whichever measured longer
            # in THIS run (assign_codes_kernel is FP32-issue-bound - its HBM fraction says how much bandwidth it leaves
            # idle, not how good it is -, focal_stream_kernel is the HBM-bound one)
            dom, dom_bytes, dom_ms = max((("focal_stream_kernel", stream_bytes, ms_stream),
                                          ("assign_codes_kernel", assign_bytes, ms_assign)), key=lambda t: t[2])
        achieved = dom_bytes / (dom_ms * 1e-3) / 1e9
        per_kernel = [] if fused else [
            {"kernel": "focal_stream_kernel", "ms": ms_stream, "share_of_forward": ms_stream / ms_fwd, "bytes": stream_bytes,
             "GBps": stream_bytes / (ms_stream * 1e-3) / 1e9, "frac": stream_bytes / (ms_stream * 1e-3) / 1e9 / hbm_peak,
             "limiter": "hbm", "traffic": _traffic("focal_stream_kernel")},
            {"kernel": "assign_codes_kernel", "ms": ms_assign, "share_of_forward": ms_assign / ms_fwd, "bytes": assign_bytes,
             "GBps": assign_bytes / (ms_assign * 1e-3) / 1e9, "frac": assign_bytes / (ms_assign * 1e-3) / 1e9 / hbm_peak,
             "limiter": "fp32 issue (ncu: 75 % of issue slots busy, anchors + GT L2-resident)",
             "traffic": _traffic("assign_codes_kernel")}]
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": "training-loss path (BASELINE configs[1]): FocalLoss fwd+bwd, 1080p, A=389205, "
                                   "G=200 GT/img, C=8, 12-d regression", "batch_per_gpu": B, "global_batch": B * world,
                       "parallelism": f"images sharded over {world} GPU(s), 5-scalar all-gather only",
                       "l2": "inputs (1.0 GB/step) exceed the 126 MB L2; no flush needed",
                       "launch_mode": mode, "ms_per_step_eager": ms_step_eager},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 12,
                    "ms_per_step": t_e2e * 1e3, "steps": e2e_steps},
            # gt_prepare, focal_fused, positives_finalize, focal_cls_grad, positives (bwd) per step
            "gpu_launches": (5 if fused else 6) * args.steps,
            "roofline": {"kernel": dom, "bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                         "frac": achieved / hbm_peak, "traffic": _traffic(dom), "peak_source": peak_src,
                         "kernels": per_kernel,
                         "ms": {"forward": ms_fwd, "backward": ms_bwd, knames[0]: ms_assign, knames[1]: ms_pos,
                                knames[2]: ms_stream},
                         # SURVEY.md §8(d) "fused fwd+bwd": 160 B per (image, anchor) - cls + reg in, dcls + dreg out - plus
                         # the anchors once: the whole step against the HBM roofline (this design never reads reg except
                         # for the positives, so its own traffic is lower: see forward / backward below)
                         "step_on_survey_bytes": {"bytes": B * A * 160 + A * 16,
                                                  "GBps": (B * A * 160 + A * 16) / (ms_step * 1e-3) / 1e9,
                                                  "frac": (B * A * 160 + A * 16) / (ms_step * 1e-3) / 1e9 / hbm_peak},
                         "forward": {"bytes": fwd_bytes, "GBps": fwd_bytes / (ms_fwd * 1e-3) / 1e9,
                                     "frac": fwd_bytes / (ms_fwd * 1e-3) / 1e9 / hbm_peak},
                         "backward": {"bytes": bwd_bytes, "GBps": bwd_bytes / (ms_bwd * 1e-3) / 1e9,
                                      "frac": bwd_bytes / (ms_bwd * 1e-3) / 1e9 / hbm_peak}},
            "losses": loss_vals, "e2e_losses": [float(x) for x in host_losses], "ms_kernels": ms_k,
        }

    # forward-only pass (validation loss, no gradient buffers): GT-centric assignment on the tagged pyramid table against
    # the anchor-centric kernel on an untagged copy of the same table (same codes, same losses)
    fwd_only = None
    if rank == 0 and world == 1:
        anc_plain = anc.clone()

        def fwd_time(table):
            with torch.no_grad():
                for _ in range(3):
                    out = ops.focal_loss_forward(cls_d, reg_d, table, ann_d)
                torch.cuda.synchronize(dev)
                a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a0.record()
                for _ in range(10):
                    out = ops.focal_loss_forward(cls_d, reg_d, table, ann_d)
                a1.record()
                torch.cuda.synchronize(dev)
            return a0.elapsed_time(a1) / 10, out
        t_gt, o_gt = fwd_time(anc)
        t_an, o_an = fwd_time(anc_plain)
        assert o_gt["gt_centric"] and not o_an["gt_centric"] and torch.equal(o_gt["assign"], o_an["assign"])
        fwd_only = {"ms_gt_centric": t_gt, "ms_anchor_centric": t_an,
                    "G_pairs_per_s_gt_centric": pairs_per_step / (t_gt * 1e-3) / 1e9}
        del anc_plain, o_gt, o_an
        line["forward_only"] = fwd_only
    del cls_d, reg_d
    torch.cuda.empty_cache()
    if rank == 0:
        if world == 1:
            cpu_value, cpu_s = cpu_loss_sample(8)
            line["cpu_baseline"] = {"value": cpu_value, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                                    "sample": f"8 of the 32 images (oracle port, forward+backward, best of 2 after 1 warm-up, {cpu_s:.2f} s/run)"}
            if not args.no_extras:
                try:
                    line["other_workloads"] = other_workloads(dev, hbm_peak)
                except Exception as e:  # the headline line must still be printed
                    line["other_workloads"] = {"error": f"{type(e).__name__}: {e}"}
        print(json.dumps(line), flush=True)
    if world > 1:
        # A captured graph holds NCCL work; tearing the process group down with it alive can block.  Everything is
        # measured and printed: synchronise, meet the other ranks once more, and leave without the teardown.
        torch.cuda.synchronize(dev)
        torch.distributed.barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-extras", action="store_true", help="skip the config 3-5 workloads")
    ap.add_argument("--no-graph", action="store_true", help="time the eager step only (no CUDA graph replay)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        args.warmup = max(args.warmup, 3)
        run_ours(args)


if __name__ == "__main__":
    main()
