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
5-scalar all-gather of dist.sharded_focal_loss.  "strong_scaling" (N > 1): the 32-image batch of BASELINE configs[1]
sharded over the ranks (32 / N images each).
Other workloads (configs 3-5: decode+NMS, homography, tracking frame) are reported in "other_workloads"; decode+NMS and
the homography shard by image / state range at N > 1 (no collective), the tracking frame is replicas only.
--impl reference: the oracle port of the reference's CPU implementation, timed on the host cores (rank 0 only).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

# stdout carries ONE JSON line: whatever NCCL prints goes to stderr.  (NCCL honours NCCL_DEBUG_FILE only above the
# VERSION level - at VERSION its banner goes to stdout - so VERSION is raised to WARN, which prints the same banner.)
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
    os.environ["NCCL_DEBUG"] = "WARN"

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


def _timed(fn, iters, dev, warm=3):
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


def _best_of(fn, repeats=3):
    """CPU baseline timing: one warm-up call, then the best of `repeats`"""
    fn()
    best = float("inf")
    for _ in range(repeats):
        t0 = time.perf_counter()
        fn()
        best = min(best, time.perf_counter() - t0)
    return best


def _bind_to_gpu_numa_node(local):
    """Pin this process (and with it the pinned host buffers it allocates afterwards: first touch) to the CPU cores of
    the GPU's NUMA node, so that 8 ranks do not all pull their host->device copies through one socket.  Returns a
    description for the bench line; silently does nothing where sysfs / affinity are not available."""
    try:
        pr = torch.cuda.get_device_properties(local)
        if all(hasattr(pr, k) for k in ("pci_domain_id", "pci_bus_id", "pci_device_id")):
            bus = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        else:
            out = subprocess.run(["nvidia-smi", f"--id={local}", "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                                 stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, timeout=20).stdout.strip()
            bus = out.lower()
            if bus.count(":") == 2 and len(bus.split(":")[0]) == 8:
                bus = bus[4:]                                        # 00000000:1B:00.0 -> 0000:1b:00.0
        base = f"/sys/bus/pci/devices/{bus}"
        node = int(open(base + "/numa_node").read())
        cpulist = open(base + "/local_cpulist").read().strip()
        cpus = set()
        for part in cpulist.split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & set(os.sched_getaffinity(0))
        if allowed:
            os.sched_setaffinity(0, allowed)
        return {"numa_node": node, "cpus_bound": len(allowed), "pci": bus}
    except Exception as e:  # noqa: BLE001
        return {"numa_node": None, "note": f"not bound ({type(e).__name__})"}


def other_workloads(dev, hbm_peak, rank, world):
    """configs 3-5 (+ the Kalman filter).  Work shards over the ranks with no collective: images for
    decode + NMS (64 -> 64 / N per GPU), state ranges for the homography (10 M -> 10 M / N), replicas only for the
    2000-object tracking frame and the filter (rank 0).  Every figure: device time, max over ranks, whole-job rate."""
    import synth
    from geom3d_b200 import ops, postprocess, tracker_geometry
    out = []
    timed = lambda fn, iters, warm=3: _timed(fn, iters, dev, warm)   # noqa: E731
    g = synth.gen(7)
    anc = synth.anchors(H_IMG, W_IMG).to(dev)
    A = anc.shape[1]
    # ---- config 3: decode + score filter + NMS, batch 64 at 1080p, ~5k pre-NMS boxes per image
    B3_total = 64
    B3 = B3_total // world
    cls = torch.rand(B3, A, C_CLS, device=dev) * 0.04
    small = synth.detection_scores(1, A, C_CLS, g)           # 200 objects x 25 anchors scoring U(0.05, 1) - image 0 of the parity test
    hot = torch.nonzero(small[0] > 0.04)
    for b in range(B3):
        shift = (hot[:, 0] + 1237 * (b + rank * B3)) % A
        cls[b, shift.to(dev), hot[:, 1].to(dev)] = small[0][hot[:, 0], hot[:, 1]].to(dev)
    reg3 = torch.randn(B3, A, 12, device=dev) * 0.1
    reg3[..., 8:12] = torch.tensor([-0.5, -0.5, 0.5, 0.5], device=dev) + torch.randn(B3, A, 4, device=dev) * 0.05
    # the path PostProcess3D takes: filter first, decode only the candidates / kept rows (no [B,A,20] tensor)
    t_fused = _max_over_ranks(timed(lambda: postprocess.detect_per_class_fused(cls, reg3, anc, score_threshold=0.05), 20, warm=3), world, dev)
    n_fused = postprocess.detect_per_class_fused(cls, reg3, anc, score_threshold=0.05)[0].numel()
    # its HBM-bound launch on its own (compact8_kernel: every class score read once), and the stand-alone decode
    thr3 = torch.full((B3 * C_CLS,), 0.05, dtype=torch.float32, device=dev)
    t_filter = timed(lambda: ops.filter_compact(cls, B3, C_CLS, A, A * C_CLS, thr3, 16384), 10)
    filt_bytes = B3 * A * C_CLS * 4
    # the device side of the tail alone (score filter + per-segment kernel + offsets; no assembly, no host read), on
    # reused buffers as the PostProcess path runs it; and the general chain (gather -> sort -> greedy NMS launches)
    plan = ops.detect_tail(cls, B3, C_CLS, A, A * C_CLS, thr3, 16384, anc, reg3, 0.5, short=True)
    assert int(plan["summary"][2]) == 0, "config 3 must run on the short-segment path"
    t_lib = timed(lambda: ops.detect_tail(cls, B3, C_CLS, A, A * C_CLS, thr3, 16384, anc, reg3, 0.5, short=True, reuse=plan), 10)
    t_lib_general = timed(lambda: ops.detect_tail(cls, B3, C_CLS, A, A * C_CLS, thr3, 16384, anc, reg3, 0.5, short=False, reuse=plan), 10)
    max_count = int(plan["count"].max())
    del plan
    entry = {"workload": f"config 3: 3D decode + scores>0.05 + per-class NMS 0.5, batch {B3_total} at 1080p, ~5k pre-NMS boxes/img, "
                         f"{B3} images per GPU",
             "metric": "decode+NMS img/s", "value": B3_total / (t_fused * 1e-3), "unit": "img/s", "n_gpus": world,
             "ms": {"filter->decode-on-the-fly->nms->assemble (PostProcess path)": t_fused, "compact8_kernel (score filter) alone": t_filter,
                    "g3d_detect_tail_short alone (filter + per-segment kernel + offsets, device side)": t_lib,
                    "g3d_detect_tail alone (general chain: filter, gather, sort, greedy NMS, scan)": t_lib_general},
             "detections_rank0": n_fused, "largest_segment_rank0": max_count,
             "note": "the tail reads the class scores once - that launch (compact8_kernel) is the HBM-bound one; behind it ONE "
                     "launch does gather + decode + sort + NMS per (image, class) segment in shared memory (tail_short_kernel, "
                     "latency-bound: ~30 us per CTA, 1.7 waves of 1024-thread CTAs), then offsets and the assembly; the rest of "
                     "the per-call time is the host: one 16-byte read-back and the Python between two calls",
             "roofline": {"kernel": "compact8_kernel", "bound": "hbm", "achieved": filt_bytes / (t_filter * 1e-3) / 1e9,
                          "peak": hbm_peak, "unit": "GB/s", "frac": filt_bytes / (t_filter * 1e-3) / 1e9 / hbm_peak,
                          "share_of_tail": t_filter / t_fused, "segment_kernel_and_offsets_ms": t_lib - t_filter,
                          "assembly_and_host_ms": t_fused - t_lib}}
    if rank == 0 and world == 1:
        t_dec = timed(lambda: ops.decode3d(anc, reg3), 5)
        dec_bytes = B3 * A * (48 + 80) + 16 * A
        boxes = ops.decode3d(anc, reg3)
        t_pipe = timed(lambda: postprocess.detect_per_class(cls, boxes, box_col=16, score_threshold=0.05), 3, warm=2)
        del boxes
        entry["ms"]["decode3d (full tensor, BBoxTransform alone)"] = t_dec
        entry["ms"]["filter+nms+assemble on the decoded tensor"] = t_pipe
        entry["unfused_img_per_s"] = B3 / ((t_dec + t_pipe) * 1e-3)
        entry["decode3d_roofline"] = {"bytes": dec_bytes, "GBps": dec_bytes / (t_dec * 1e-3) / 1e9,
                                      "frac": dec_bytes / (t_dec * 1e-3) / 1e9 / hbm_peak}
        # CPU port of the reference on ONE of the 64 images (decode + scores > 0.05 + per-class nms + cat)
        from oracle import decode_oracle, nms_oracle
        cls1, reg1, anc_h = cls[:1].cpu(), reg3[:1].cpu(), anc.cpu()
        try:    # the reference calls torchvision.ops.nms (C++ CPU kernel); fall back to the oracle's restatement of it
            from torchvision.ops import nms as cpu_nms
            nms_name = "torchvision.ops.nms"
        except Exception:   # noqa: BLE001
            cpu_nms, nms_name = nms_oracle.nms, "oracle nms"

        def cpu3():
            dec1 = decode_oracle.decode3d(anc_h, reg1)[0]
            for c in range(C_CLS):
                m = cls1[0, :, c] > 0.05
                if int(m.sum()):
                    kept = cpu_nms(dec1[m][:, 16:20].contiguous(), cls1[0, m, c], 0.5)
                    _ = dec1[m][kept]
        t_cpu3 = _best_of(cpu3)
        entry["cpu_baseline"] = {"value": 1.0 / t_cpu3, "unit": "img/s", "cores": os.cpu_count(), "kind": "port",
                                 "sample": f"1 of the 64 images (oracle decode3d + scores>0.05 + per-class {nms_name} + gather; "
                                           f"best of 3 after a warm-up, {t_cpu3:.3f} s)"}
    out.append(entry)
    del cls, reg3
    torch.cuda.empty_cache()
    # ---- config 4: homography, 10 M states x 18 cameras, state ranges per rank
    P, Hm = synth.camera_matrices(18)
    Pd, Hd = torch.from_numpy(P).to(dev), torch.from_numpy(Hm).to(dev)
    d_total = 10_000_000
    d = d_total // world
    st, cam = synth.vehicle_states(d, synth.gen(7 + rank))
    st, cam = st.to(dev), cam.to(dev)
    t_s2i = _max_over_ranks(timed(lambda: ops.state_to_im(st, Pd, cam, wrapper=True), 5), world, dev)
    s2i_bytes = d * (24 + 1 + 128)
    im = ops.state_to_im(st, Pd, cam, wrapper=True)
    hts = st[:, 4].contiguous()
    t_i2s = _max_over_ranks(timed(lambda: ops.im_to_state(im, hts, Hd, cam, wrapper=True), 5), world, dev)
    # 128 B of float64 image points + height + camera in, 24 B state out per object; the fused kernel
    # needs only the bottom face (64 B of the 128)
    i2s_bytes = d * (128 + 8 + 1 + 24)
    i2s_touched = d * (64 + 8 + 1 + 24)
    del im
    entry = {"workload": f"config 4: state_to_im / im_to_state, {d_total // 1_000_000}M states, one of 18 cameras each (float64 out), "
                         f"{d} states per GPU",
             "metric": "state_to_im M states/s", "value": d_total / (t_s2i * 1e-3) / 1e6, "unit": "M states/s", "n_gpus": world,
             "ms": {"state_to_im": t_s2i, "im_to_state": t_i2s},
             "im_to_state_M_per_s": d_total / (t_i2s * 1e-3) / 1e6,
             "roofline": {"kernel": "state_to_im_kernel", "bound": "hbm", "achieved": s2i_bytes / (t_s2i * 1e-3) / 1e9,
                          "peak": hbm_peak, "unit": "GB/s", "frac": s2i_bytes / (t_s2i * 1e-3) / 1e9 / hbm_peak,
                          "im_to_state_frac": i2s_bytes / (t_i2s * 1e-3) / 1e9 / hbm_peak,
                          "im_to_state_frac_bytes_touched": i2s_touched / (t_i2s * 1e-3) / 1e9 / hbm_peak}}
    if rank == 0 and world == 1:
        d_all = 1_000_000
        t_all = timed(lambda: ops.state_to_im(st[:d_all], Pd, None, wrapper=True, all_cams=True), 3)
        all_bytes = d_all * (24 + 18 * 128)
        entry["ms"]["state_to_im_all18_1M"] = t_all
        entry["all_cameras_M_state_cams_per_s"] = d_all * 18 / (t_all * 1e-3) / 1e6
        entry["roofline"]["all_cameras_frac"] = all_bytes / (t_all * 1e-3) / 1e9 / hbm_peak
        from oracle import homography_oracle
        n_cpu = 1_000_000
        st_h, cam_h = st[:n_cpu].cpu(), cam[:n_cpu].cpu().long()
        P_h = torch.from_numpy(P)[cam_h]
        t_cpu4 = _best_of(lambda: homography_oracle.wrapper_state_to_im(st_h, P_h[:, 0], P_h[:, 1]))
        entry["cpu_baseline"] = {"value": n_cpu / t_cpu4 / 1e6, "unit": "M states/s", "cores": os.cpu_count(), "kind": "port",
                                 "sample": f"1 M of the 10 M states (oracle wrapper_state_to_im; best of 3 after a warm-up, {t_cpu4:.3f} s)"}
    out.append(entry)
    del st, cam
    torch.cuda.empty_cache()
    if rank != 0:
        return out
    # ---- config 5: tracking frame, 2000 objects: association matrix + space NMS + image NMS (replicas only: rank 0)
    from oracle import homography_oracle, tracker_oracle
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

    def cpu5():
        tracker_oracle.association_cost(s5h, j5h)
        tracker_oracle.space_nms(j5h, sc5h, 0.1)
        tracker_oracle.im_nms(homography_oracle.wrapper_state_to_im(j5h, P5[:, 0], P5[:, 1]), sc5h, 0.3)
    t_cpu5 = _best_of(cpu5)
    out.append({"workload": "config 5: 2000 objects: footprint association matrix (f64) + space NMS 0.1 + image NMS 0.3",
                "metric": "tracking-frame geometry frames/s", "value": 1e3 / t_graph, "unit": "frames/s", "n_gpus": 1,
                "ms": {"frame (one CUDA graph, FrameGeometry)": t_graph, "frame (separate drop-in calls)": t_frame},
                "separate_calls_frames_per_s": 1e3 / t_frame,
                "note": "replicas only (SURVEY.md §8e: the 2000 x 2000 matrix is too small to shard); launch/latency bound; graph "
                        "frame: 5 small input copies + one graph launch + one 8-byte read of the keep-list lengths",
                "cpu_baseline": {"value": 1.0 / t_cpu5, "unit": "frames/s", "cores": os.cpu_count(), "kind": "port",
                                 "sample": f"one frame (oracle association_cost + space_nms + state_to_im + im_nms; best of 3 "
                                           f"after a warm-up, {t_cpu5:.3f} s)"}})
    if world > 1:
        return out
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
    Xc0, Pc0, Dc, dtc = Xk[:n_cpu].cpu(), Pk[:n_cpu].cpu(), Dk[:n_cpu].cpu(), dtk[:n_cpu].cpu()
    zc = zk[:n_cpu].cpu()

    def cpuk():
        Xc, Pc = kf_oracle.predict(Xc0, Pc0, Dc, dtc, Fk, Qk)
        kf_oracle.update(Xc, Pc, torch.arange(n_cpu), zc, Hk, Rk, torch.zeros(Mk))
    t_cpuk = _best_of(cpuk)
    out.append({"workload": "SURVEY 8(f)-4: Torch_KF predict (per-object dt) + update, 1M objects, 6 states / 5 measurements",
                "metric": "Kalman predict+update M objects/s", "value": nk / ((t_pred + t_upd) * 1e-3) / 1e6,
                "unit": "M objects/s", "n_gpus": 1, "ms": {"predict": t_pred, "update": t_upd},
                "roofline": {"kernel": "kf_predict_kernel", "bound": "hbm", "achieved": pred_bytes / (t_pred * 1e-3) / 1e9,
                             "peak": hbm_peak, "unit": "GB/s", "frac": pred_bytes / (t_pred * 1e-3) / 1e9 / hbm_peak,
                             "update_frac": upd_bytes / (t_upd * 1e-3) / 1e9 / hbm_peak},
                "cpu_baseline": {"value": n_cpu / t_cpuk / 1e6, "unit": "M objects/s", "cores": os.cpu_count(), "kind": "port",
                                 "sample": f"100 k of the 1 M objects (oracle predict + update, torch CPU bmm / inverse; best of 3 "
                                           f"after a warm-up, {t_cpuk:.3f} s)"}})
    return out


def _loss_inputs(B, A, rank, dev, first_image_from_cpu):
    """synthetic batch of B images.  Image 0 of rank 0 is generated on the CPU exactly like the image of
    tests/test_gpu_losses.py::test_cfg2_one_image_1080p_200gt_vs_oracle (synth.gen(100)), so the GPU arm's first image can be
    cross-checked against the oracle before anything is timed; the rest comes from device generators."""
    import synth
    g = synth.gen(100 + rank)
    ann0 = synth.gt_annotations_3d(1, G_PER_IMG, H_IMG, W_IMG, g)
    cls0, reg0 = synth.head_outputs(1, A, C_CLS, R_REG, g)
    ann_rest = synth.gt_annotations_3d(B - 1, G_PER_IMG, H_IMG, W_IMG, g) if B > 1 else ann0[:0]
    ann_h = torch.cat((ann0, ann_rest)).contiguous().pin_memory()
    torch.manual_seed(100 + rank)
    cls_d = torch.rand(B, A, C_CLS, device=dev) * 0.1
    reg_d = torch.randn(B, A, R_REG, device=dev) * 0.1
    if first_image_from_cpu:
        cls_d[0].copy_(cls0[0])
        reg_d[0].copy_(reg0[0])
    return ann_h, cls_d, reg_d, (ann0, cls0, reg0)


def _time_loss_steps(step, steps, warmup, world, dev, allow_graph=True):
    """ms per step of `step()` (forward + backward): eager, and as one CUDA-graph replay per step.  Returns
    (best ms, mode, eager ms, graph ms or None, losses of the last step)."""
    for _ in range(warmup):
        step()
    torch.cuda.synchronize(dev)
    if world > 1:
        torch.distributed.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        losses = step()
    e1.record()
    torch.cuda.synchronize(dev)
    if world > 1:
        torch.distributed.barrier()
    losses = losses.detach().clone()
    ms_eager = _max_over_ranks(e0.elapsed_time(e1), world, dev) / steps
    ms_step, mode, ms_graph = ms_eager, "eager (one Python call per step)", None
    if allow_graph:
        graph, ok = None, 1
        try:
            side = torch.cuda.Stream(dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                for _ in range(3):
                    step()
            torch.cuda.current_stream(dev).wait_stream(side)
            torch.cuda.synchronize(dev)
            graph = torch.cuda.CUDAGraph()
            # thread_local: the NCCL watchdog thread's CUDA calls must not invalidate the capture (world > 1)
            with torch.cuda.graph(graph, capture_error_mode="thread_local"):
                g_losses = step().detach()
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
            for _ in range(warmup):
                graph.replay()
            torch.cuda.synchronize(dev)
            if world > 1:
                torch.distributed.barrier()
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            g0.record()
            for _ in range(steps):
                graph.replay()
            g1.record()
            torch.cuda.synchronize(dev)
            if world > 1:
                torch.distributed.barrier()
            ms_graph = _max_over_ranks(g0.elapsed_time(g1), world, dev) / steps
            same = torch.tensor([1 if torch.equal(g_losses, losses) else 0], dtype=torch.int32, device=dev)
            if world > 1:
                torch.distributed.all_reduce(same, op=torch.distributed.ReduceOp.MIN)
            if int(same.item()) and ms_graph < ms_step:
                ms_step, mode = ms_graph, "CUDA graph replay (one launch per step)"
        if graph is not None:
            graph.reset()
            del graph
    return ms_step, mode, ms_eager, ms_graph, losses


def run_ours(args):
    rank, world, local = _dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (use --impl reference for the CPU arm)")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    numa = _bind_to_gpu_numa_node(local)          # before any pinned allocation
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=dev)
    import synth
    from geom3d_b200 import dist as gdist
    from geom3d_b200 import losses_impl, ops

    hbm_peak, peak_src = _peaks()
    B = B_PER_GPU
    # the anchor table as the model gets it (retinanet/model.py:306: self.anchors(img_batch)): the drop-in Anchors module
    # writes it on the device and tags it as the regular pyramid, which lets the loss run its GT-centric assignment
    from geom3d_b200.anchors_impl import Anchors
    anc = Anchors()(torch.zeros(1, 3, H_IMG, W_IMG, device=dev))
    assert torch.equal(anc.cpu(), synth.anchors(H_IMG, W_IMG))
    A = anc.shape[1]
    ann_h, cls_raw, reg_raw, image0 = _loss_inputs(B, A, rank, dev, first_image_from_cpu=True)
    cls_d, reg_d = cls_raw.requires_grad_(True), reg_raw.requires_grad_(True)
    ann_d = ann_h.to(dev)
    ones = torch.ones(3, device=dev)

    # ---- parity gate: before anything is timed, image 0 of this arm against the oracle (the CPU arm's implementation)
    parity = None
    if rank == 0:
        from oracle import losses_oracle
        ann0, cls0, reg0 = image0
        t0 = time.perf_counter()
        ref = losses_oracle.focal_loss(cls0, reg0, anc.cpu(), ann0)
        t_oracle = time.perf_counter() - t0
        with torch.no_grad():
            per_image = losses_impl.focal_loss(cls_d.detach(), reg_d.detach(), anc, ann_d)[2].cpu()
        want = torch.tensor([float(ref[0]), float(ref[1]), float(ref[2]), float(ref[3][0][2].sum())])
        err = ((per_image[0].double() - want.double()).abs() / want.double().abs().clamp(min=1e-30)).tolist()
        assert max(err[:3]) < 1e-5 and err[3] == 0.0, f"GPU image 0 differs from the oracle: {per_image[0].tolist()} vs {want.tolist()}"
        parity = {"image0_gpu": per_image[0].tolist(), "image0_oracle": want.tolist(), "max_rel_err_losses": max(err[:3]),
                  "num_pos_equal": err[3] == 0.0, "oracle_seconds": t_oracle,
                  "note": "(cls, reg, vp, num_pos) of image 0 of the timed batch, checked before timing; same image as "
                          "tests/test_gpu_losses.py::test_cfg2_one_image_1080p_200gt_vs_oracle"}

    def make_step(c, r, a_d, sharded, trace_holder=None):
        def step():
            c.grad = None
            r.grad = None
            tr = trace_holder[0] if trace_holder else None
            if sharded:
                losses = gdist.sharded_focal_loss(c, r, anc, a_d, trace_events=tr)
            else:
                losses = losses_impl.focal_loss(c, r, anc, a_d, trace_events=tr)[0]
            losses.backward(ones)
            return losses
        return step

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # ---- headline: weak scaling, B images per rank
    step_device = make_step(cls_d, reg_d, ann_d, world > 1)
    ms_step, mode, ms_step_eager, ms_graph, losses = _time_loss_steps(step_device, args.steps, args.warmup, world, dev,
                                                                       allow_graph=not args.no_graph)
    # ---- per-kernel times: traced eager steps (the event records serialise the launches that otherwise overlap)
    kev = [[torch.cuda.Event(enable_timing=True) for _ in range(6)] for _ in range(args.steps)]
    bwd_ev = [[torch.cuda.Event(enable_timing=True) for _ in range(2)] for _ in range(args.steps)]
    for tr in kev + bwd_ev:
        for e in tr:
            e.record()      # creates the CUDA events outside the timed region
    torch.cuda.synchronize(dev)
    for i in range(args.steps):
        cls_d.grad = None
        reg_d.grad = None
        if world > 1:
            l3 = gdist.sharded_focal_loss(cls_d, reg_d, anc, ann_d, trace_events=kev[i])
        else:
            l3 = losses_impl.focal_loss(cls_d, reg_d, anc, ann_d, trace_events=kev[i])[0]
        bwd_ev[i][0].record()
        l3.backward(ones)
        bwd_ev[i][1].record()
    torch.cuda.synchronize(dev)
    del l3
    clocks = sampler.stop() if rank == 0 else None
    knames = ["loss_prologue_kernel", "assign_pairs_kernel", "assign_resolve_kernel", "focal_stream8_kernel", "positives_kernel"]
    ms_k = [_event_ms([(e[i], e[i + 1]) for e in kev]) / args.steps for i in range(5)]
    ms_bwd = _event_ms([(e[0], e[1]) for e in bwd_ev]) / args.steps
    ms_fwd = sum(ms_k)
    pairs_per_step = world * B * A * G_PER_IMG
    value = pairs_per_step / (ms_step * 1e-3) / 1e9
    loss_vals = [float(x) for x in losses.detach().cpu()]
    with torch.no_grad():
        per_image_vals = losses_impl.focal_loss(cls_d.detach(), reg_d.detach(), anc, ann_d)[2].cpu().tolist()

    # ---- strong scaling (BASELINE configs[1] as written: the 32-image batch sharded over the GPUs, 32 / N images each)
    strong = None
    if world > 1 and B_PER_GPU % world == 0:
        Bs = B_PER_GPU // world
        cs = cls_d.detach()[:Bs].clone().requires_grad_(True)
        rs = reg_d.detach()[:Bs].clone().requires_grad_(True)
        step_strong = make_step(cs, rs, ann_d[:Bs].contiguous(), True)
        ms_s, mode_s, ms_s_eager, _, _ = _time_loss_steps(step_strong, args.steps, args.warmup, world, dev, allow_graph=not args.no_graph)
        strong = {"global_batch": B_PER_GPU, "images_per_gpu": Bs, "ms_per_step": ms_s, "ms_per_step_eager": ms_s_eager,
                  "launch_mode": mode_s, "value": B_PER_GPU * A * G_PER_IMG / (ms_s * 1e-3) / 1e9, "unit": UNIT,
                  "note": "same global batch as N = 1: speed-up = ms_per_step(N = 1) / this; the fixed part of a step (six "
                          "launches of latency-bound work + the 5-scalar exchange) does not shrink with the shard"}
        del cs, rs

    # ---- end to end through the public module (default constructor), from pinned host buffers
    cls_h = torch.empty((B, A, C_CLS), dtype=torch.float32).pin_memory()
    reg_h = torch.empty((B, A, R_REG), dtype=torch.float32).pin_memory()
    cls_h.copy_(cls_d.detach())
    reg_h.copy_(reg_d.detach())
    module = losses_impl.FocalLoss()
    cls_in = torch.empty_like(cls_d).requires_grad_(True)
    reg_in = torch.empty_like(reg_d).requires_grad_(True)

    def copy_in():
        with torch.no_grad():
            cls_in.copy_(cls_h, non_blocking=True)
            reg_in.copy_(reg_h, non_blocking=True)
        return ann_h.to(dev, non_blocking=True)

    def step_e2e():
        cls_in.grad = None
        reg_in.grad = None
        ann_in = copy_in()
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
    # the roof of that number: the same host->device copies alone, all ranks at once
    torch.cuda.synchronize(dev)
    if world > 1:
        torch.distributed.barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        copy_in()
    torch.cuda.synchronize(dev)
    t_copy_local = (time.perf_counter() - t0) / e2e_steps
    t_copy = _max_over_ranks(t_copy_local, world, dev)
    h2d_rates = [h2d / t_copy_local / 1e9]
    if world > 1:
        gathered = [None] * world
        torch.distributed.all_gather_object(gathered, (h2d / t_copy_local / 1e9, numa))
        h2d_rates = gathered
    del cls_h, reg_h, cls_in, reg_in

    if rank == 0:
        # algorithmic bytes.  The sweep (focal_stream8_kernel) moves, per (image, anchor): 32 B of
        # classification in, 32 B of classification gradient out, 48 B of regression-gradient zeros out = 112 B - everything
        # SURVEY §8(d)'s fused forward+backward figure (160 B) contains except the 48 B read of `reg`, which this design
        # touches only on the ~1 % positive rows (positives_kernel).  The other launches move little and are latency-bound.
        stream_bytes = B * A * (C_CLS * 4 * 2 + R_REG * 4)
        chain_bytes = B * A * 4 * 2 + A * 16 + ann_h.numel() * 4               # key zero fill + keys of touched chunks, anchors, GT
        ms_stream = ms_k[3]
        achieved = stream_bytes / (ms_stream * 1e-3) / 1e9
        step_bytes = stream_bytes + chain_bytes
        per_kernel = [{"kernel": n, "ms": m, "share_of_traced_step": m / (ms_fwd + ms_bwd)} for n, m in zip(knames, ms_k)]
        per_kernel.append({"kernel": "backward: autograd dispatch on the host + focal_bwd_kernel (verifies the expected upstream gradients, exits)", "ms": ms_bwd,
                           "share_of_traced_step": ms_bwd / (ms_fwd + ms_bwd)})
        per_kernel[3].update({"bytes": stream_bytes, "GBps": achieved, "frac": achieved / hbm_peak, "limiter": "hbm",
                              "traffic": _traffic("focal_stream8_kernel")})
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": "training-loss path (BASELINE configs[1]): FocalLoss fwd+bwd, 1080p, A=389205, "
                                   "G=200 GT/img, C=8, 12-d regression", "batch_per_gpu": B, "global_batch": B * world,
                       "parallelism": f"images sharded over {world} GPU(s); only exchange: 5 scalars per rank, stored into the peers' memory over NVLink",
                       "l2": "inputs (1.0 GB/step) exceed the 126 MB L2; no flush needed",
                       "launch_mode": mode, "ms_per_step_eager": ms_step_eager, "ms_per_step_graph": ms_graph},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 12,
                    "ms_per_step": t_e2e * 1e3, "steps": e2e_steps,
                    "module": "losses.FocalLoss() as shipped (default constructor: lazy empty-batch check, no host sync)",
                    "h2d_only_ms": t_copy * 1e3, "h2d_roof_fraction": t_copy / t_e2e,
                    "h2d_GBps_per_rank": h2d_rates if world == 1 else [x[0] for x in h2d_rates],
                    "numa": numa if world == 1 else [x[1] for x in h2d_rates],
                    "note": "the step is bound by the host->device copy of its 1 GB of inputs: h2d_roof_fraction = the share "
                            "of the step that the same copies take alone; each rank's pinned buffers are allocated on the "
                            "NUMA node of its GPU"},
            # prologue, pairs, resolve, stream, positives + the backward's verification launch
            "gpu_launches": 6 * args.steps,
            "roofline": {"kernel": "focal_stream8_kernel", "bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                         "frac": achieved / hbm_peak, "traffic": _traffic("focal_stream8_kernel"), "peak_source": peak_src,
                         "bytes_per_launch": stream_bytes, "ms_per_launch": ms_stream,
                         "kernels": per_kernel,
                         # the whole step on the bytes it has to move (cls in, dcls + dreg out, key fill + touched keys) ...
                         "step_on_bytes_moved": {"bytes": step_bytes, "GBps": step_bytes / (ms_step * 1e-3) / 1e9,
                                                 "frac": step_bytes / (ms_step * 1e-3) / 1e9 / hbm_peak},
                         # ... and on SURVEY.md §8(d)'s 160 B per (image, anchor), 48 B of which (reading reg) this design never moves
                         "step_on_survey_bytes": {"bytes": B * A * 160 + A * 16,
                                                  "GBps": (B * A * 160 + A * 16) / (ms_step * 1e-3) / 1e9,
                                                  "frac": (B * A * 160 + A * 16) / (ms_step * 1e-3) / 1e9 / hbm_peak}},
            "parity": parity,
            "losses": loss_vals, "e2e_losses": [float(x) for x in host_losses],
            "positives_per_image_mean": sum(p[3] for p in per_image_vals) / len(per_image_vals),
        }
        if strong is not None:
            line["strong_scaling"] = strong

    # forward-only pass (validation loss, no gradient buffers): GT-centric assignment on the tagged pyramid table against
    # the anchor-centric kernel on an untagged copy of the same table (same codes, same losses)
    if rank == 0 and world == 1:
        anc_plain = anc.clone()

        def fwd_time(table):
            with torch.no_grad():
                t = _timed(lambda: ops.focal_loss_forward(cls_d, reg_d, table, ann_d, want_assign=False), 10, dev)
                return t, ops.focal_loss_forward(cls_d, reg_d, table, ann_d)
        t_gt, o_gt = fwd_time(anc)
        t_an, o_an = fwd_time(anc_plain)
        assert o_gt["gt_centric"] and not o_an["gt_centric"] and torch.equal(o_gt["assign"], o_an["assign"])
        line["forward_only"] = {"ms_gt_centric": t_gt, "ms_anchor_centric": t_an,
                                "G_pairs_per_s_gt_centric": pairs_per_step / (t_gt * 1e-3) / 1e9}
        # the training step on an untagged table (any anchor tensor): anchor-centric assignment
        cp, rp = cls_d.detach().clone().requires_grad_(True), reg_d.detach().clone().requires_grad_(True)

        def step_plain():
            cp.grad = None
            rp.grad = None
            l3 = losses_impl.focal_loss(cp, rp, anc_plain, ann_d)[0]
            l3.backward(ones)
            return l3
        line["anchor_centric_step_ms"] = _timed(step_plain, 5, dev)
        del anc_plain, o_gt, o_an, cp, rp
        # opt-in FocalLoss(persistent_grad=True): gradient buffers and workspace kept from step to step, so the 48 B/anchor
        # of zeros of the regression gradient are not written again (only the previous step's positive rows are cleared).
        # Not the headline: it changes who owns the gradient tensors (see losses_impl.FocalLoss).  Same kernels as the
        # module (forward-with-gradients + the backward's verification launch), bit-identical gradients (checked here).
        pg = ops.PersistentGrads()
        cdet, rdet = cls_d.detach(), reg_d.detach()

        def step_keep():
            f = ops.focal_loss_forward(cdet, rdet, anc, ann_d, want_assign=False, grad_expected=1.0, persistent=pg)
            ops.focal_loss_backward(f, ones)
            return f["losses"]
        ms_keep, mode_keep, ms_keep_eager, _, _ = _time_loss_steps(step_keep, args.steps, args.warmup, 1, dev,
                                                                   allow_graph=not args.no_graph)
        fresh = ops.focal_loss_forward(cdet, rdet, anc, ann_d, want_assign=False, grad_expected=1.0)
        ops.focal_loss_backward(fresh, ones)
        same = bool(torch.equal(pg.bufs[0], fresh["dcls"]) and torch.equal(pg.bufs[1], fresh["dreg"]))
        assert same, "persistent gradient buffers differ from freshly written ones"
        keep_bytes = B * A * 64 + A * 16
        line["persistent_grad"] = {"ms_per_step": ms_keep, "ms_per_step_eager": ms_keep_eager, "launch_mode": mode_keep,
                                   "G_pairs_per_s": pairs_per_step / (ms_keep * 1e-3) / 1e9,
                                   "gradients_equal_default_path": same,
                                   "step_bytes": keep_bytes, "step_frac_of_hbm_peak": keep_bytes / (ms_keep * 1e-3) / 1e9 / hbm_peak,
                                   "note": "FocalLoss(persistent_grad=True), opt-in: cls in + dcls out = 64 B per (image, anchor)"}
        del pg, fresh, cdet, rdet
    del cls_d, reg_d
    torch.cuda.empty_cache()
    extras = None
    if not args.no_extras:
        try:
            extras = other_workloads(dev, hbm_peak, rank, world)
        except Exception as e:  # the headline line must still be printed
            extras = {"error": f"{type(e).__name__}: {e}"}
            if world > 1:
                raise
    if rank == 0:
        if world == 1:
            cpu_value, cpu_s = cpu_loss_sample(8)
            line["cpu_baseline"] = {"value": cpu_value, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                                    "sample": f"8 of the 32 images (oracle port, forward+backward, best of 2 after 1 warm-up, {cpu_s:.2f} s/run)"}
        if extras is not None:
            line["other_workloads"] = extras
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
