#!/usr/bin/env python
"""bench.py -- tracked features/sec of the pyramidal-LK hot path (BASELINE.json metric) on N B200s of one node.

Workload ("step" = one pass of the hot path over one batch): config C3 of BASELINE.json -- synthetic 1241x376
frame pairs, 8192 corners per pair, 4-level 21x21 pyramidal LK (30 iterations, eps 0.01), `--pairs` pairs PER GPU
(default 4096; weak scaling: every rank tracks its own contiguous block of pairs, no collective on the data path).
The pairs are `--base-pairs` distinct synthetic pairs (tools/synth.py recipe) tiled to `--pairs`
separate device buffers, so the per-step working set (~19 GB incl. pyramids) is far larger than L2.

  value     device-resident: inputs already in HBM; per step = pyramids of both frames + Scharr + LK + outputs in HBM.
  e2e       the same work through the host-buffer C-ABI entry point dr3lk_track_batch_host (pinned HOST images and
            points in, results back in host memory; H2D/D2H inside the timed region).
  roofline  LK kernel alone: algorithmic bytes (from the kernel's own per-feature iteration counts)
            / its CUDA-event duration, against MEASURED_PEAKS.json hbm_gbs.
  cpu_baseline  the reference's CPU path (cv2.calcOpticalFlowPyrLK, all host threads) on a bounded sample.

`--impl reference` times only that CPU path (rank 0) and prints the same JSON shape.
"""
import argparse
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "tracked features/sec at 1/2/4/8 B200 (1241x376, 4-level 21x21 pyramidal LK)"
UNIT = "features/s"
WIN, MAX_LEVEL, CRIT, FLAGS = (21, 21), 3, (3, 30, 0.01), 0
W_IMG, H_IMG, CORNERS = 1241, 376, 8192

# The headline workload is C3 (BASELINE.json configs[2], the one `metric` is quoted on).  --workload runs the other named
# shapes through exactly the same code (informational; the driver only runs the default):
#   name: (description, W, H, window, maxLevel, default pairs per GPU, default distinct pairs)
WORKLOADS = {
    "c3": ("C3: synthetic 1241x376 frame pairs x 8192 corners, 4-level 21x21 LK (30 it, eps 0.01)", 1241, 376, (21, 21), 3, 4096, 32),
    "kitti": ("KITTI: the 9 bundled consecutive pairs data/kitti0..9 (1240x376) tiled, cv2 FAST corners of each previous frame, "
              "4-level 21x21 LK (30 it, eps 0.01)", 1240, 376, (21, 21), 3, 2304, 9),
    "c4": ("C4: synthetic 3840x2160 frame pairs x 50000 corners, 5-level 31x31 LK (30 it, eps 0.01)", 3840, 2160, (31, 31), 4, 64, 2),
    "c5": ("C5: semi-dense 4-px lattice (29140 points) on synthetic 1241x376 frame pairs, 4-level 21x21 LK (30 it, eps 0.01)",
           1241, 376, (21, 21), 3, 1024, 16),
}


def make_workload(name, nb, seed0):
    """nb distinct frame pairs of the workload: prev (nb,H,W) u8, next, pts (N,2) f32, offsets (nb+1,) i32."""
    from tools import synth
    if name == "c3":
        return synth.make_batch(nb, W_IMG, H_IMG, CORNERS, seed0=seed0)
    if name == "c5":
        return synth.make_batch(nb, W_IMG, H_IMG, CORNERS, seed0=seed0, points="lattice")
    if name == "c4":
        return synth.make_batch(nb, W_IMG, H_IMG, 50000, seed0=1000 + seed0, min_dist=5)
    if name == "kitti":
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from _common import load_gray
        from oracle import cv2_ref
        frames = [load_gray("kitti%d.png" % i) for i in range(10)]
        sel = [i % 9 for i in range(nb)]
        pts = [cv2_ref.fast_corners(frames[i])[0] for i in sel]
        offs = np.concatenate([[0], np.cumsum([len(p) for p in pts])]).astype(np.int32)
        return (np.stack([frames[i] for i in sel]), np.stack([frames[i + 1] for i in sel]), np.concatenate(pts).astype(np.float32), offs)
    raise SystemExit("unknown workload " + name)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--pairs", type=int, default=0, help="frame pairs per GPU per step (0: the workload's default, 4096 for c3)")
    ap.add_argument("--base-pairs", type=int, default=0, help="distinct pairs generated per rank (0: the workload's default, 32 for c3)")
    ap.add_argument("--cpu-sample-pairs", type=int, default=0, help="pairs per reference-arm step; the cpu_baseline leg uses 3x (0: 512 for c3)")
    ap.add_argument("--single-process", action="store_true", help="drive all --gpus devices from ONE process through dr3lk_multi (host-buffer "
                    "path only; the driver's contract arm is one process per GPU under torchrun)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    return ap.parse_args()


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                pass
        sm, mx, reasons = [], 0.0, set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx = max(mx, float(r[2]))
                for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        # "under load": samples in the upper half of what was seen (the sampler also sees idle gaps)
        load = sorted(sm)[len(sm) // 2:] if sm else []
        return {"sm_mhz": float(np.median(load)) if load else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


def cpu_reference(prev, nxt, pts, offs, n_pairs, threads=None):
    """The reference's CPU path on `n_pairs` pairs: cv2.calcOpticalFlowPyrLK per pair (pyramids included, as the
    reference calls it with level-0 images only), all host threads.  Falls back to the C oracle port (OpenMP)."""
    from oracle import cv2_ref
    import oracle
    cores = os.cpu_count() or 1
    n_pairs = min(n_pairs, len(offs) - 1)
    tracked = feats = 0
    if cv2_ref.HAVE_CV2:
        cv2_ref.cv2.setNumThreads(threads or cores)
        kind, used = "reference", cv2_ref.cv2.getNumThreads()
        t0 = time.perf_counter()
        for b in range(n_pairs):
            p = pts[offs[b]:offs[b + 1]]
            _, st, _ = cv2_ref.calc_optical_flow_pyr_lk(prev[b], nxt[b], p, None, WIN, MAX_LEVEL, CRIT, FLAGS)
            tracked += int(st.sum()); feats += len(p)
        dt = time.perf_counter() - t0
        what = "cv2 %s calcOpticalFlowPyrLK" % cv2_ref.CV2_VERSION
    else:
        kind, used = "port", oracle.num_threads()
        t0 = time.perf_counter()
        for b in range(n_pairs):
            p = pts[offs[b]:offs[b + 1]]
            _, st, _ = oracle.calc_optical_flow_pyr_lk(prev[b], nxt[b], p, None, WIN, MAX_LEVEL, CRIT, FLAGS)
            tracked += int(st.sum()); feats += len(p)
        dt = time.perf_counter() - t0
        what = "oracle/lk_oracle.c (OpenMP)"
    return {"tracked": tracked, "features": feats, "seconds": dt, "kind": kind, "cores": used, "what": what, "pairs": n_pairs}


def measure_latency(dr3, ctx):
    """Per-call latency of the reference's own call pattern on the bundled KITTI frames (host buffers in, host results out,
    synchronous): C1 = one calcOpticalFlowPyrLK call kitti0 -> kitti1 with the reference detector's corners (<= 546) and with
    4607 cv2 FAST corners; C2 = the chain kitti0..9 through the streaming entry point (one call per new frame)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from _common import load_gray
    from oracle import cv2_ref
    frames = [load_gray("kitti%d.png" % i) for i in range(10)]
    xy, _, _ = ctx.fast_detect(frames[0])
    few = xy.astype(np.float32)
    many = cv2_ref.fast_corners(frames[0])[0] if cv2_ref.HAVE_CV2 else np.tile(few, (8, 1))

    def med_us(fn, n=200, warm=20):
        for _ in range(warm):
            fn()
        ts = []
        for _ in range(n):
            t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
        return 1e6 * float(np.median(ts))

    out = {"c1_call_us_%d_points" % len(few): med_us(lambda: ctx.calc_optical_flow_pyr_lk(frames[0], frames[1], few)),
           "c1_call_us_%d_points" % len(many): med_us(lambda: ctx.calc_optical_flow_pyr_lk(frames[0], frames[1], many)),
           "c1_reference_params_30x30_us_%d_points" % len(few): med_us(
               lambda: ctx.calc_optical_flow_pyr_lk(frames[0], frames[1], few, few, (30, 30), 4, (3, 1000, 1e-3), dr3.USE_INITIAL_FLOW))}

    # the same frames in page-locked memory (dr3lk_host_alloc) with rows at the device pitch (width rounded up to 16): the
    # library hands them to the copy engine without staging
    hh, ww = frames[0].shape
    pins = [dr3.PinnedArray((hh, (ww + 15) // 16 * 16), np.uint8) for _ in frames]
    for pa, f in zip(pins, frames):
        pa.array[:, :ww] = f
    pf = [pa.array[:, :ww] for pa in pins]
    out["c1_call_us_%d_points_pinned_images" % len(few)] = med_us(lambda: ctx.calc_optical_flow_pyr_lk(pf[0], pf[1], few))

    def chain(fr=frames):
        pyr = dr3.Pyramid(ctx, fr[0])
        cur = many
        for i in range(1, 10):
            p, s, _, nxt = ctx.track_frame(pyr, fr[i], cur, keep_next=2)
            pyr.close(); pyr = nxt
            cur = p[s == 1]
        pyr.close()
        return len(cur)
    out["c2_chain_kitti0_9_ms"] = med_us(chain, n=30, warm=5) / 1e3
    out["c2_chain_kitti0_9_ms_pinned_images"] = med_us(lambda: chain(pf), n=30, warm=5) / 1e3
    out["c2_survivors"] = int(chain())
    assert int(chain(pf)) == out["c2_survivors"]
    for pa in pins:
        pa.free()
    out["what"] = "median wall-clock of the synchronous host-buffer calls (H2D + kernels + D2H + sync inside)"
    # the same call from compiled C++ through the header shim (3dr_b200/host/call_latency.cpp): no Python binding in the way
    exe = os.path.join(ROOT, "3dr_b200", "host", "call_latency")
    if os.path.exists(exe):
        import subprocess, tempfile
        with tempfile.TemporaryDirectory() as td:
            for name, f in (("a.pgm", frames[0]), ("b.pgm", frames[1])):
                with open(os.path.join(td, name), "wb") as fh:
                    fh.write(b"P5\n%d %d\n255\n" % (f.shape[1], f.shape[0]))
                    fh.write(np.ascontiguousarray(f).tobytes())
            np.savetxt(os.path.join(td, "pts.txt"), few, fmt="%.9g")
            r = subprocess.run([exe, os.path.join(td, "a.pgm"), os.path.join(td, "b.pgm"), os.path.join(td, "pts.txt"), "300"],
                               capture_output=True, text=True)
            out["compiled_cpp_caller"] = json.loads(r.stdout) if r.returncode == 0 else {"failed": (r.stderr or r.stdout)[-300:]}
            # C2 from C++: the nine-frame chain with the same starting points as c2_chain_kitti0_9_ms above
            more = []
            for i in range(2, 10):
                more.append(os.path.join(td, "f%d.pgm" % i))
                with open(more[-1], "wb") as fh:
                    fh.write(b"P5\n%d %d\n255\n" % (frames[i].shape[1], frames[i].shape[0]))
                    fh.write(np.ascontiguousarray(frames[i]).tobytes())
            np.savetxt(os.path.join(td, "many.txt"), many, fmt="%.9g")
            r = subprocess.run([exe, os.path.join(td, "a.pgm"), os.path.join(td, "b.pgm"), os.path.join(td, "many.txt"), "300"] + more,
                               capture_output=True, text=True)
            out["compiled_cpp_chain"] = json.loads(r.stdout) if r.returncode == 0 else {"failed": (r.stderr or r.stdout)[-300:]}
    return out


_REAL_STDOUT = None


def emit(obj):
    """The ONE JSON line of the contract, written to the process's original stdout."""
    line = json.dumps(obj) + "\n"
    if _REAL_STDOUT is not None:
        os.write(_REAL_STDOUT, line.encode())
    else:
        sys.stdout.write(line)
        sys.stdout.flush()


def main():
    global _REAL_STDOUT, WIN, MAX_LEVEL, W_IMG, H_IMG, CORNERS
    args = parse()
    desc, W_IMG, H_IMG, WIN, MAX_LEVEL, d_pairs, d_base = WORKLOADS[args.workload]
    args.pairs = args.pairs or d_pairs
    args.base_pairs = args.base_pairs or d_base
    data_kind = "bundled KITTI frames (data/kitti0..9), tiled" if args.workload == "kitti" else "synthetic"
    args.cpu_sample_pairs = args.cpu_sample_pairs or {"c3": 512, "kitti": 512, "c5": 128, "c4": 16}[args.workload]
    # libraries (NCCL version banner, torchrun notices) print to fd 1: keep stdout clean for the JSON line
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    # the SAME dict in both arms (the reference arm times a bounded sample of this workload, stated in its cpu_baseline.sample)
    cfg = {"workload": desc,
           "pairs_per_gpu": args.pairs, "corners_per_pair": CORNERS, "distinct_pairs_per_gpu": args.base_pairs,
           "sharding": "independent frame pairs per GPU, no collective", "win": list(WIN), "max_level": MAX_LEVEL,
           "l2": "inputs (%.1f GB/GPU/step) larger than L2" % ((2 * args.pairs * W_IMG * H_IMG) / 1e9)}
    extra = {}  # arm-specific notes go next to `config`, not into it

    from tools import synth

    # One process per GPU: run it (and allocate its pinned host buffers, first touch) on the CPUs next to that GPU, so the
    # eight ranks of a box do not pull their H2D traffic across the socket interconnect.  Our arm only.
    if args.impl == "ours" and world > 1 and not os.environ.get("DR3LK_NO_AFFINITY"):
        try:
            import pynvml
            pynvml.nvmlInit()
            words = (os.cpu_count() + 63) // 64
            mask = pynvml.nvmlDeviceGetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local_rank), words)
            cpus = {64 * i + b for i, m in enumerate(mask) for b in range(64) if (m >> b) & 1} & os.sched_getaffinity(0)
            if cpus:
                os.sched_setaffinity(0, cpus)
                extra["cpu_affinity"] = "GPU-local CPUs (%d of %d)" % (len(cpus), os.cpu_count())
        except Exception as e:  # no NVML / no permission: run unpinned
            extra["cpu_affinity"] = "unpinned (%s)" % type(e).__name__

    # ---------------------------------------------------------------- reference arm: CPU only, rank 0 only
    if args.impl == "reference":
        if rank != 0:
            return
        nb = max(1, min(args.base_pairs, args.cpu_sample_pairs))
        prev, nxt, pts, offs = make_workload(args.workload, nb, 1000)
        CORNERS = int(round(len(pts) / nb))
        cfg["corners_per_pair"] = CORNERS
        idx = np.arange(args.cpu_sample_pairs) % nb
        P, N = prev[idx], nxt[idx]
        lens = np.diff(offs)[idx]
        o2 = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
        p2 = np.concatenate([pts[offs[i]:offs[i + 1]] for i in idx])
        for _ in range(args.warmup):
            cpu_reference(P, N, p2, o2, min(8, args.cpu_sample_pairs))
        tot_t = tot_tr = tot_f = 0
        r = None
        for _ in range(args.steps):
            r = cpu_reference(P, N, p2, o2, args.cpu_sample_pairs)
            tot_t += r["seconds"]; tot_tr += r["tracked"]; tot_f += r["features"]
        val = tot_tr / tot_t
        sample = "%d pairs x %d corners per step (bounded sample of the %d-pair workload), %s" % (args.cpu_sample_pairs, CORNERS, args.pairs, r["what"])
        emit(({"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
                          "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / args.steps, "higher_is_better": True, "scaling": "weak",
                          "vs_baseline": None, "dtype": "int32 fixed-point + fp32 solve", "data": data_kind, "config": cfg,
                          "cpu_baseline": {"value": val, "unit": UNIT, "cores": r["cores"], "kind": r["kind"], "sample": sample},
                          "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                          "submitted_features_per_s": tot_f / tot_t, "cpu_sample_pairs_per_step": args.cpu_sample_pairs,
                          "host_cpus": os.cpu_count()}))
        return

    # ---------------------------------------------------------------- our arm
    import torch
    import torch.distributed as dist
    dr3 = importlib.import_module("3dr_b200")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
    if args.single_process:
        # ONE process, all --gpus devices through the C ABI's multi-GPU entry point (dr3lk_multi_track_batch_host): one worker
        # thread + context per device, contiguous blocks of pairs, host buffers in / out.  Not the driver's contract arm.
        G = args.gpus
        nb = max(1, min(args.base_pairs, args.pairs))
        prev_b, next_b, pts_b, offs_b = make_workload(args.workload, nb, 1000)
        total = args.pairs * G
        idx = np.arange(total) % nb
        offs = np.concatenate([[0], np.cumsum(np.diff(offs_b)[idx])]).astype(np.int32)
        n_feat = int(offs[-1])
        hp = dr3.PinnedArray((total, H_IMG, W_IMG), np.uint8)
        hn = dr3.PinnedArray((total, H_IMG, W_IMG), np.uint8)
        hpts = dr3.PinnedArray((n_feat, 2), np.float32)
        o_np, o_st, o_err = dr3.PinnedArray((n_feat, 2), np.float32), dr3.PinnedArray((n_feat,), np.uint8), dr3.PinnedArray((n_feat,), np.float32)
        for i in range(nb):
            sel = np.where(idx == i)[0]
            hp.array[sel] = prev_b[i]; hn.array[sel] = next_b[i]
        hpts.array[...] = np.concatenate([pts_b[offs_b[i]:offs_b[i + 1]] for i in idx])
        with dr3.MultiContext(list(range(G))) as mc:
            def step():
                mc.track_batch_host(hp.array, hn.array, hpts.array, offs, None, WIN, MAX_LEVEL, CRIT, FLAGS, out=(o_np.array, o_st.array, o_err.array, None))
            for _ in range(max(1, args.warmup)):
                step()
            l0 = mc.launch_count
            t0 = time.perf_counter()
            for _ in range(args.steps):
                step()
            wall = time.perf_counter() - t0
            launches = mc.launch_count - l0
        tracked = int(o_st.array.sum())
        # every copy of a distinct pair must come out identical whichever device tracked it
        first = {}
        ok = True
        for b in range(total):
            r = first.setdefault(int(idx[b]), b)
            if r != b:
                ok = ok and bool(np.array_equal(o_np.array[offs[b]:offs[b + 1]].view(np.uint32), o_np.array[offs[r]:offs[r + 1]].view(np.uint32))
                                 and np.array_equal(o_st.array[offs[b]:offs[b + 1]], o_st.array[offs[r]:offs[r + 1]]))
        val = tracked * args.steps / wall
        emit({"metric": METRIC, "value": val, "unit": UNIT, "n_gpus": G, "steps": args.steps, "warmup": args.warmup,
              "ms_per_step": 1e3 * wall / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
              "dtype": "int32 fixed-point windows + fp32 2x2 solve", "data": data_kind, "config": cfg, "mode": "single-process (dr3lk_multi)",
              "gpu_launches": int(launches), "replicas_consistent": ok,
              "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": int(2 * total * H_IMG * W_IMG + 8 * n_feat + 4 * (total + 1)),
                      "d2h_bytes_per_step": int(13 * n_feat), "api": "dr3lk_multi_track_batch_host (pinned host buffers, one process)",
                      "h2d_GBps_per_gpu": (2 * total * H_IMG * W_IMG + 8 * n_feat) / G / (wall / args.steps) / 1e9},
              "note": "value == e2e in this mode: host buffers in, host results out, wall clock around the synchronous calls"})
        return
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)

    nb = max(1, min(args.base_pairs, args.pairs))
    prev_b, next_b, pts_b, offs_b = make_workload(args.workload, nb, 1000 + rank * nb)
    idx = np.arange(args.pairs) % nb
    lens = np.diff(offs_b)[idx]
    offs = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    pts = np.concatenate([pts_b[offs_b[i]:offs_b[i + 1]] for i in idx]).astype(np.float32)
    n_feat = int(offs[-1])
    CORNERS = int(round(n_feat / args.pairs))
    cfg["corners_per_pair"] = CORNERS
    t_idx = torch.from_numpy(idx).to(dev)
    prev_d = torch.from_numpy(prev_b).to(dev)[t_idx].contiguous()   # (pairs, H, W) distinct device buffers
    next_d = torch.from_numpy(next_b).to(dev)[t_idx].contiguous()
    pts_d = torch.from_numpy(pts).to(dev)
    nxt_d = torch.zeros_like(pts_d)
    st_d = torch.zeros(n_feat, dtype=torch.uint8, device=dev)
    err_d = torch.zeros(n_feat, dtype=torch.float32, device=dev)
    stats_d = torch.zeros(n_feat, dtype=torch.int32, device=dev)

    ctx = dr3.Context(local_rank)
    # a dedicated (non-default) stream shared by torch and the library, so the CUDA events below see the kernels
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)

    def step_device(stats=False):
        ctx.track_batch(prev_d.data_ptr(), next_d.data_ptr(), W_IMG, H_IMG, W_IMG, W_IMG * H_IMG, args.pairs, pts_d.data_ptr(),
                        nxt_d.data_ptr(), st_d.data_ptr(), err_d.data_ptr(), offs, stats_d.data_ptr() if stats else None,
                        WIN, MAX_LEVEL, CRIT, FLAGS)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # ---- device-resident timing (value) + LK-kernel roofline
    for _ in range(max(args.warmup, 1)):
        step_device(stats=True)
    torch.cuda.synchronize()
    tracked_per_step = int(st_d.sum().item())

    def replicas_consistent():
        """Size-independent self-check at full scale: the batch tiles `nb` distinct pairs, so every copy of a pair must
        produce bit-identical positions, status, err and iteration counts wherever it sits in the batch."""
        first = {}
        for b in range(args.pairs):
            first.setdefault(int(idx[b]), b)
        ok = True
        for b in range(args.pairs):
            r = first[int(idx[b])]
            if r == b:
                continue
            a0, a1, r0, r1 = int(offs[b]), int(offs[b + 1]), int(offs[r]), int(offs[r + 1])
            ok = ok and bool(torch.equal(nxt_d[a0:a1].view(torch.int32), nxt_d[r0:r1].view(torch.int32)) and torch.equal(st_d[a0:a1], st_d[r0:r1])
                             and torch.equal(err_d[a0:a1].view(torch.int32), err_d[r0:r1].view(torch.int32)) and torch.equal(stats_d[a0:a1], stats_d[r0:r1]))
        return ok

    replicas_ok = replicas_consistent()
    stats_h = stats_d.cpu().numpy().view(np.uint32)
    alg_bytes = dr3.algorithmic_bytes(stats_h, WIN)
    iters_per_feat = float(dr3.decode_stats(stats_h)[0].mean())
    ctx.profile_read()
    ctx.set_profiling(True)
    launches0 = ctx.launch_count
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        step_device()
    e1.record(stream)
    barrier()
    clocks = sampler.stop()
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    launches = ctx.launch_count - launches0
    lk_ms, lk_n, pyr_ms, _ = ctx.profile_read()
    ctx.set_profiling(False)
    tracked_all = sum_over_ranks(tracked_per_step)
    feats_all = sum_over_ranks(n_feat)
    value = tracked_all * args.steps / (ms_total * 1e-3)
    hbm_peak, peak_src = peaks()
    lk_ms_avg = lk_ms / max(lk_n, 1)
    achieved = alg_bytes / (lk_ms_avg * 1e-3) / 1e9
    # DRAM traffic of the LK launch: bytes/feature from the committed ncu --set full capture, scaled to this launch
    traffic, traffic_src = None, None
    try:
        if args.workload != "c3":
            raise KeyError("the committed capture is of the C3 kernel")
        tfile = "lk_traffic_r02.json" if os.path.exists(os.path.join(ROOT, "profiles", "lk_traffic_r02.json")) else "lk_traffic_r01.json"
        with open(os.path.join(ROOT, "profiles", tfile)) as f:
            tj = json.load(f)
        traffic = tj["dram_bytes_per_feature"] * n_feat
        traffic_src = "dram__bytes_read+write.sum per feature from %s, x %d features" % (tj["source"], n_feat)
    except Exception:
        pass
    # What the number means: `achieved` / `frac` are the north_star metric -- ALGORITHMIC bytes per second of the LK kernel
    # against the measured HBM peak.  The kernel's working set is L2-resident (DRAM traffic ~3 % of the algorithmic bytes),
    # so this is not HBM utilisation; the limiter is instruction issue / the integer multiply pipe (`limiter`, from the
    # committed ncu capture).
    lim = {}
    try:
        with open(os.path.join(ROOT, "profiles", "lk_fast_r02_limiter.json")) as f:
            lim = json.load(f)
    except Exception:
        pass
    roof = {"bound": "issue", "metric_denominator": "hbm", "kernel": "lk_fast_kernel (LK kernel alone, all levels in one launch)",
            "achieved": achieved, "peak": hbm_peak,
            "unit": "GB/s", "frac": achieved / hbm_peak, "effective_GBps": achieved, "traffic": traffic, "traffic_source": traffic_src,
            "dram_frac_of_peak": (traffic / (lk_ms_avg * 1e-3) / 1e9 / hbm_peak) if traffic else None, "peak_source": peak_src,
            "limiter": lim or None,
            "algorithmic_bytes_per_launch": alg_bytes, "lk_ms_per_launch": lk_ms_avg, "pyramid_ms_per_step": pyr_ms / max(lk_n, 1),
            "lk_iterations_per_feature": iters_per_feat,
            "note": "algorithmic bytes = sum over features of (5T*levels_with_template + T*iterations + T*err_pass + 21), T=(%d+1)^2 "
                    "(SURVEY.md 8d); the kernel's working set is L2-resident so DRAM traffic is far below this" % WIN[0]}

    # ---- end to end through the host-buffer C ABI
    e2e = None
    if not args.no_e2e:
        hp = dr3.PinnedArray((args.pairs, H_IMG, W_IMG), np.uint8)
        hn = dr3.PinnedArray((args.pairs, H_IMG, W_IMG), np.uint8)
        hpts = dr3.PinnedArray((n_feat, 2), np.float32)
        o_np, o_st, o_err = dr3.PinnedArray((n_feat, 2), np.float32), dr3.PinnedArray((n_feat,), np.uint8), dr3.PinnedArray((n_feat,), np.float32)
        for i in range(nb):
            sel = np.where(idx == i)[0]
            hp.array[sel] = prev_b[i]; hn.array[sel] = next_b[i]
        hpts.array[...] = pts

        def step_host():
            return ctx.track_batch_host(hp.array, hn.array, hpts.array, offs, None, WIN, MAX_LEVEL, CRIT, FLAGS,
                                        out=(o_np.array, o_st.array, o_err.array, None))
        for _ in range(max(1, min(args.warmup, 2))):
            step_host()
        barrier()
        t0 = time.perf_counter()
        e0.record(stream)
        for _ in range(args.steps):
            step_host()
        e1.record(stream)
        barrier()
        wall = time.perf_counter() - t0
        ms_e2e = max_over_ranks(max(e0.elapsed_time(e1), 0.0))
        wall = max_over_ranks(wall)
        tracked_host = sum_over_ranks(int(o_st.array.sum()))
        # parity of the two entry points on the same inputs
        same = bool(np.array_equal(o_st.array, st_d.cpu().numpy()) and np.array_equal(o_np.array.view(np.uint32), nxt_d.cpu().numpy().view(np.uint32)))
        e2e = {"value": tracked_host * args.steps / wall, "unit": UNIT,
               "h2d_bytes_per_step": int(2 * args.pairs * H_IMG * W_IMG + 8 * n_feat + 4 * (args.pairs + 1)),
               "d2h_bytes_per_step": int(13 * n_feat), "ms_per_step": 1e3 * wall / args.steps, "timer": "host wall clock around the synchronous call, max over ranks",
               "cuda_event_ms_per_step": ms_e2e / args.steps, "api": "dr3lk_track_batch_host (pinned host buffers)",
               "matches_device_path": same,
               "h2d_GBps_per_gpu": (2 * args.pairs * H_IMG * W_IMG + 8 * n_feat) / (wall / args.steps) / 1e9}
        for a in (hp, hn, hpts, o_np, o_st, o_err):
            a.free()

    # ---- single-call latency (the reference's real workload: one frame pair, a few hundred points per call)
    latency = None
    if args.workload == "kitti" and rank == 0 and world == 1:
        latency = measure_latency(dr3, ctx)

    # ---- CPU baseline on the host cores (rank 0, N=1 only)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        n_s = 3 * args.cpu_sample_pairs  # ~10-30 s of CPU work on the box's cores
        ci = np.arange(n_s) % nb
        c_off = np.concatenate([[0], np.cumsum(np.diff(offs_b)[ci])]).astype(np.int32)
        c_pts = np.concatenate([pts_b[offs_b[i]:offs_b[i + 1]] for i in ci])
        c_prev, c_next = prev_b[ci], next_b[ci]
        cpu_reference(c_prev, c_next, c_pts, c_off, min(8, n_s))
        r = cpu_reference(c_prev, c_next, c_pts, c_off, n_s)
        r1 = cpu_reference(c_prev, c_next, c_pts, c_off, min(n_s, 24), threads=1)
        cpu = {"value": r["tracked"] / r["seconds"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"],
               "sample": "%d pairs x %d corners (%.1f s), %s" % (r["pairs"], CORNERS, r["seconds"], r["what"]),
               "single_thread_value": r1["tracked"] / r1["seconds"], "host_cpus": os.cpu_count()}
        # parity of what was timed: the GPU results against the same CPU call on EVERY distinct pair of the sample
        from oracle import cv2_ref
        if cv2_ref.HAVE_CV2:
            cv2_ref.cv2.setNumThreads(os.cpu_count() or 1)
            p_all = nxt_d.cpu().numpy(); s_all = st_d.cpu().numpy()
            agree = n_all = n_both = n_over = 0
            max_d = 0.0
            for b in range(nb):  # batch position b holds distinct pair b (idx = arange % nb)
                sl = slice(int(offs[b]), int(offs[b + 1]))
                p_cv, s_cv, _ = cv2_ref.calc_optical_flow_pyr_lk(prev_b[b], next_b[b], pts_b[offs_b[b]:offs_b[b + 1]], None, WIN, MAX_LEVEL, CRIT, FLAGS)
                both = (s_cv == 1) & (s_all[sl] == 1)
                d = np.linalg.norm(p_cv.astype(np.float64) - p_all[sl], axis=1)
                agree += int((s_cv == s_all[sl]).sum()); n_all += len(s_cv); n_both += int(both.sum())
                n_over += int((d[both] > 0.01).sum()); max_d = max(max_d, float(d[both].max()) if both.any() else 0.0)
            cpu["parity"] = {"pairs": nb, "features": n_all, "status_agree": agree / max(n_all, 1), "jointly_tracked": n_both,
                             "n_over_0.01px": n_over, "max_dpos_px": max_d,
                             "gates": "north_star: status agreement >= 0.999, |dpos| <= 0.01 px on jointly tracked points"}

    if rank == 0:
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
               "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
               "dtype": "int32 fixed-point windows + fp32 2x2 solve", "data": data_kind, "config": cfg,
               "submitted_features_per_s": feats_all * args.steps / (ms_total * 1e-3), "tracked_fraction": tracked_all / feats_all,
               "gpu_launches": int(launches), "clocks": clocks, "roofline": roof,
               "replicas_consistent": replicas_ok}
        out.update(extra)
        if e2e:
            out["e2e"] = e2e
        if cpu:
            out["cpu_baseline"] = cpu
            # north_star target: >= 100x the reference's host-CPU tracked features/s on THIS box's cores, end to end
            ref_v = cpu["value"]
            out["target_100x_host_cpu"] = {"host_cores_used": cpu["cores"], "cpu_value": ref_v, "needed": 100.0 * ref_v,
                                           "device_resident_ratio": value / ref_v, "e2e_ratio": (e2e["value"] / ref_v) if e2e else None,
                                           "met": bool(e2e and e2e["value"] >= 100.0 * ref_v)}
        if latency:
            out["latency"] = latency
        emit(out)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
