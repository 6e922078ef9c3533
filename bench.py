#!/usr/bin/env python
"""bench.py — GeoMaskMaker + ORB frames/s at 640x480 on N B200s (BASELINE.json metric), with roofline and CPU baseline.

  python bench.py --gpus N --steps K --warmup W            # our arm (N>1: launched by torchrun, one rank per GPU)
  python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port) on the host cores

A "step" is one pass of the hot path over one batch: every one of the `batch` independent RGB-D streams of a rank
advances by one frame (gray x2, ORB extraction, AddNewImage products, Farneback flow (t-5,t), Mahalanobis scatter,
min-max/threshold mask).  Streams are sharded across ranks with no collective on the data path: weak scaling.

  value : frames/s, inputs already resident in HBM (gd_frontend_step_staged), CUDA events on the handle's stream,
          max over ranks.
  e2e   : the same through the reference-facing C ABI with pinned HOST buffers: H2D of BGR + the raw 16-bit depth image
          (gd_frontend_step_u16, what Tracking.cc:234-235 receives) and D2H of mask + keypoints + descriptors inside the
          timed region; `e2e.f32_depth` repeats it with the depth already converted to metres on the host (gd_frontend_step).
  parity_checked : after the timed region one stream's results of the LAST timed step are compared with the CPU oracle.
torch is used only for the multi-rank barrier / max-reduce (torch.distributed, NCCL); the product path is the C ABI.
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

W, H = 640, 480  # BASELINE.json headline resolution; --res WxH runs the resolution sweep of configs[4]
N_PX = W * H
METRIC = "GeoMaskMaker+ORB frames/sec at 640x480"
# algorithmic (compulsory) bytes per frame of the whole pipeline, steady state = 318.5 B/px
ALGO_BYTES_PER_FRAME = 318.5 * N_PX
FB_LEVEL_SUM = 1.328125  # (1 + 1/4 + 1/16 + 1/64) pixels over the 4 Farneback levels
# algorithmic bytes per frame of each kernel family (DESIGN.md "kernels" table)
FAMILY_BYTES = {
    "K0_gray": (3 + 1 + 1) * N_PX,
    "K1a_blur_resample": (4 * 1 + 4 * FB_LEVEL_SUM) * N_PX,
    "K1a_polyexp": (4 + 20) * FB_LEVEL_SUM * N_PX,
    "K1b_flow_iter": 3 * 56 * FB_LEVEL_SUM * N_PX,      # fused form (GD_FLOW_FUSED=1)
    "K1b_matrices": 3 * 68 * FB_LEVEL_SUM * N_PX,       # split form: R0 20 + R1 20 + flow 8 in, M 20 out
    "K1b_box_solve": 3 * 28 * FB_LEVEL_SUM * N_PX,      # split form: M 20 in, flow 8 out
    # M ping-pong form (default): per level matrices once (R0 20 + R1 20 + coarse flow 2, M 20 out), two fused
    # box/solve + next-matrices launches (M 20 in, R0 20 + R1 20, M 20 out) and a last box/solve (M 20 in, flow 8 out)
    "K1b_box_matrices": 2 * 80 * FB_LEVEL_SUM * N_PX,
    "K1b_flow_upsample": (8 * 0.328125 / 4 + 8 * 0.328125) * N_PX,
    "K2a_depth_edge": 5 * N_PX,
    "K2b_mahalanobis": 22 * N_PX,
    "K3a_minmax": 8 * N_PX,                             # 64-bit scatter keys
    "K3b_normalize_mask": 9 * N_PX,
    "K3_minmax_mask": 9 * N_PX,                         # cluster form: keys read once, mask written once
    "K4a_pyramid_resize": 2 * 2.094 * N_PX,
    "K4b_fast_cells": 3.094 * N_PX,
    "K4c_quadtree": 8 * 8000 * 4,
    "K4e_blur7": 2 * 3.094 * N_PX,
    "K4de_orient_describe": 1500 * (749 + 512 + 60),
}


# what actually bounds each kernel family on B200 (ncu, profiles/): reported beside the HBM fraction so that a low HBM
# fraction of a compute-bound kernel is not misread
FAMILY_LIMITER = {
    "K0_gray": "hbm",
    "K1a_blur_resample": "issue slots / L2 (narrow levels)",
    "K1a_polyexp": "issue slots (83 % busy); conversion (XU) pipe 58 %, LSU 61 % (fp64 horizontal pass like OpenCV)",
    "K1b_matrices": "hbm",
    "K1b_box_solve": "issue slots (66 % busy): f32 adds of the window sums + shared-memory exchange",
    "K1b_box_matrices": "hbm (69 % of the measured peak at level 0) + issue slots (59 % busy, 3 CTAs/SM)",
    "K2a_depth_edge": "issue slots (71 % busy; persistent CTAs with register prefetch, f32 interval test, fp64 only for undecided pixels)",
    "K2b_mahalanobis": "conversion (XU) pipe 70 % busy: 65 f32<->f64 conversions per pixel (bit-exact OpenCV accumulation widths)",
    "K3a_minmax": "hbm",
    "K3b_normalize_mask": "hbm",
    "K3_minmax_mask": "hbm",
    "K4a_pyramid_resize": "issue slots (byte gathers), L2 resident",
    "K4b_fast_cells": "issue slots (90 % busy: 4-point pass, 16-point network, cell bookkeeping), L2 resident",
    "K4c_quadtree": "latency (one CTA per level and stream)",
    "K4e_blur7": "latency / issue slots (rows requested seven iterations ahead), L2 resident",
    "K4de_orient_describe": "latency / gathers",
}


def _load_limiter_util():
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "limiters.json")))
        return {k: v for k, v in d.items() if not k.startswith("_")}
    except Exception:
        return {}


LIMITER_UTIL = _load_limiter_util()


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device = device
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.device)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def _gen_frame(args):
    stream, f, res = args
    synth = importlib.import_module("gd-slam_b200.synth")
    s = synth.SyntheticStream(stream, res[0], res[1])
    fr = s.frame(f)
    return stream, f, fr.bgr, fr.depth_m


def make_data(n_distinct, n_frames, seed0):
    """n_distinct synthetic streams x n_frames frames (seeded, SURVEY 8d generator), generated on all host cores."""
    from concurrent.futures import ProcessPoolExecutor

    synth = importlib.import_module("gd-slam_b200.synth")
    jobs = [(seed0 + s, f, (W, H)) for s in range(n_distinct) for f in range(n_frames)]
    bgr = np.empty((n_distinct, n_frames, H, W, 3), np.uint8)
    dep = np.empty((n_distinct, n_frames, H, W), np.float32)
    with ProcessPoolExecutor(max_workers=min(os.cpu_count() or 1, 16)) as ex:
        for s, f, b, d in ex.map(_gen_frame, jobs, chunksize=2):
            bgr[s - seed0, f] = b
            dep[s - seed0, f] = d
    poses = {}
    for s in range(n_distinct):
        st = synth.SyntheticStream.__new__(synth.SyntheticStream)
        st.roll, st.t = 0.0, np.asarray((0.002, -0.0012, 0.0))
        for f in range(n_frames):
            poses[(s, f)] = st.pair_pose((f - 5) % n_frames if f >= 5 else 0, f) if f >= 5 else st.pair_pose(0, 0)
    return bgr, dep, poses


# ------------------------------------------------------------------------------------------------ CPU reference arm
_CPU = {}


def workload_name():
    """config.workload: identical in both arms (the driver compares the strings)."""
    return f"synthetic {W}x{H} RGB-D streams, full GeoMaskMaker + ORB (TUM3 intrinsics, ORB 1500/1.2/8/20/7)"


def cpu_model():
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("model name"):
                return ln.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def use_native_oracle():
    """Point the CPU legs at an -O3 -march=native build of the oracle made on THIS machine; the
    portable -O2 parity build is the fallback.  Returns the description that goes into the JSON line."""
    from oracle import pyoracle as po

    p = po.build_native()
    if p:
        os.environ["GD_ORACLE_LIB"] = p  # inherited by the worker processes
        return "-O3 -march=native -ffp-contract=off (built on this host)"
    return "-O2 -ffp-contract=off (portable parity build; the native build failed)"


def _cpu_init(res):
    """Worker initialiser: one host process = one stream; frames of the pair (t-5, t) are generated once."""
    global W, H
    W, H = res
    from oracle import pyoracle as po

    synth = importlib.import_module("gd-slam_b200.synth")
    po.lib()
    s = synth.SyntheticStream(os.getpid() % 97, W, H)
    _CPU.update(po=po, K=synth.intrinsics(W, H), a=s.frame(0), b=s.frame(5), pose=s.pair_pose(0, 5))


def _cpu_step(n_frames):
    """Structured like the reference: every frame recomputes both gray images, both Farneback pyramids and both depth-edge
    maps (GeoMaskMaker.cc:158-199), then ORB on the new frame (Tracking.cc:238)."""
    po, K, a, b, (R, T) = _CPU["po"], _CPU["K"], _CPU["a"], _CPU["b"], _CPU["pose"]
    t0 = time.perf_counter()
    for _ in range(n_frames):
        po.orb_extract(po.gray(b.bgr, 1))
        po.geomask_pair(a.bgr, b.bgr, a.depth_m, b.depth_m, K, R, T)
    return n_frames, time.perf_counter() - t0


def _cpu_stage_split(_):
    """ms per frame of each stage of the CPU path on ONE core."""
    po, K, a, b, (R, T) = _CPU["po"], _CPU["K"], _CPU["a"], _CPU["b"], _CPU["pose"]
    out = {}

    def t(name, fn):
        t0 = time.perf_counter()
        r = fn()
        out[name] = (time.perf_counter() - t0) * 1e3
        return r

    g0 = t("gray_x2", lambda: (po.gray(a.bgr, 0), po.gray(b.bgr, 0)))
    flow = t("farneback_pair", lambda: po.farneback(g0[0], g0[1]))
    e = t("depth_edge_x2", lambda: (po.depth_edge(a.depth_m, K), po.depth_edge(b.depth_m, K)))
    dist = t("mahalanobis_loop", lambda: po.mahalanobis(flow, a.depth_m, b.depth_m, e[0], e[1], K, R, T)[0])
    t("normalize_threshold", lambda: po.normalize_threshold(dist))
    go = po.gray(b.bgr, 1)
    t("orb_extract", lambda: po.orb_extract(go))
    return out


class CpuReference:
    """The reference's CPU path (oracle port) on all host cores: one worker process per core, kept alive across steps."""

    def __init__(self, procs=None):
        from concurrent.futures import ProcessPoolExecutor

        self.build = use_native_oracle()
        self.procs = procs or (os.cpu_count() or 1)
        self.ex = ProcessPoolExecutor(max_workers=self.procs, initializer=_cpu_init, initargs=((W, H),))
        list(self.ex.map(_cpu_step, [0] * self.procs))  # start the workers, build their frames

    def step(self, frames_per_proc):
        res = list(self.ex.map(_cpu_step, [frames_per_proc] * self.procs))
        return sum(r[0] for r in res) / max(r[1] for r in res)

    def stage_split(self):
        return {k: round(v, 2) for k, v in list(self.ex.map(_cpu_stage_split, [0]))[0].items()}

    def describe(self):
        return {"cpu_model": cpu_model(), "nproc": os.cpu_count(), "oracle_build": self.build}

    def close(self):
        self.ex.shutdown()


def run_reference(args, rank, world):
    if rank != 0:
        return
    ref = CpuReference()
    procs = ref.procs
    # bounded sample: keep the whole run within ~2 minutes whatever K the driver passes (one frame takes ~0.25 s per core)
    per_step = max(1, min(args.ref_frames_per_proc, int(120.0 / (0.3 * max(1, args.steps + args.warmup)))))
    for _ in range(max(0, args.warmup)):
        ref.step(per_step)
    rates = []
    t_all = time.perf_counter()
    for _ in range(args.steps):
        rates.append(ref.step(per_step))
    ms = (time.perf_counter() - t_all) * 1e3 / max(1, args.steps)
    split = ref.stage_split()
    desc = ref.describe()
    ref.close()
    v = float(statistics.median(rates))
    sample = (f"{procs} host processes x {per_step} frame(s) per step ({W}x{H} pair (t-5,t) + ORB on the new frame), oracle port "
              "(ORB part = the reference's algorithm restated; equal to its verbatim compile on the fixtures)")
    out = {"impl": "reference", "metric": METRIC, "value": v, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
           "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f32", "data": "synthetic",
           "config": {"workload": workload_name(),
                      "note": "reference's CPU path restated (oracle/): OpenCV-4.13 semantics, one process per host core, "
                              "reference-structured (both pyramids / edge maps recomputed per call); the real reference "
                              "additionally allocates cv::Mat temporaries and calls Mat::inv() per pixel, so it is slower than this"},
           "cpu_baseline": {"value": v, "unit": "frames/s", "cores": procs, "kind": "port", "sample": sample,
                            "stage_ms_one_core": split, **desc},
           "e2e": {"value": v, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


# ------------------------------------------------------------------------------------------------ our arm
def bind_host_to_gpu(device):
    """Pin this rank's host threads (and, by first touch, its pinned staging buffers) to the CPUs NVML reports as local
    to the GPU: the e2e leg moves ~30 GB/s per GPU through host memory, which must not cross the socket interconnect.
    Best effort: returns the number of CPUs bound to, or None (GD_BENCH_AFFINITY=0 disables)."""
    if os.environ.get("GD_BENCH_AFFINITY", "1") == "0":
        return None
    try:
        import pynvml
        import torch

        pynvml.nvmlInit()
        p = torch.cuda.get_device_properties(device)
        bus = "%08x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        h = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
        pynvml.nvmlDeviceSetCpuAffinity(h)
        return len(os.sched_getaffinity(0))
    except Exception as e:  # affinity is an optimisation, never a requirement
        print(f"bench.py: host affinity not set ({type(e).__name__}: {e})", file=sys.stderr)
        return None


def check_last_step_against_oracle(res, bgr_ref, bgr_cur, dep_ref, dep_cur, K, R, T):
    """One stream of the last timed step against the CPU oracle: ORB exact, mask agreement >= 99.9 %."""
    from oracle import pyoracle as po

    mask, kp, desc = res
    rkp, rdesc, _ = po.orb_extract(po.gray(bgr_cur, 1))
    orb_ok = len(kp) == len(rkp) and all(np.array_equal(kp[f], rkp[f]) for f in kp.dtype.names) and np.array_equal(desc, rdesc)
    mo = po.geomask_pair(bgr_ref, bgr_cur, dep_ref, dep_cur, K, R, T)
    agree = float((mask == mo).mean())
    return {"orb_bit_exact": bool(orb_ok), "mask_agreement": agree, "n_keypoints": int(len(kp)),
            "dynamic_fraction": float((mask == 0).mean()), "ok": bool(orb_ok and agree >= 0.999)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=int(os.environ.get("GD_BENCH_BATCH", "64")), help="streams per GPU")
    ap.add_argument("--streams-total", type=int, default=0,
                    help="strong-scaling variant of SURVEY 8d config 4: this many streams in total, split evenly over the GPUs "
                         "(overrides --batch; e.g. 8 -> 8/4/2/1 streams per GPU on 1/2/4/8 GPUs)")
    ap.add_argument("--slots", type=int, default=12, help="distinct frames kept resident per stream")
    ap.add_argument("--distinct", type=int, default=4, help="distinct synthetic streams generated per rank")
    ap.add_argument("--ref-frames-per-proc", type=int, default=3)
    ap.add_argument("--e2e-handles", type=int, default=4, help="independent handles (host threads) the e2e leg drives per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-strong", action="store_true", help="skip the 8-streams-in-total (config 4 as written) side measurement")
    ap.add_argument("--no-getrt", action="store_true", help="skip the side measurement with the GetRt stage enabled")
    ap.add_argument("--quick", action="store_true", help="device-resident leg and kernel profile only (A/B runs of kernel variants)")
    ap.add_argument("--probe-handles", action="store_true",
                    help="experiment: device-resident throughput with the streams split over 1 / 2 / 4 handles (host threads)")
    ap.add_argument("--probe-pcie", action="store_true",
                    help="only time bare pinned cudaMemcpyAsync H2D / D2H on all ranks at once and print the rates")
    ap.add_argument("--res", default="640x480", help="WxH of the synthetic streams (configs[4]: 1280x720, 1920x1080)")
    args = ap.parse_args()
    global W, H, N_PX, ALGO_BYTES_PER_FRAME, METRIC
    W, H = (int(v) for v in args.res.lower().split("x"))
    scale_px = (W * H) / N_PX
    N_PX = W * H
    ALGO_BYTES_PER_FRAME = 318.5 * N_PX
    for k in list(FAMILY_BYTES):
        if k not in ("K4c_quadtree", "K4de_orient_describe"):
            FAMILY_BYTES[k] *= scale_px
    if (W, H) != (640, 480):
        METRIC = f"GeoMaskMaker+ORB frames/sec at {W}x{H}"

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    capi = importlib.import_module("gd-slam_b200.capi")
    synth = importlib.import_module("gd-slam_b200.synth")
    capi.lib()
    if capi.device_count() < 1:
        raise SystemExit("bench.py: no CUDA device (libgdslam_cuda has no CPU fallback)")
    dist = None
    host_cpus = bind_host_to_gpu(local_rank) if world > 1 else None  # a single rank keeps every core of the box
    if world > 1:
        # stdout carries the single JSON line.  NCCL honours NCCL_DEBUG_FILE only above the VERSION level (the image sets
        # NCCL_DEBUG=VERSION, whose banner goes to stdout), so raise VERSION/unset to WARN and send the log to stderr.
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() in ("VERSION", "NONE", ""):
            os.environ["NCCL_DEBUG"] = "WARN"
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        import torch
        import torch.distributed as dist

        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    device = local_rank
    if args.streams_total > 0:
        if args.streams_total % world:
            raise SystemExit("--streams-total must be a multiple of the number of GPUs")
        args.batch = args.streams_total // world
    B, S, K_, Wm = args.batch, args.slots, args.steps, max(3, args.warmup)

    def barrier():
        if dist is not None:
            dist.barrier()

    sharding = importlib.import_module("gd-slam_b200.sharding")

    def max_over_ranks(x):
        return sharding.max_over_ranks(x, dist, f"cuda:{local_rank}" if dist is not None else None)

    def min_over_ranks(x):
        return -max_over_ranks(-x)

    # ---- host-side copy ceiling: bare pinned cudaMemcpyAsync on every rank at the same time (no kernel involved)
    def probe():
        barrier()
        h2d = capi.probe_copy(device, 256 << 20, 8, True)
        barrier()
        d2h = capi.probe_copy(device, 256 << 20, 8, False)
        barrier()
        return min_over_ranks(h2d), min_over_ranks(d2h)

    if args.probe_pcie:
        h2d, d2h = probe()
        if rank == 0:
            print(json.dumps({"probe": "pinned cudaMemcpyAsync, 8 x 256 MiB per direction, all ranks at once, slowest rank",
                              "n_gpus": world, "h2d_gbs_per_gpu": h2d, "d2h_gbs_per_gpu": d2h,
                              "h2d_gbs_box": h2d * world, "d2h_gbs_box": d2h * world, "host_cpus_per_rank": host_cpus}), flush=True)
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- data: every rank owns its own `B` streams (weak scaling); `distinct` seeded streams are generated per rank and
    #      replicated over the batch with a frame offset
    D = min(args.distinct, B)
    bgr, dep, poses = make_data(D, S, seed0=sharding.stream_seed(rank, 0, B))
    K = synth.intrinsics(W, H)
    fe = capi.Frontend(K, W, H, batch=B, device=device, staged_slots=S)
    hb = capi.pinned_empty((S, B, H, W, 3), np.uint8)
    hd = capi.pinned_empty((S, B, H, W), np.float32)
    hd16 = capi.pinned_empty((S, B, H, W), np.uint16)  # the raw TUM depth image: metres = u16 * (1 / 5000.f), exactly
    Rs = np.zeros((S, B, 3, 3), np.float32)
    Ts = np.zeros((S, B, 3), np.float32)
    d16 = np.rint(dep.astype(np.float64) * 5000.0).astype(np.uint16)
    assert np.array_equal(d16.astype(np.float32) * np.float32(1.0 / 5000.0), dep)
    for s in range(S):
        for b in range(B):
            src, off = b % D, (b // D) % S
            f = (s + off) % S
            hb[s, b] = bgr[src, f]
            hd[s, b] = dep[src, f]
            hd16[s, b] = d16[src, f]
            Rs[s, b], Ts[s, b] = poses[(src, f)]
        fe.stage(s, hb[s], hd[s])
    del bgr, dep, d16

    # L2 hygiene: the default batch streams far more than the 126 MB L2 per step; small batches flush it before every step
    ws_mb = B * (2 * 8.2 + 2.2 + 2 * 2.5 + 2.5 + 3.3) * N_PX / 307200
    flush = ws_mb < 2 * 126
    step_i = [0]

    def step_staged():
        s = step_i[0] % S
        if flush:
            fe.flush_l2()
        fe.step_staged(s, Rs[s], Ts[s])
        step_i[0] += 1

    # ---- device-resident throughput (value)
    for _ in range(6 + Wm):  # fill the 6-frame ring, then W warm-up steps
        step_staged()
    fe.sync()
    barrier()
    sampler = ClockSampler(device)
    sampler.start()
    l0 = fe.launch_count()
    fe.sync()
    fe.timer_begin()
    for _ in range(K_):
        step_staged()
    ms_total = fe.timer_end()
    fe.sync()
    launches = fe.launch_count() - l0
    barrier()
    ms_total = max_over_ranks(ms_total)
    value = world * B * K_ / (ms_total * 1e-3)

    if args.probe_handles:  # how much does interleaving independent groups of streams help? (wall clock, synchronised ends)
        res = {}
        for NHp in (1, 2, 4):
            Bp = B // NHp
            hs = [capi.Frontend(K, W, H, batch=Bp, device=device, staged_slots=S) for _ in range(NHp)]
            for i, hnd in enumerate(hs):
                for s in range(S):
                    hnd.stage(s, hb[s, i * Bp:(i + 1) * Bp], hd[s, i * Bp:(i + 1) * Bp])

            def loop(i, n, ev):
                ev.wait()
                for k in range(n):
                    hs[i].step_staged(k % S, Rs[k % S, i * Bp:(i + 1) * Bp], Ts[k % S, i * Bp:(i + 1) * Bp])
                hs[i].sync()

            def run(n):
                ev = threading.Event()
                th = [threading.Thread(target=loop, args=(i, n, ev)) for i in range(NHp)]
                for t in th:
                    t.start()
                t0 = time.perf_counter()
                ev.set()
                for t in th:
                    t.join()
                return time.perf_counter() - t0

            run(6 + 12 + 5)
            dt = run(K_)
            res[NHp] = B * K_ / dt
            for hnd in hs:
                hnd.close()
        print(json.dumps({"probe": "device-resident frames/s with the batch split over N handles", "streams": B,
                          "single_handle_event_timed": value, "by_handles": res}), flush=True)
        return

    # ---- parity of what was just timed: stream 0 (and the last stream) of the LAST timed step against the CPU oracle
    parity = None
    if rank == 0:
        last = (step_i[0] - 1) % S
        ref_slot = (step_i[0] - 1 - 5) % S
        res = fe.fetch()
        checks = [check_last_step_against_oracle(res[b], hb[ref_slot, b], hb[last, b], hd[ref_slot, b], hd[last, b], K,
                                                 Rs[last, b], Ts[last, b]) for b in sorted({0, B - 1})]
        parity = {"ok": all(c["ok"] for c in checks), "streams": sorted({0, B - 1}), "checks": checks,
                  "what": "ORB keypoints/descriptors bit-exact and mask agreement >= 0.999 vs the CPU oracle on the last timed step"}

    # ---- SURVEY 8d config 4 as written: 8 streams in total over the GPUs (strong scaling, latency-bound regime)
    strong = None
    if not args.no_strong and not args.quick and args.streams_total == 0 and 8 % world == 0 and B >= 8 // world:
        bs = 8 // world
        fs = capi.Frontend(K, W, H, batch=bs, device=device, staged_slots=S)
        for s in range(S):
            fs.stage(s, hb[s, :bs], hd[s, :bs])
        k, t_w = 0, time.perf_counter()
        while k < 6 + 2 * 6 + 40 or time.perf_counter() - t_w < 1.0:  # ring, graph capture of the ring phases, then at least
            fs.step_staged(k % S, Rs[k % S, :bs], Ts[k % S, :bs])     # 1 s of load: the GPU idled during the oracle check
            k += 1                                                    # above and needs that long to be back at full clocks
            if k % 64 == 0:
                fs.sync()
        fs.sync()
        barrier()
        KS = 2 * K_  # a step takes well under a millisecond here: three repetitions of twice as many steps, median reported
        reps = []
        for _ in range(3):
            fs.timer_begin()
            for k in range(KS):
                fs.step_staged(k % S, Rs[k % S, :bs], Ts[k % S, :bs])
            reps.append(max_over_ranks(fs.timer_end()))
        ms_s = sorted(reps)[1]
        fs.close()
        strong = {"streams_total": 8, "streams_per_gpu": bs, "value": 8 * KS / (ms_s * 1e-3), "unit": "frames/s",
                  "ms_per_step": ms_s / KS, "steps": KS, "repetitions_ms_per_step": [r / KS for r in reps], "scaling": "strong",
                  "note": "BASELINE configs[3] as written; per-step working set below the L2 size, not flushed"}

    # ---- end to end through the C ABI with pinned host buffers (e2e)
    #      The batch is driven as `args.e2e_handles` independent handles (B / handles streams each) from as many host
    #      threads: gd_frontend_step* is synchronous per handle (the reference's contract), so one handle's PCIe copies
    #      overlap the other's kernels.  Every step still moves all B frames host->device and all results device->host.
    NH = max(1, min(args.e2e_handles, B))
    while B % NH:
        NH -= 1
    if NH == 1:
        fes = [fe]
    else:
        fes = [capi.Frontend(K, W, H, batch=B // NH, device=device) for _ in range(NH)]
    Bh = B // NH

    def host_loop(i, nsteps, start_evt, u16):
        f = fes[i]
        sl = slice(i * Bh, (i + 1) * Bh)
        start_evt.wait()
        for k in range(nsteps):
            s = k % S
            if u16:
                f.step_u16(hb[s, sl], hd16[s, sl], Rs[s, sl], Ts[s, sl])
            else:
                f.step(hb[s, sl], hd[s, sl], Rs[s, sl], Ts[s, sl])
        f.sync()

    def run_host(nsteps, u16):
        ev = threading.Event()
        th = [threading.Thread(target=host_loop, args=(i, nsteps, ev, u16)) for i in range(NH)]
        for t in th:
            t.start()
        t0 = time.perf_counter()
        ev.set()
        for t in th:
            t.join()
        return time.perf_counter() - t0

    Ke = 2 if args.quick else K_
    run_host(6 + 3 + min(K_, 24), True)  # fill the rings of the e2e handles + warm-up (copy engines, host threads, clocks)
    barrier()
    e2e_s = run_host(Ke, True) * (K_ / Ke)
    barrier()
    e2e_s = max_over_ranks(e2e_s)
    run_host(3, False)
    barrier()
    e2e32_s = run_host(Ke, False) * (K_ / Ke)
    clocks = sampler.stop()
    barrier()
    e2e32_s = max_over_ranks(e2e32_s)
    e2e_value = world * B * K_ / e2e_s
    d2h = B * (N_PX + fe.cap * (28 + 32) + 4)
    h2d_u16 = B * (N_PX * 3 + N_PX * 2)
    h2d_f32 = B * (N_PX * 3 + N_PX * 4)
    if NH > 1:
        for f in fes:
            f.close()
    ceil_h2d, ceil_d2h = probe()

    # ---- the same step with GeoMaskMaker::GetRt's GPU half included (cv::ORB features of the new frame, matching against the
    #      frame five steps back, the 100 solvePnPRansac points fetched to the host); the pose itself is still the given one
    #      (cv::solvePnPRansac is host OpenCV in the reference and stays with the caller)
    with_getrt = None
    if not args.no_getrt and not args.quick:
        fg = capi.Frontend(K, W, H, batch=B, device=device, staged_slots=S, getrt=True)
        for s in range(S):
            fg.stage(s, hb[s], hd[s])
        for k in range(6 + Wm):
            fg.step_staged(k % S, Rs[k % S], Ts[k % S])
        fg.sync()
        barrier()
        fg.timer_begin()
        for k in range(K_):
            fg.step_staged((6 + Wm + k) % S, Rs[(6 + Wm + k) % S], Ts[(6 + Wm + k) % S])
        ms_g = max_over_ranks(fg.timer_end())
        pts = fg.fetch_getrt()
        npts = [len(p[0]) for p in pts]
        fg.profile(True)
        for k in range(2):
            fg.step_staged((6 + Wm + K_ + k) % S, Rs[(6 + Wm + K_ + k) % S], Ts[(6 + Wm + K_ + k) % S])
        fg.profile(False)
        gfam = {n: ms / 2 for n, ms, _ in fg.profile_read() if n.startswith("G")}
        fg.close()
        with_getrt = {"value": world * B * K_ / (ms_g * 1e-3), "unit": "frames/s", "ms_per_step": ms_g / K_,
                      "getrt_ms_per_frame_amortised": (ms_g - ms_total) / K_ / B,
                      "points_per_stream_min_max": [int(min(npts)), int(max(npts))],
                      "getrt_kernel_ms_per_step_serialised": gfam,
                      "note": "GetRt up to solvePnPRansac on the GPU every step (features cached per ring slot); pose = given"}

    # ---- per-kernel-family device time (events on the handle's stream, serialised) -> dominant kernel + roofline
    fe.profile(True)
    PSTEPS = 3
    for _ in range(PSTEPS):
        step_staged()
    fe.profile(False)
    fam = fe.profile_read()
    peak, peak_src = load_peaks()
    if any(name == "K1b_box_matrices" for name, _, _ in fam):  # M ping-pong form: one launch of each per level
        FAMILY_BYTES["K1b_matrices"] = 62 * FB_LEVEL_SUM * N_PX
        FAMILY_BYTES["K1b_box_solve"] = 28 * FB_LEVEL_SUM * N_PX
    tot_ms = sum(ms for _, ms, _ in fam) or 1.0
    families = {}
    for name, ms, ln in fam:
        by = FAMILY_BYTES.get(name, 0.0) * B * PSTEPS
        gbs = by / (ms * 1e-3) / 1e9 if ms > 0 else None
        families[name] = {"ms_per_step": ms / PSTEPS, "launches_per_step": ln / PSTEPS, "share": ms / tot_ms,
                          "achieved_gbs": gbs, "frac_of_hbm_peak": gbs / peak if gbs else None,
                          "limiter": FAMILY_LIMITER.get(name)}
        if name in LIMITER_UTIL:  # ncu: how busy the limiting resource is (static, from profiles/limiters.json)
            families[name]["limiter_util"] = LIMITER_UTIL[name]
    dom = max(fam, key=lambda x: x[1])
    dom_name, dom_ms, dom_ln = dom
    dom_bytes_per_launch = FAMILY_BYTES.get(dom_name, 0.0) * B * PSTEPS / max(1, dom_ln)
    achieved = dom_bytes_per_launch / (dom_ms / max(1, dom_ln) * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            tj = json.load(open(tp))
            traffic = tj.get(dom_name)
            if traffic is not None:  # captured at tj["_streams"] streams of 640x480 per launch; per-launch traffic scales with batch and pixels
                traffic = traffic * B / float(tj.get("_streams", B)) * (N_PX / float(640 * 480))
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "kernel": dom_name, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "peak_source": peak_src, "algorithmic_bytes_per_launch": dom_bytes_per_launch,
                "avg_launch_ms": dom_ms / max(1, dom_ln),
                "pipeline": {"algorithmic_bytes_per_frame": ALGO_BYTES_PER_FRAME,
                             "achieved": value / world * ALGO_BYTES_PER_FRAME / 1e9, "frac": value / world * ALGO_BYTES_PER_FRAME / 1e9 / peak},
                "families": families}

    if flush:
        l2 = ("L2 flushed by a 256 MiB memset before every step, inside the timed region (per-step working set %.0f MB per GPU "
              "is not larger than twice the 126 MB L2)" % ws_mb)
    else:
        l2 = ("inputs larger than L2: per-step working set %.0f MB per GPU (ring of polynomial-expansion pyramids + staged "
              "frames), 126 MB L2" % ws_mb)
    out = {"metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": K_, "warmup": Wm,
           "ms_per_step": ms_total / K_, "higher_is_better": True, "scaling": "strong" if args.streams_total > 0 else "weak",
           "vs_baseline": None, "dtype": "f32",
           "data": "synthetic",
           "config": {"workload": workload_name(),
                      "note": "steady state: per-image products cached in the 6-deep device ring",
                      "streams_per_gpu": B, "frames_per_step": B * world, "resident_frames_per_stream": S,
                      "l2_hygiene": l2, "sharding": "independent streams per GPU, no collective"},
           "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": h2d_u16, "d2h_bytes_per_step": d2h,
                   "ms_per_step": e2e_s * 1e3 / K_, "handles_per_gpu": NH, "streams_per_handle": Bh,
                   "depth_input": "raw 16-bit TUM depth, converted on the device (gd_frontend_step_u16)",
                   "h2d_gbs": h2d_u16 * K_ / e2e_s / 1e9, "d2h_gbs": d2h * K_ / e2e_s / 1e9,
                   "h2d_ceiling_gbs": ceil_h2d, "d2h_ceiling_gbs": ceil_d2h,
                   "ceiling_note": "per GPU: bare pinned cudaMemcpyAsync (8 x 256 MiB) on all ranks at once, slowest rank",
                   "f32_depth": {"value": world * B * K_ / e2e32_s, "ms_per_step": e2e32_s * 1e3 / K_,
                                 "h2d_bytes_per_step": h2d_f32, "h2d_gbs": h2d_f32 * K_ / e2e32_s / 1e9},
                   "host_cpus_per_rank": host_cpus},
           "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline}
    if parity is not None:
        out["parity_checked"] = parity["ok"]
        out["parity"] = parity
    if strong is not None:
        out["strong_scaling_8_streams"] = strong
    if with_getrt is not None:
        out["with_getrt_stage"] = with_getrt

    if args.quick:
        out["e2e"]["note"] = "--quick: 2 steps only, not a measurement"
    if rank == 0 and world == 1 and not args.no_cpu_baseline and not args.quick:
        t0 = time.perf_counter()
        ref = CpuReference()
        per = 3
        v = ref.step(per)
        split = ref.stage_split()
        desc = ref.describe()
        ref.close()
        out["cpu_baseline"] = {"value": v, "unit": "frames/s", "cores": ref.procs, "kind": "port",
                               "sample": f"{ref.procs} host processes x {per} frames of the same workload through the oracle "
                                         f"(reference-structured: both pyramids/edge maps per call), {time.perf_counter() - t0:.1f} s wall",
                               "stage_ms_one_core": split, **desc}
    fe.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
