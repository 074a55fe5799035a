#!/usr/bin/env python3
"""bench.py — map frames/s of the per-frame vision path on B200 (BASELINE.json metric).

A "step" is one pass of the whole hot path (fused pixel pass + on-device line search) over one batch of
synthetic 2560x1440 captures per GPU (BASELINE config 3: dense SL markers + lines, batch 1024).  Frames are
independent, so N GPUs each process their own 1024 frames with no collective ("weak" scaling);
torch.distributed (NCCL) is only used for the barrier and for gathering per-rank timings.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Prints ONE JSON line (rank 0).  Keys beyond the base contract:
  roofline       HBM roofline of the bandwidth-bound kernel (k_pixels2 + k_px_seams), CUDA-event timed in the step
  roofline_step  the same for the whole step (algorithmic bytes of a step / ms_per_step / peak)
  kernels        per-kernel time inside the timed region (pixel pass, line search) + LSD work figures
  per_rank       ms_per_step / pixel pass / line search of every rank (names the cause of any N-GPU shortfall)
  check          the GPU's lines of the first frames compared with the CPU oracle inside this run
  cpu_baseline   the CPU oracle (a restatement of vision-cpu) timed on this box's host cores
  e2e            same metric through the C ABI with HOST frames (H2D + D2H inside the timed region), 1024 frames per
                 step at every N through a bounded ring of pinned buffers; with the measured bus ceiling beside it
  latency        one frame through the trait calls (config 2: 1920x1080; and 2560x1440): wall-clock + device time
  config4        BASELINE config 4 (3840x2160, 4096 frames frame-sharded over the ranks; one rank: a 1/8 shard)
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "map frames/sec (1440p)"
UNIT = "frames/s"
W, H = 2560, 1440
MAX_GAP = 15
WORKLOAD = "2560x1440 synthetic captures, dense SL markers + lines (8 marker lines, 12 blips, 2 scale labels), batch 1024 per GPU"
RING = 128  # pinned host buffers of the end-to-end leg


def algorithmic_bytes_per_frame(roi) -> int:
    """read button ROI + map ROI (BGRA); write ui_map (RGBA) + lsd_image + ocr_out
    + scales + 32 lines + flags."""
    bw, bh = roi["button"][2:]
    w, h = roi["map"][2:]
    qw, qh = roi["brq"][2:]
    return 4 * (bw * bh + w * h) + 4 * w * h + w * h + 2 * qw * qh + 512 + 8


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for nme, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def cpu_oracle(frames_np, threads: int, budget_s: float = 20.0):
    """Time the CPU oracle (oracle/ — a port of vision-cpu) on a bounded sample of the workload; keep its lines."""
    from oracle import oracle_py as O
    O.set_threads(threads)
    o = O.Oracle()
    o.process_frame(frames_np[0], True, 0, MAX_GAP)  # warm-up (thread pool, page faults)
    t0 = time.perf_counter()
    lines = []
    for f in frames_np:
        lines.append(o.process_frame(f, True, 0, MAX_GAP)[1])
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    return len(lines) / dt, lines


def run_reference(args, rank: int, world: int) -> None:
    """--impl reference: the reference's CPU implementation of the path.  Its Rust crate cannot be built here (no
    cargo/rustc, crates.io deps absent), so this times the oracle port with all host threads on a bounded sample of
    the same workload.  Nothing of the product is imported or loaded: the frames come from the oracle's own build of
    the shared generator (oracle_py.synth_frames)."""
    if rank != 0:
        return
    from oracle import oracle_py as O
    cores = os.cpu_count() or 1
    n_sample = 8
    frames = O.synth_frames(W, H, 0, n_sample)
    O.set_threads(cores)
    o = O.Oracle()
    for _ in range(args.warmup):
        o.process_frame(frames[0], True, 0, MAX_GAP)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        for f in frames:
            o.process_frame(f, True, 0, MAX_GAP)
    dt = time.perf_counter() - t0
    fps = args.steps * n_sample / dt
    sample = f"{n_sample} frames of the workload per step (of 1024), {cores} OpenMP threads"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8+f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": fps, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def gather(world, dist, torch, values):
    """[rank][k] of a list of floats."""
    if world == 1:
        return [list(values)]
    t = torch.tensor(values, device="cuda", dtype=torch.float64)
    out = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(out, t)
    return [[float(x) for x in o.cpu()] for o in out]


def bus_ceiling(torch, barrier, seconds=0.25):
    """Pinned-copy bandwidth of this rank, every rank at once (as in the e2e leg): uploads alone, downloads alone, and
    both directions together.  -> GB/s (h2d_alone, d2h_alone, h2d_with_d2h, d2h_with_h2d)"""
    nbytes = 128 << 20
    hin = torch.empty(nbytes, dtype=torch.uint8).pin_memory(); hout = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    din = torch.empty(nbytes, dtype=torch.uint8, device="cuda"); dout = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def run(up, down):
        torch.cuda.synchronize()
        barrier()
        e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        with torch.cuda.stream(s1):
            e[0].record()
        with torch.cuda.stream(s2):
            e[2].record()
        reps, t0 = 0, time.perf_counter()
        while time.perf_counter() - t0 < seconds:
            if up:
                with torch.cuda.stream(s1):
                    din.copy_(hin, non_blocking=True)
            if down:
                with torch.cuda.stream(s2):
                    hout.copy_(dout, non_blocking=True)
            reps += 1
            if reps % 8 == 0:
                torch.cuda.synchronize()
        with torch.cuda.stream(s1):
            e[1].record()
        with torch.cuda.stream(s2):
            e[3].record()
        torch.cuda.synchronize()
        return (reps * nbytes / (e[0].elapsed_time(e[1]) * 1e-3) / 1e9 if up else 0.0,
                reps * nbytes / (e[2].elapsed_time(e[3]) * 1e-3) / 1e9 if down else 0.0)
    run(True, True)
    up_alone = run(True, False)[0]
    down_alone = run(False, True)[1]
    up_both, down_both = run(True, True)
    return up_alone, down_alone, up_both, down_both


def run_e2e(args, vision, frames, roi, n, world, barrier, torch, dist, oracle_lines):
    """Same metric through the C ABI with pinned HOST frames in and host results out: every step streams n frames
    through a ring of RING pinned buffers (smh_b200_process_host_stream) — 1024 frames per step at every N."""
    import numpy as np
    ring = min(RING, n)
    host_frames = torch.empty((ring, H, W, 4), dtype=torch.uint8).pin_memory()
    host_frames.copy_(frames[:ring])
    _, _, mw, mh = roi["map"]; _, _, qw, qh = roi["brq"]

    def outs_for(l8):
        return {"ui_map": torch.empty((ring, mh, mw) if l8 else (ring, mh, mw, 4), dtype=torch.uint8).pin_memory(),
                "ocr_out": torch.empty((ring, qh, qw), dtype=torch.uint8).pin_memory(),
                "scales": torch.empty((ring, qh, qw), dtype=torch.uint8).pin_memory(),
                "lines": torch.empty((ring, 32, 4), dtype=torch.float32).pin_memory(),
                "n_lines": torch.empty((ring,), dtype=torch.int32).pin_memory(),
                "red_pixels": torch.empty((ring,), dtype=torch.int32).pin_memory(),
                "map_open": torch.empty((ring,), dtype=torch.uint8).pin_memory()}

    gx0 = roi["map"][0] & ~3
    nvec = (roi["map"][0] - gx0 + mw + 3) // 4
    h2d = n * (min(4 * nvec, W - gx0) * 4 * mh + roi["button"][2] * roi["button"][3] * 4)
    res = {}
    e_steps = max(3, min(args.steps, 5))
    for name, l8 in (("rgba", False), ("l8", True)):
        outs = outs_for(l8)
        ptrs = {k: v.data_ptr() for k, v in outs.items()}

        def step():
            vision.process_host_batch(None, frames_ptr=host_frames.data_ptr(), shape=(ring, H, W), max_gap=MAX_GAP, grayscale=True,
                                      chunk=32, want=("ui_map", "ocr_out", "scales"), out=dict(ptrs), ui_l8=l8, ring=ring, n_total=n)
        for _ in range(2):
            step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e_steps):
            step()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        per_rank = [r[0] for r in gather(world, dist, torch, [dt])]
        # the ring's slots hold the results of the last `ring` frames = frames 0 .. ring-1 again: compare with the oracle
        ok = None
        if oracle_lines:
            ok = all(int(outs["n_lines"][i]) == len(ol) and np.array_equal(outs["lines"][i, :len(ol)].numpy().view(np.uint32), ol.view(np.uint32))
                     for i, ol in enumerate(oracle_lines[:ring]))
        # grayscale ui_map: only its L plane crosses the bus (widened to RGBA by the library's host threads, or handed over as L8)
        d2h = n * (mw * mh + 2 * qw * qh + 512 + 9)
        res[name] = {"value": world * n * e_steps / max(per_rank), "per_rank_frames_per_s": [n * e_steps / t for t in per_rank],
                     "d2h_bytes_per_step": d2h, "lines_match_oracle": ok}
        del outs
    bw = gather(world, dist, torch, list(bus_ceiling(torch, barrier)))
    h2d_pf, d2h_pf = h2d / n, res["rgba"]["d2h_bytes_per_step"] / n
    # A frame needs its upload and its download.  While the shorter of the two is on the bus the other one runs at its
    # two-way rate, for the rest of its bytes at its one-way rate: the least time the longer direction can take.
    def frame_time(b):
        up1, dn1, up2, dn2 = (x * 1e9 for x in b)
        t_dn, t_up = d2h_pf / dn2, h2d_pf / up2            # each direction alone at its two-way rate
        if t_up >= t_dn:                                    # uploads are the longer direction
            return t_dn + max(h2d_pf - t_dn * up2, 0.0) / up1
        return t_up + max(d2h_pf - t_up * dn2, 0.0) / dn1
    ceiling = sum(1.0 / frame_time(b) for b in bw)
    return {"value": res["rgba"]["value"], "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": res["rgba"]["d2h_bytes_per_step"],
            "frames_per_step": n, "steps": e_steps, "ring_frames": ring,
            "per_rank_frames_per_s": res["rgba"]["per_rank_frames_per_s"], "lines_match_oracle": res["rgba"]["lines_match_oracle"],
            "ui_l8": {"value": res["l8"]["value"], "per_rank_frames_per_s": res["l8"]["per_rank_frames_per_s"],
                      "note": "SMH_BATCH_UI_L8: the caller takes ui_map as its L plane (no widening pass on the host)"},
            "bus": {"per_rank_pinned_copy_GBs": {"h2d_alone": [b[0] for b in bw], "d2h_alone": [b[1] for b in bw],
                                                 "h2d_with_d2h": [b[2] for b in bw], "d2h_with_h2d": [b[3] for b in bw]},
                    "ceiling_frames_per_s": ceiling, "frac_of_ceiling": res["rgba"]["value"] / ceiling if ceiling else None,
                    "note": "every rank copying 128 MiB pinned blocks at once; per rank and frame: the shorter direction's time at its two-way rate, "
                            "during which the longer direction moves at its two-way rate, + the rest of the longer direction's bytes at its "
                            "one-way rate; ceiling = sum over ranks"},
            "note": "smh_b200_process_host_stream: a ring of pinned host frames -> map + button rectangles up -> kernels -> ui_map/ocr/scales/lines "
                    "back into a ring of pinned result buffers, 3 chunks of 32 frames in flight; the grayscale ui_map crosses the bus as its L plane "
                    "and is widened to RGBA in the caller's buffer by the library's host threads inside the timed region"}


def run_latency(vision, synth, torch):
    """One frame through the trait calls in the orchestrator's order (src/vision/mod.rs:66-223): from a pageable host frame
    (what the reference's capture hands over) and from a pinned one (a capture loop that allocates its frame pool pinned)."""
    out = {}
    for name, (w, h) in (("1920x1080", (1920, 1080)), ("2560x1440", (W, H))):
        pageable = synth.frames_host(synth.SynthParams(w, h), 0, 1)[0]
        pinned = torch.from_numpy(pageable).pin_memory().numpy()
        for kind, frame in (("", pageable), ("pinned_frame", pinned)):
            wall, dev, shares = [], [], None
            for it in range(12):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                vision.load_frame(frame)
                got = vision.crop_to_map(True)
                assert got is not None
                vision.isolate_map_markers(); vision.mask_marker_lines()
                vision.find_marker_lines(MAX_GAP)
                vision.ocr_preprocess(); vision.find_scales_preprocess(0)
                dt = (time.perf_counter() - t0) * 1e3
                ts = vision.timeshares()
                if it >= 2:
                    wall.append(dt); dev.append(sum(v[1] for v in ts.values())); shares = ts
            rec = {"wall_ms": statistics.median(wall), "device_ms": statistics.median(dev),
                   "device_ms_without_upload": statistics.median(dev) - shares["load_frame"][1],
                   "stages_host_ms_device_ms": {k: [round(v[0], 4), round(v[1], 4)] for k, v in shares.items() if v[0] or v[1]}}
            if kind:
                out[name][kind] = rec
            else:
                out[name] = rec
    return out


def run_config4(args, vision, synth, shard, torch, dist, rank, world, barrier, roi_of, peak):
    """BASELINE config 4: 3840x2160, 4096 frames frame-sharded over the ranks (strong scaling, no collective)."""
    W4, H4, total = 3840, 2160, 4096
    if world == 1:
        first, cnt, note = 0, total // 8, "one GPU: the shard one rank of eight would take (512 frames)"
    else:
        first, last = shard.shard_range(rank, world, total)
        cnt, note = last - first, f"4096 frames split into {world} contiguous ranges"
    frames = torch.empty((cnt, H4, W4, 4), dtype=torch.uint8, device="cuda")
    synth.frames_device(vision, synth.SynthParams(W4, H4), first, cnt, frames.data_ptr())
    torch.cuda.synchronize()
    st = torch.cuda.current_stream().cuda_stream
    for _ in range(2):
        vision.process_batch(frames.data_ptr(), cnt, W4, H4, max_gap=MAX_GAP, grayscale=True, stream=st)
    barrier()
    steps = 3
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    for _ in range(steps):
        vision.process_batch(frames.data_ptr(), cnt, W4, H4, max_gap=MAX_GAP, grayscale=True, stream=st)
    e1.record()
    vision.process_batch(frames.data_ptr(), cnt, W4, H4, max_gap=MAX_GAP, grayscale=True, skip_lsd=True, stream=st)
    e2.record()
    barrier()
    ms, px = e0.elapsed_time(e1) / steps, e1.elapsed_time(e2)
    per = gather(world, dist, torch, [ms, px, float(cnt)])
    del frames
    torch.cuda.empty_cache()
    bpf = algorithmic_bytes_per_frame(roi_of(W4, H4))
    t_max = max(p[0] for p in per)
    return {"resolution": [W4, H4], "frames_total": int(sum(p[2] for p in per)), "scaling": "strong" if world > 1 else "one shard",
            "value": sum(p[2] for p in per) / (t_max * 1e-3), "unit": UNIT, "ms_per_step": t_max, "batches_in_flight": 1,
            "per_rank_ms_per_step": [p[0] for p in per], "per_rank_pixel_pass_ms": [p[1] for p in per],
            "pixel_pass_roofline_frac": [bpf * p[2] / (p[1] * 1e-3) / 1e9 / peak for p in per], "note": note}


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=1024, help="frames per GPU per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="profiling runs: skip the end-to-end, latency and config-4 legs")
    ap.add_argument("--in-flight", type=int, default=3,
                    help="batches in flight: steps alternate between this many handles/streams (1 = one stream)")
    ap.add_argument("--serial", action="store_true",
                    help="one launch per kernel on one stream (no chunk overlap): the form the ncu launch list is taken on")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import numpy as np
    import torch
    import torch.distributed as dist

    import smh_b200  # noqa: F401
    from squad_mortar_helper_b200 import shard, synth
    from squad_mortar_helper_b200.vision import B200Vision, roi as roi_of

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the backend has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        # stdout carries ONE JSON line: NCCL's own log (the version banner this image's NCCL_DEBUG setting prints
        # on the first collective) goes to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":  # the banner is printf'ed to stdout whatever the log file
            del os.environ["NCCL_DEBUG"]
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    n = args.frames
    roi = roi_of(W, H)
    bpf = algorithmic_bytes_per_frame(roi)
    vision = B200Vision.init(local_rank)
    p = synth.SynthParams(W, H)  # DENSE scene
    # frame-wise shard: rank r owns frames [r*n, (r+1)*n) of the global batch (no data-path collective)
    first, _last = shard.weak_shard(rank, n)
    frames = torch.empty((n, H, W, 4), dtype=torch.uint8, device="cuda")
    synth.frames_device(vision, p, first, n, frames.data_ptr())
    torch.cuda.synchronize()
    stream = torch.cuda.current_stream().cuda_stream
    fp = frames.data_ptr()

    # Batches in flight: consecutive steps alternate between `--in-flight` handles, each with its own output
    # planes and its own CUDA stream (double buffering): the line search of a batch is latency-bound and leaves issue
    # slots idle that the next batch's HBM-bound pixel pass fills.  Every step is still one smh_b200_process_batch
    # call over the whole batch, all K steps complete inside the timed region, and a step's results stay valid until
    # its handle's next call.
    n_fly = 1 if args.serial else max(1, args.in_flight)
    handles = [vision] + [B200Vision.init(local_rank) for _ in range(n_fly - 1)]
    fly_streams = [torch.cuda.Stream() for _ in range(n_fly)] if n_fly > 1 else [torch.cuda.current_stream()]
    step_no = [0]
    last = [None] * n_fly

    def step():
        k = step_no[0] % n_fly
        step_no[0] += 1
        last[k] = handles[k].process_batch(fp, n, W, H, max_gap=MAX_GAP, grayscale=True, serial=args.serial,
                                           stream=fly_streams[k].cuda_stream)

    def fork():
        if n_fly > 1:
            ev = torch.cuda.Event(); ev.record()
            for s_ in fly_streams:
                s_.wait_event(ev)

    def join():
        if n_fly > 1:
            for s_ in fly_streams:
                torch.cuda.current_stream().wait_stream(s_)

    def step_split(ev):
        # per-kernel durations: the pixel pass (k_pixels2 + k_px_seams) and the line search over the whole batch, each timed alone
        vision.process_batch(fp, n, W, H, max_gap=MAX_GAP, grayscale=True, skip_lsd=True, stream=stream)
        ev[1].record()
        vision.process_batch(fp, n, W, H, max_gap=MAX_GAP, grayscale=True, skip_pixels=True, stream=stream)

    for _ in range(max(args.warmup, 3)):
        step()
    join()
    barrier()
    # one nvidia-smi poller for the whole job (rank 0's GPU): eight of them querying the driver every 100 ms
    # perturb the launches of all ranks
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = sum(h_.launch_count() for h_ in handles)
    barrier()
    t_start = torch.cuda.Event(enable_timing=True); t_end = torch.cuda.Event(enable_timing=True)
    t_start.record()
    fork()
    for k in range(args.steps):
        step()
    join()
    t_end.record()
    barrier()
    launches = sum(h_.launch_count() for h_ in handles) - launches0

    def dev_bytes(ptr, nbytes):
        class _H:
            pass
        hd = _H()
        hd.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (int(ptr), False), "version": 2}
        return torch.as_tensor(hd, device="cuda").cpu().numpy()
    # every handle processed the same frames: their results must agree (guards the in-flight scheme)
    sums = [int(dev_bytes(b_.n_lines, n * 4).view(np.uint32).sum()) for b_ in last if b_ is not None]
    if len(set(sums)) != 1:
        raise SystemExit(f"bench: handles disagree on the number of lines found: {sums}")
    lines_checksum = sums[0]
    gpu_n_lines = dev_bytes(last[0].n_lines, n * 4).view(np.uint32).copy()
    gpu_lines = dev_bytes(last[0].lines, n * 512).view(np.float32).reshape(n, 32, 4).copy()
    total_ms = t_start.elapsed_time(t_end)
    # second timed pass, same inputs, kernels one after the other with events between them
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(3)] for _ in range(args.steps)]
    step_split(evs[0])
    barrier()
    for k in range(args.steps):
        evs[k][0].record()
        step_split(evs[k])
        evs[k][2].record()
    barrier()
    clocks = sampler.stop()
    px_ms = statistics.mean(e[0].elapsed_time(e[1]) for e in evs)
    lsd_ms = statistics.mean(e[1].elapsed_time(e[2]) for e in evs)
    per_rank = gather(world, dist, torch, [total_ms / args.steps, px_ms, lsd_ms])
    ms_per_step = max(r[0] for r in per_rank)  # MAX over ranks
    px_ms_max, lsd_ms_max = max(r[1] for r in per_rank), max(r[2] for r in per_rank)
    value = world * n / (ms_per_step * 1e-3)

    # LSD work figures (one extra, untimed, instrumented launch)
    b = vision.process_batch(fp, n, W, H, max_gap=MAX_GAP, grayscale=True, lsd_stats=True, skip_pixels=True, stream=stream)
    torch.cuda.synchronize()
    stats = dev_bytes(b.lsd_stats, n * 24).view(np.uint64).reshape(n, 3)
    seeds_pf, steps_pf, pruned_pf = float(stats[:, 0].mean()), float(stats[:, 1].mean()), float(stats[:, 2].mean())

    peaks_path = ROOT / "MEASURED_PEAKS.json"
    if peaks_path.exists():
        peak, peak_src = float(json.loads(peaks_path.read_text())["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
    else:
        peak, peak_src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"

    # ---- the CPU oracle on the first frames: the baseline AND the check of this run's GPU output ----------------
    cpu, check, oracle_lines = None, None, []
    if rank == 0 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        sample_frames = frames[:32].cpu().numpy()
        fps_cpu, oracle_lines = cpu_oracle(sample_frames, cores, 15.0)
        cpu = {"value": fps_cpu, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{len(oracle_lines)} frames of the same workload, C oracle (restated vision-cpu) with {cores} OpenMP threads"}
        bad = [i for i, ol in enumerate(oracle_lines)
               if int(gpu_n_lines[i]) != len(ol) or not np.array_equal(gpu_lines[i, :len(ol)].view(np.uint32), ol.view(np.uint32))]
        check = {"frames": len(oracle_lines), "lines": int(sum(len(ol) for ol in oracle_lines)), "mismatching_frames": bad,
                 "what": "n_lines and every f32 end point of the timed steps' output, bit for bit against the CPU oracle"}
        if bad:
            raise SystemExit(f"bench: GPU lines differ from the oracle on frames {bad}")

    # ---- end to end, single-frame latency, config 4 ---------------------------------------------------------------
    e2e = latency = config4 = None
    if not args.no_e2e:
        e2e = run_e2e(args, vision, frames, roi, n, world, barrier, torch, dist, oracle_lines)
        if rank == 0:
            latency = run_latency(vision, synth, torch)
        barrier()
        del frames
        for h_ in handles[1:]:
            h_.close()
        handles = handles[:1]
        torch.cuda.empty_cache()
        config4 = run_config4(args, vision, synth, shard, torch, dist, rank, world, barrier, roi_of, peak)

    if rank == 0:
        achieved = bpf * n / (px_ms * 1e-3) / 1e9
        # DRAM bytes of the pixel pass from the committed `ncu --set full` capture (per frame; scaled to this launch)
        traffic, traffic_src = None, None
        tpath = ROOT / "profiles" / "k_pixels_traffic.json"
        if tpath.exists():
            tj = json.loads(tpath.read_text())
            traffic, traffic_src = tj["dram_bytes_per_frame_1440p"] * n, tj["source"]
        dominant = "k_lsd2" if lsd_ms > px_ms else "k_pixels2"
        # k_lsd2 is bound by instruction issue / latency, not by HBM or the tensor pipes: its roofline is the
        # SMs' issue rate (4 schedulers x 1 warp instruction per clock per SM).  Instructions per frame come from
        # the committed ncu capture of the same workload, the duration is this run's.
        lsd_issue = None
        ipath = ROOT / "profiles" / "k_lsd_instructions.json"
        if ipath.exists():
            ij = json.loads(ipath.read_text())
            props = torch.cuda.get_device_properties(local_rank)
            sm_clock_ghz = (clocks.get("sm_mhz") or 1965.0) / 1e3
            peak_ginst = props.multi_processor_count * 4 * sm_clock_ghz
            ach_ginst = ij["warp_instructions_per_frame_1440p_dense"] * n / (lsd_ms * 1e-3) / 1e9
            lsd_issue = {"bound": "issue", "achieved": ach_ginst, "peak": peak_ginst, "unit": "G warp-inst/s",
                         "frac": ach_ginst / peak_ginst, "instructions_source": ij["source"],
                         "note": "peak = SMs x 4 schedulers x SM clock under load"}
        line = {
            "impl": "ours",
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8+f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "frames_per_gpu": n, "resolution": [W, H], "max_gap": MAX_GAP,
                       "grayscale": True, "synth": p.as_dict(),
                       "l2": f"inputs {n * W * H * 4 / 1e9:.1f} GB per GPU >> 126 MB L2, no flush needed",
                       "parallelism": f"frame-sharded x{world}, no collectives",
                       "batches_in_flight": n_fly},
            "roofline": {"kernel": "k_pixels2 (one launch: plain and OCR-quadrant CTAs interleaved) + k_px_seams", "bound": "hbm", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": bpf * n, "kernel_ms": px_ms,
                         "note": "of measured copy bandwidth; HBM-bound pixel pass over the whole batch timed alone "
                                 "with CUDA events (second timed pass). k_lsd2 is not HBM-bound (see kernels)"},
            "roofline_step": {"bound": "hbm", "achieved": bpf * n / (ms_per_step * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                              "frac": bpf * n / (ms_per_step * 1e-3) / 1e9 / peak,
                              "note": "algorithmic bytes of one step / ms_per_step: the whole step (pixel pass + line search) against the HBM roofline"},
            "kernels": {"dominant_by_time": dominant,
                        "timing": "second timed pass of the same steps: the pixel pass, then the line search, over the whole batch, events between",
                        "serial_ms_per_step": px_ms + lsd_ms,
                        "step_form": "one smh_b200_process_batch call per step; "
                                     + (f"{n_fly} batches in flight: steps alternate between {n_fly} handles, each with its own planes and CUDA stream"
                                        if n_fly > 1 else "one handle, one stream"),
                        "in_flight": n_fly,
                        "overlap_speedup": (px_ms + lsd_ms) / (total_ms / args.steps),
                        "k_pixels2": {"ms": px_ms, "share": px_ms / (px_ms + lsd_ms)},
                        "k_lsd2": {"ms": lsd_ms, "share": lsd_ms / (px_ms + lsd_ms), "bound": "instruction issue / latency of the per-frame seed chain",
                                   "seeds_per_frame": seeds_pf, "seeds_pruned_per_frame": pruned_pf, "ray_steps_per_frame": steps_pf,
                                   "ray_steps_per_s": steps_pf * n / (lsd_ms * 1e-3), "issue_roofline": lsd_issue}},
            "per_rank": {"ms_per_step": [r[0] for r in per_rank], "pixel_pass_ms": [r[1] for r in per_rank], "line_search_ms": [r[2] for r in per_rank],
                         "max_ms": {"step": ms_per_step, "pixel_pass": px_ms_max, "line_search": lsd_ms_max},
                         "note": "every rank has different frames (global frame index seeds the generator): the line search is data-dependent, the pixel pass is not"},
            "frames_per_s_pixels_only": n / (px_ms * 1e-3),
            "lines_checksum": lines_checksum,
            "check": check,
            "clocks": clocks,
            "e2e": e2e,
            "latency": latency,
            "config4": config4,
            "gpu_launches": int(launches),
            "cpu_baseline": cpu,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    for h_ in handles:
        h_.close()


if __name__ == "__main__":
    main()
