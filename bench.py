#!/usr/bin/env python3
"""bench.py — map frames/s of the per-frame vision path on B200 (BASELINE.json metric).

A "step" is one pass of the whole hot path (fused pixel pass + on-device line search) over one
batch of synthetic 2560x1440 captures per GPU (BASELINE config 3: dense SL markers + lines,
batch 1024).  Frames are independent, so N GPUs each process their own 1024 frames with no
collective ("weak" scaling); torch.distributed (NCCL) is only used for the barrier and the
max-over-ranks timing.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Prints ONE JSON line (rank 0).  Keys beyond the base contract:
  roofline      HBM roofline of the bandwidth-bound kernel (k_pixels), CUDA-event timed in the step
  kernels       per-kernel time share inside the timed region (k_pixels, k_lsd) + LSD work figures
  cpu_baseline  the CPU oracle (a restatement of vision-cpu) timed on this box's host cores
  e2e           same metric through the C ABI with HOST frames (H2D + D2H inside the timed region)
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


def cpu_oracle_fps(frames_np, threads: int, budget_s: float = 20.0):
    """Time the CPU oracle (oracle/ — a port of vision-cpu) on a bounded sample of the workload."""
    from oracle import oracle_py as O
    O.set_threads(threads)
    o = O.Oracle()
    o.process_frame(frames_np[0], True, 0, MAX_GAP)  # warm-up (thread pool, page faults)
    t0 = time.perf_counter()
    done = 0
    for f in frames_np:
        o.process_frame(f, True, 0, MAX_GAP)
        done += 1
        if time.perf_counter() - t0 > budget_s:
            break
    dt = time.perf_counter() - t0
    return done / dt, done


def run_reference(args, rank: int, world: int) -> None:
    """--impl reference: the reference's CPU implementation of the path.  Its Rust crate cannot be
    built here (no cargo/rustc, crates.io deps absent), so this times the oracle port with all
    host threads on a bounded sample of the same workload."""
    if rank != 0:
        return
    import smh_b200  # noqa: F401
    from squad_mortar_helper_b200 import synth
    cores = os.cpu_count() or 1
    n_sample = 8
    p = synth.SynthParams(W, H)
    frames = synth.frames_host(p, 0, n_sample)
    from oracle import oracle_py as O
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


def run_e2e(args, vision, frames, roi, n, world, barrier):
    """Same metric through smh_b200_process_host_batch: pinned HOST frames in, host results out."""
    import torch
    import torch.distributed as dist
    ne = min(args.e2e_frames or (n if world <= 2 else (512 if world <= 4 else 256)), n)
    host_frames = torch.empty((ne, H, W, 4), dtype=torch.uint8).pin_memory()
    host_frames.copy_(frames[:ne])
    _, _, mw, mh = roi["map"]; _, _, qw, qh = roi["brq"]
    outs = {"ui_map": torch.empty((ne, mh, mw, 4), dtype=torch.uint8).pin_memory(),
            "ocr_out": torch.empty((ne, qh, qw), dtype=torch.uint8).pin_memory(),
            "scales": torch.empty((ne, qh, qw), dtype=torch.uint8).pin_memory(),
            "lines": torch.empty((ne, 32, 4), dtype=torch.float32).pin_memory(),
            "n_lines": torch.empty((ne,), dtype=torch.int32).pin_memory(),
            "red_pixels": torch.empty((ne,), dtype=torch.int32).pin_memory(),
            "map_open": torch.empty((ne,), dtype=torch.uint8).pin_memory()}
    out_ptrs = {k: v.data_ptr() for k, v in outs.items()}
    torch.cuda.synchronize()

    def e2e_step():
        vision.process_host_batch(None, frames_ptr=host_frames.data_ptr(), shape=(ne, H, W), max_gap=MAX_GAP,
                                  grayscale=True, chunk=32, want=("ui_map", "ocr_out", "scales"), out=dict(out_ptrs))

    for _ in range(2):
        e2e_step()
    barrier()
    e_steps = max(3, min(args.steps, 5))
    t0 = time.perf_counter()
    for _ in range(e_steps):
        e2e_step()
    torch.cuda.synchronize()
    e_dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e_dt], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e_dt = float(t.cpu()[0])
    e2e_fps = world * ne * e_steps / e_dt
    # what smh_b200_process_host_batch uploads per frame: the map rows as wide as the pixel pass's
    # 16-byte vector grid, plus the button rectangle
    gx0 = roi["map"][0] & ~3
    nvec = (roi["map"][0] - gx0 + mw + 3) // 4
    h2d = ne * (min(4 * nvec, W - gx0) * 4 * mh + roi["button"][2] * roi["button"][3] * 4)
    # grayscale ui_map: only its L plane crosses the bus (host threads of the library widen it to RGBA)
    ui_l8 = os.environ.get("SMH_B200_UI_L8", "1") != "0"
    d2h = ne * (mw * mh * (1 if ui_l8 else 4) + 2 * qw * qh + 512 + 9)
    e2e_check = int(outs["n_lines"].sum())

    return {"value": e2e_fps, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
            "frames_per_step": ne, "steps": e_steps, "lines_checksum": e2e_check,
            "note": "smh_b200_process_host_batch: pinned host frames -> map + button rectangles up -> kernels -> ui_map/ocr/scales/lines back, "
                    "3 chunks in flight; the grayscale ui_map crosses the bus as its L plane and is widened to RGBA in the caller's buffer "
                    "by the library's host threads inside the timed region" if ui_l8 else
                    "smh_b200_process_host_batch: pinned host frames -> map + button rectangles up -> kernels -> ui_map/ocr/scales/lines back, 3 chunks in flight"}


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=1024, help="frames per GPU per step")
    ap.add_argument("--e2e-frames", type=int, default=0,
                    help="host frames per e2e step (default: the whole batch; 512 per rank beyond 2 GPUs, 256 beyond 4, to bound pinned host memory)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="profiling runs: skip the end-to-end leg")
    ap.add_argument("--in-flight", type=int, default=2,
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
    # planes and its own CUDA stream (double buffering): the line search of a batch is a couple of waves of
    # long sequential per-frame chains, and its tail leaves most SMs idle — the next batch's pixel pass and
    # line search fill them.  Every step is still one smh_b200_process_batch call over the whole batch, all
    # K steps complete inside the timed region, and a step's results stay valid until its handle's next call.
    n_fly = 1 if args.serial else max(1, args.in_flight)
    handles = [vision] + [B200Vision.init(local_rank) for _ in range(n_fly - 1)]
    fly_streams = [torch.cuda.Stream() for _ in range(n_fly)] if n_fly > 1 else [torch.cuda.current_stream()]
    step_no = [0]
    last = [None] * n_fly

    def step():
        # the product call: one smh_b200_process_batch over the whole batch; --serial forces one handle,
        # one stream (what the ncu launch list is taken on)
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
        # per-kernel durations: one launch of each kernel over the whole batch, each timed alone
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

    def dev_u32(ptr, count):
        class _H:
            pass
        hd = _H()
        hd.__cuda_array_interface__ = {"shape": (count * 4,), "typestr": "|u1", "data": (int(ptr), False), "version": 2}
        return torch.as_tensor(hd, device="cuda").cpu().numpy().view(np.uint32)
    # every handle processed the same frames: their results must agree (guards the in-flight scheme)
    sums = [int(dev_u32(b_.n_lines, n).sum()) for b_ in last if b_ is not None]
    if len(set(sums)) != 1:
        raise SystemExit(f"bench: handles disagree on the number of lines found: {sums}")
    lines_checksum = sums[0]
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
    if world > 1:
        t = torch.tensor([total_ms, px_ms, lsd_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, px_ms, lsd_ms = (float(x) for x in t.cpu())
    ms_per_step = total_ms / args.steps
    value = world * n / (ms_per_step * 1e-3)

    # LSD work figures (one extra, untimed, instrumented launch)
    b = vision.process_batch(fp, n, W, H, max_gap=MAX_GAP, grayscale=True, lsd_stats=True, skip_pixels=True, stream=stream)
    torch.cuda.synchronize()

    class _H:
        pass
    hld = _H()
    hld.__cuda_array_interface__ = {"shape": (n * 24,), "typestr": "|u1", "data": (int(b.lsd_stats), False), "version": 2}
    stats = torch.as_tensor(hld, device="cuda").cpu().numpy().view(np.uint64).reshape(n, 3)
    seeds_pf, steps_pf, pruned_pf = float(stats[:, 0].mean()), float(stats[:, 1].mean()), float(stats[:, 2].mean())

    # ---- e2e: host frames through the C ABI, H2D + D2H inside the timed region ------------
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(args, vision, frames, roi, n, world, barrier)

    if rank == 0:
        peaks_path = ROOT / "MEASURED_PEAKS.json"
        if peaks_path.exists():
            peak, peak_src = float(json.loads(peaks_path.read_text())["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
        else:
            peak, peak_src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"
        achieved = bpf * n / (px_ms * 1e-3) / 1e9
        # DRAM bytes of k_pixels from the committed `ncu --set full` capture (per frame; scaled to this launch)
        traffic, traffic_src = None, None
        tpath = ROOT / "profiles" / "k_pixels_traffic.json"
        if tpath.exists():
            tj = json.loads(tpath.read_text())
            traffic, traffic_src = tj["dram_bytes_per_frame_1440p"] * n, tj["source"]
        cpu = None
        if not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            sample_frames = frames[:32].cpu().numpy()
            fps_cpu, done = cpu_oracle_fps(sample_frames, cores, 15.0)
            cpu = {"value": fps_cpu, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": f"{done} frames of the same workload, C oracle (restated vision-cpu) with {cores} OpenMP threads"}
        dominant = "k_lsd" if lsd_ms > px_ms else "k_pixels"
        # k_lsd is bound by instruction issue / latency, not by HBM or the tensor pipes: its roofline is the
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
                         "note": "duration includes k_pack_bits (~8 % of it); peak = SMs x 4 schedulers x SM clock under load"}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8+f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "frames_per_gpu": n, "resolution": [W, H], "max_gap": MAX_GAP,
                       "grayscale": True, "synth": p.as_dict(),
                       "l2": f"inputs {n * W * H * 4 / 1e9:.1f} GB per GPU >> 126 MB L2, no flush needed",
                       "parallelism": f"frame-sharded x{world}, no collectives",
                       "batches_in_flight": n_fly},
            "roofline": {"kernel": "k_pixels", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": bpf * n, "kernel_ms": px_ms,
                         "note": "of measured copy bandwidth; HBM-bound pixel pass, one launch over the whole batch timed alone "
                                 "with CUDA events (second timed pass). k_lsd is not HBM-bound (see kernels)"},
            "kernels": {"dominant_by_time": dominant,
                        "timing": "second timed pass of the same steps: one launch per kernel over the whole batch, events between",
                        "serial_ms_per_step": px_ms + lsd_ms,
                        "step_form": "one smh_b200_process_batch call per step (one launch per kernel); "
                                     + (f"{n_fly} batches in flight: steps alternate between {n_fly} handles, each with its own planes and CUDA stream"
                                        if n_fly > 1 else "one handle, one stream"),
                        "in_flight": n_fly,
                        "overlap_speedup": (px_ms + lsd_ms) / ms_per_step,
                        "k_pixels": {"ms": px_ms, "share": px_ms / (px_ms + lsd_ms)},
                        "k_lsd": {"ms": lsd_ms, "share": lsd_ms / (px_ms + lsd_ms), "bound": "instruction issue / L1 latency",
                                  "seeds_per_frame": seeds_pf, "seeds_pruned_per_frame": pruned_pf, "ray_steps_per_frame": steps_pf,
                                  "ray_steps_per_s": steps_pf * n / (lsd_ms * 1e-3), "issue_roofline": lsd_issue}},
            "frames_per_s_pixels_only": n / (px_ms * 1e-3),
            "lines_checksum": lines_checksum,
            "clocks": clocks,
            "e2e": e2e,
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
