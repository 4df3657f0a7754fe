#!/usr/bin/env python
"""bench.py -- FeatureCorrelation DB frame-offsets/s on B200 (BASELINE.json config 5).

  python bench.py --gpus 1 --steps K --warmup W            own arm (CUDA path through the C ABI)
  torchrun ... bench.py --gpus N --steps K --warmup W      one rank per GPU, weak scaling
  python bench.py --impl reference ...                     the reference's CPU algorithm (oracle port)

A step is ONE complete search (K1 scan + K2 selection + ordered merge) of the 2 s punch-in window over
the database resident in HBM: 6000 synthetic feature files x 51 680 frames (1000 h, 13 MFCC + loudness,
fft 1024 / step 512 @ 44.1 kHz) PER GPU, numMatches 100, numPerFile 1, minSpacing 0.5 s.  `value` is
evaluated frame-offsets of all ranks / max-over-ranks device time.  `e2e` is the same search through the
same C-ABI calls starting from HOST buffers (pinned): DB upload + normalise/transpose + search + result
download inside the timed region.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FRAMES_PER_FILE = 51680
STEP = 512
W_IN_SAMPLES = 88200
FLOP_PER_OFFSET = 2 * 14 * 172 + 4 * 14 + 32   # 4904
BYTES_PER_OFFSET = 56 + 8                      # 56 B DB frame read + (sim, boost) curve written


def corr_config(N, num_matches=100, num_per_file=1, min_spacing=22050):
    return N.CorrConfig(STEP, 0, W_IN_SAMPLES, 0.5, 0, 0, 0, 0.5, 44100, 352800, 8.0, num_matches, num_per_file,
                        min_spacing)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe).  The process is started
    before the warm-up (its start-up alone is longer than a short timed region), every row is stamped on arrival, and
    only rows that arrived between mark_begin() and mark_end() count."""

    def __init__(self, index: int):
        self.index = index
        self.rows = []      # (arrival time, fields)
        self.proc = None
        self.t0 = self.t1 = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap,clocks_event_reasons.active")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [x.strip() for x in line.split(",")]))

    def mark_begin(self):
        self.t0 = time.perf_counter()

    def mark_end(self):
        self.t1 = time.perf_counter()

    def in_window(self):
        lo = self.t0 if self.t0 is not None else -1e300
        hi = self.t1 if self.t1 is not None else 1e300
        return [r for (t, r) in list(self.rows) if lo <= t <= hi and len(r) >= 7]

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        rows = self.in_window()
        sm = [float(r[0]) for r in rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in rows if r[1].replace(".", "").isdigit()]
        pw = [float(r[2]) for r in rows if r[2].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(r[3 + k].lower() == "active" for r in rows)]
        mask = 0
        for r in rows:
            try:
                mask |= int(r[7], 16) if len(r) > 7 else 0
            except ValueError:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_min_mhz": min(sm) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "power_w_max": max(pw) if pw else None, "reasons": reasons,
                "reasons_mask": hex(mask), "samples": len(sm)}


def host_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def reference_arm(args):
    """The reference's own CPU algorithm (oracle port: no JVM exists on the box), all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from concurrent.futures import ThreadPoolExecutor
    from oracle import oracle as O
    from strugatzki_b200 import synth
    cores = host_threads()
    mu, sigma, floor0, norm = synth.default_profile(14)
    files_per_thread = 2
    files = [synth.synth_file(synth.BASE_SEED, 1 + i, FRAMES_PER_FILE, mu, sigma, floor0)
             for i in range(files_per_thread)]
    inp = synth.synth_file(synth.BASE_SEED, 0, 900, mu, sigma, floor0)
    p = O.CorrParams(step_size=STEP, input=inp, punch_in=(0, W_IN_SAMPLES), norm=norm, num_matches=100,
                     num_per_file=1, min_spacing=22050)
    offsets_per_thread = O.corr_num_offsets(p, [f.shape[0] for f in files])
    O.lib()

    def one(_):
        return O.corr_search(O.CorrParams(**{k: getattr(p, k) for k in (
            "step_size", "input", "punch_in", "punch_in_weight", "punch_out", "punch_out_weight", "min_punch",
            "max_punch", "norm", "max_boost", "num_matches", "num_per_file", "min_spacing")}), files)

    times = []
    with ThreadPoolExecutor(cores) as ex:
        for it in range(args.warmup + args.steps):
            t = time.perf_counter()
            list(ex.map(one, range(cores)))
            dt = time.perf_counter() - t
            if it >= args.warmup:
                times.append(dt)
    total = offsets_per_thread * cores
    ms = 1e3 * float(np.mean(times))
    value = total / (ms * 1e-3)
    sample = (f"{cores} threads x {files_per_thread} files x {FRAMES_PER_FILE} frames "
              f"({total} offsets per step); data pre-loaded in RAM (the reference's 1-frame AudioFile.read per "
              f"offset is excluded)")
    line = {"impl": "reference", "metric": "FeatureCorrelation DB frame-offsets/sec", "value": value,
            "unit": "offsets/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": config_block(args.gpus), "config_note": "bounded sample of the same workload",
            "cpu_baseline": {"value": value, "unit": "offsets/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "offsets/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)
    return 0


def config_block(n_gpus, files=6000):
    c = {"workload": "FeatureCorrelation punch-in 2 s (W=172 frames) over 1000 h synthetic feature DB per GPU "
                     "(BASELINE.json configs[4])",
         "files_per_gpu": files, "frames_per_file": FRAMES_PER_FILE, "channels": 14, "temporalWeight": 0.5,
         "numMatches": 100, "numPerFile": 1, "minSpacing": 22050, "sharding": f"file-range x{n_gpus}",
         "cache": "DB per GPU (17.4 GB) >> 126 MB L2, no reuse between steps"}
    return c


_JSON_FD = None


def emit(line: dict) -> None:
    """The ONE JSON line of this run, on the real stdout (see main: libraries that print to fd 1 -- NCCL announces its
    version there -- are pointed at stderr for the whole run)."""
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)              # anything else written to stdout goes to stderr
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--files", type=int, default=6000, help="DB files per GPU (default = 1000 h)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "native" else args.warmup

    if args.impl == "reference":
        return reference_arm(args)

    import torch
    import torch.distributed as dist
    from strugatzki_b200 import _native as N, engine, synth
    from strugatzki_b200.distributed import sharded_search

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=device)

    ctx = engine.Context(local_rank)
    ext_stream = torch.cuda.ExternalStream(ctx.stream, device=device)

    # ---- live roofline denominators on this GPU ----
    ffma_peak = ctx.measure_peak(0)
    peaks_file = {}
    try:
        peaks_file = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks_file.get("hbm_gbs", 6650.0))
    hbm_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks_file else "fallback (B200_PROFILING.md)"

    # ---- database: generated on the device, then the normal prepare path ----
    mu, sigma, floor0, norm = synth.default_profile(14)
    files = args.files
    db = engine.Database(ctx, 14, norm)
    db.reserve(files * FRAMES_PER_FILE, files)
    db.add_synth_many(synth.BASE_SEED, 1 + rank * files, files, FRAMES_PER_FILE, mu, sigma, float(floor0))   # one launch
    inp = synth.synth_file(synth.BASE_SEED, 0, 900, mu, sigma, floor0)
    rng = np.random.default_rng(1234 + rank)
    needles = []
    for k in range(8):  # planted noisy copies of the query at known places
        f, off = int(rng.integers(0, files)), int(rng.integers(0, FRAMES_PER_FILE - 172))
        db.patch(f, off, synth.plant(inp[:172], 77 + rank, k))
        needles.append((rank * files + f, off * STEP))
    db.finalize()
    cfg = corr_config(N)
    job = engine.CorrelationJob(db, cfg, inp)
    n_off_local = job.num_offsets

    phases = {}

    def step():
        return sharded_search(job, device, None, trace=phases) if world > 1 else job.run()

    def sync_all():
        ctx.synchronize()
        torch.cuda.synchronize(device)
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize(device)

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()     # before the warm-up: nvidia-smi needs longer to start than a short timed region lasts
    for _ in range(args.warmup):
        res = step()
    sync_all()
    sampler.mark_begin()
    launches0 = ctx.launch_count
    phases.clear()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    scan_ms, select_ms = [], []
    t_wall = time.perf_counter()
    ev0.record(ext_stream)
    for _ in range(args.steps):
        res = step()
        tm = job.timing()
        scan_ms.append(tm["scan_ms"])
        select_ms.append(tm["select_ms"])
    ev1.record(ext_stream)
    sync_all()
    wall_ms = 1e3 * (time.perf_counter() - t_wall)
    dev_ms = ev0.elapsed_time(ev1)
    launches = ctx.launch_count - launches0
    phases = dict(phases)   # the extra steps below must not count into the per-step phase times
    # a timed region of K x 8 ms can fall between two nvidia-smi rows: the same steps keep running, untimed, until at
    # least five rows have been taken under this load (every rank takes part: the steps are collective at N > 1)
    extra = 0
    for _ in range(40):
        flag = torch.tensor([1.0 if (rank == 0 and len(sampler.in_window()) < 5) else 0.0], device=device)
        if world > 1:
            dist.broadcast(flag, 0)
        if float(flag[0]) == 0.0:
            break
        for _ in range(max(args.steps, 10)):
            sharded_search(job, device, None) if world > 1 else job.run()
        sync_all()
        extra += max(args.steps, 10)
    sampler.mark_end()
    clocks = sampler.stop() if rank == 0 else None
    if clocks is not None:
        clocks["extra_load_steps"] = extra   # identical untimed steps run behind the timed ones while sampling

    t = torch.tensor([dev_ms, wall_ms, float(n_off_local), float(np.mean(scan_ms))], dtype=torch.float64, device=device)
    if world > 1:
        mx = t.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = t.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        dev_ms_max, total_off, scan_ms_max = float(mx[0]), float(sm[2]), float(mx[3])
    else:
        dev_ms_max, total_off, scan_ms_max = dev_ms, float(n_off_local), float(np.mean(scan_ms))
    ms_per_step = dev_ms_max / args.steps
    value = total_off / (ms_per_step * 1e-3)

    # sanity inside the bench: every planted needle of this rank must be found where it was planted
    found = {(m["file"], m["start"]) for m in res}
    missing = [nd for nd in needles if nd not in found] if world == 1 else []

    # ---- end to end from HOST buffers through the same C-ABI calls ----
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(args, torch, device, ctx, engine, N, synth, rank, files, mu, sigma, floor0, norm, inp, cfg, world,
                      dist)

    # ---- secondary metric of BASELINE.json: SelfSimilarity cells/s (+ segmentation), bounded workloads ----
    secondary = None
    if rank == 0 and world == 1:
        secondary = run_secondary(ctx, engine, N, synth, norm)

    if world > 1:
        secondary = run_secondary_sharded(torch, dist, device, ctx, engine, N, synth, norm, rank, world)

    # ---- CPU baseline: the oracle port, one host thread (the reference is single-threaded) ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_baseline(args, synth, norm, inp)

    if rank == 0:
        scan_s = float(np.mean(scan_ms)) * 1e-3
        flops = n_off_local * FLOP_PER_OFFSET / scan_s / 1e12
        tc = os.environ.get("SGZ_CORR_TC", "1") != "0"     # K1 on the tensor cores (default) or the FFMA2 kernel
        traffic = None   # dram__bytes_read + dram__bytes_write of one K1 launch, from the committed ncu capture
        traffic_file = ("r02_k_corr_tc2_traffic.json" if os.environ.get("SGZ_CORR_TC2", "1") != "0" else
                        "r01_k_corr_tc_traffic.json") if tc else "r01_k_corr_traffic.json"
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", traffic_file)))
            if int(tj["offsets_per_launch"]) == int(n_off_local):
                traffic = float(tj["dram_bytes_per_launch"])
        except Exception:
            pass
        hbm = {"achieved": n_off_local * BYTES_PER_OFFSET / scan_s / 1e9, "peak": hbm_peak, "unit": "GB/s",
               "frac": n_off_local * BYTES_PER_OFFSET / scan_s / 1e9 / hbm_peak, "peak_source": hbm_src,
               "algorithmic_bytes_per_offset": BYTES_PER_OFFSET}
        if tc:
            # tensor-core K1 (corr_tc2.cuh): per tile of 8192 offsets, 14 channels x 3 split-FP16 products x
            # ceil((63 + W) / 16) MMAs of M128 x N64 x K16.  An MMA reads its operands from shared memory: 6 KB (A 4 KB +
            # B 2 KB), 2 KB when A comes from the operand collector -> 14 KB per K step of three MMAs, 109 B per clock at the
            # MMA rate the tensor pipe could sustain.  Together with the bulk copies that fill the ring (731 KB per tile)
            # and the epilogue's loads / stores the SM's 128 B/clock shared-memory / L1 data pipe is the binding resource
            # of this formulation ("smem_floor_ms"); see DESIGN.md.
            tensor_peak = float(peaks_file.get("bf16_tflops_sustained", 1399.0))
            t2 = os.environ.get("SGZ_CORR_TC2", "1") != "0"
            if t2:
                ks, tile, ncols, kname = (63 + 172 + 15) // 16, 8192, 64, "sgz::k_corr_tc2"
                mma_cycles = (48 + 32 + 48) / 3.0
            else:
                ks, tile, ncols, kname = (32 + 172 + 15) // 16, 4096, 32, "sgz::k_corr_tc"
                mma_cycles = 44.5
            mmas_per_tile = 14 * 3 * ks
            exec_flop_per_offset = mmas_per_tile * 2 * 128 * ncols * 16 / tile
            tiles_per_sm = -(-n_off_local // tile) / 148.0
            floor_ms = tiles_per_sm * mmas_per_tile * mma_cycles / 1.965e6
            smem_bytes_per_tile = 14 * ks * 14336 + 14 * (2 * 16768 + 18432) + 4096 * 128 + 3000 * 128 if t2 else None
            # Which roof binds?  Algorithmic intensity = 4 904 flop / 64 B = 77 flop/B; with the tensor cores as the math
            # roof the ridge is at tensor_peak / hbm_peak (= 214 flop/B with the measured 1 399 TFLOP/s and 6 549 GB/s):
            # the kernel sits LEFT of the ridge, so the roofline model bounds it by HBM (min(peak, AI x BW) = 502
            # TFLOP/s-equivalent).  "frac" is therefore the HBM fraction; the tensor-pipe view is kept next to it.
            ridge = tensor_peak * 1e3 / hbm_peak
            intensity = FLOP_PER_OFFSET / BYTES_PER_OFFSET
            tensor = {"achieved": flops, "peak": tensor_peak, "unit": "TFLOP/s", "frac": flops / tensor_peak,
                      "peak_source": "dense bf16 (= fp16) matmul, sustained, MEASURED_PEAKS.json",
                      "executed_tflops": n_off_local * exec_flop_per_offset / scan_s / 1e12,
                      "executed_frac": n_off_local * exec_flop_per_offset / scan_s / 1e12 / tensor_peak,
                      "executed_flop_per_offset": exec_flop_per_offset,
                      "mma_floor_ms": floor_ms, "frac_of_mma_floor": floor_ms / float(np.mean(scan_ms)),
                      "fp32_ffma_peak": ffma_peak, "algorithmic_vs_fp32_ffma_peak": flops / ffma_peak}
            if smem_bytes_per_tile:
                tensor["smem_floor_ms"] = tiles_per_sm * smem_bytes_per_tile / 128.0 / 1.965e6
                tensor["frac_of_smem_floor"] = tensor["smem_floor_ms"] / float(np.mean(scan_ms))
            roofline = {"kernel": kname + " (K1 sliding-window Pearson correlation on tcgen05: split-FP16 operands from "
                                          "precomputed planes by bulk copy, N = 64 Hankel tiles, fused window statistics / sim / "
                                          "file maxima)" if t2 else kname + " (round-1 N = 32 tensor-core kernel)",
                        "bound": "hbm" if intensity < ridge else "tensor",
                        "achieved": hbm["achieved"] if intensity < ridge else flops,
                        "peak": hbm_peak if intensity < ridge else tensor_peak,
                        "unit": "GB/s" if intensity < ridge else "TFLOP/s",
                        "frac": hbm["frac"] if intensity < ridge else flops / tensor_peak,
                        "peak_source": hbm_src if intensity < ridge else tensor["peak_source"],
                        "intensity_flop_per_byte": intensity, "ridge_flop_per_byte": ridge,
                        "why_not_higher": "power: with the tensor cores, HBM and 16 epilogue warps busy at once the GPU sits at its "
                                          "1000 W cap and the SM clock averages 1.48 GHz INSIDE the kernel (clock64 / %globaltimer per "
                                          "CTA; 1.96 GHz with the MMAs ablated; nvidia-smi's samples between launches read higher) -- "
                                          "a tile of 8192 offsets takes 40.8 k cycles (20.8 us at 1.965 GHz = frac 0.57, 27.6 us at the "
                                          "capped clock).  In cycles: three split-FP16 products for FP32-grade precision = 630 MMAs per "
                                          "tile, 28.9 k cycles of MMA issue (an M128xN64xK16 MMA fetches 6 KB of operands from shared "
                                          "memory at 128 B/clock; the band of the taps matrix already runs as N = 16/32/48 MMAs at its "
                                          "edges), 7 k cycles waiting for operands / accumulators, the epilogue (25 k) hidden behind "
                                          "the next tile's MMAs; data movement alone takes 25 k cycles per tile at full clock "
                                          "(profiles/r02_k_corr_tc2_power_ablation.txt)",
                        "tensor": tensor}
        else:
            roofline = {"kernel": "sgz::k_corr (K1 sliding-window Pearson correlation, FFMA2)", "bound": "fp32_ffma",
                        "achieved": flops, "peak": ffma_peak, "unit": "TFLOP/s", "frac": flops / ffma_peak,
                        "peak_source": "FP32 FFMA micro-benchmark run live by this process "
                                       "(MEASURED_PEAKS.json has no FP32 figure; SURVEY.md 8d)"}
        roofline.update({"algorithmic_flop_per_offset": FLOP_PER_OFFSET, "offsets_per_launch": n_off_local,
                         "launch_ms": float(np.mean(scan_ms)), "traffic": traffic,
                         "traffic_source": f"profiles/{traffic_file} (ncu --set full of this command)",
                         "algorithmic_bytes_per_launch": n_off_local * BYTES_PER_OFFSET, "hbm": hbm})
        line = {
            "metric": "FeatureCorrelation DB frame-offsets/sec", "value": value, "unit": "offsets/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16x2 split operands, f32 accumulate, window statistics f32 centred on exact f64 sums",
            "data": "synthetic (device-generated integer-hash features, planted needles)",
            "config": config_block(world, files),
            "roofline": roofline,
            "breakdown_ms": {"k1_scan": float(np.mean(scan_ms)), "k2_select_kernels": float(np.mean(select_ms)),
                             "note": "k1_scan = the k_corr_tc2 launch alone (an event sits right behind it); k2_select_kernels = "
                                     "what follows it on the device: exact Double re-evaluation of the offsets that can reach "
                                     "the result (the returned sims are the reference's bit for bit), per-file boosts",
                             "wall_per_step": wall_ms / args.steps,
                             "rank0_host_phases": {k: (v / args.steps if k == "rounds" else 1e3 * v / args.steps)
                                                   for k, v in phases.items()} if world > 1 else None},
            "gpu_launches": int(launches), "clocks": clocks,
            "e2e": e2e, "cpu_baseline": cpu, "secondary": secondary,
            "matches": len(res), "needles_missing": len(missing), "top_sim": res[0]["sim"] if res else None,
        }
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def run_e2e(args, torch, device, ctx, engine, N, synth, rank, files, mu, sigma, floor0, norm, inp, cfg, world, dist):
    import psutil
    bytes_per_file = FRAMES_PER_FILE * 14 * 4
    avail = psutil.virtual_memory().available
    e2e_files = files
    while e2e_files * bytes_per_file * 1.3 * max(world, 1) > avail * 0.6 and e2e_files > 60:
        e2e_files //= 2
    # raw (un-normalised) features of this rank's first e2e_files files -> pinned host memory, planar per file
    host = torch.empty((e2e_files, 14, FRAMES_PER_FILE), dtype=torch.float32, pin_memory=True)
    raw = engine.Database(ctx, 14, None)
    chunk = 200
    hnp = host.numpy()
    for c0 in range(0, e2e_files, chunk):
        raw = engine.Database(ctx, 14, None)
        raw.add_synth_many(synth.BASE_SEED, 1 + rank * files + c0, min(chunk, e2e_files - c0), FRAMES_PER_FILE, mu, sigma,
                           float(floor0))
        raw.finalize()
        for i in range(c0, min(c0 + chunk, e2e_files)):
            hnp[i] = raw.read(i - c0, 0, FRAMES_PER_FILE)
        raw.close()
    base = host.data_ptr()

    def e2e_step():
        db2 = engine.Database(ctx, 14, norm)
        db2.reserve(e2e_files * FRAMES_PER_FILE, e2e_files)
        for i in range(e2e_files):
            db2.add_file_ptr(base + i * bytes_per_file, FRAMES_PER_FILE, N.LAYOUT_PLANAR_LE | N.LAYOUT_HOST_STABLE)
        db2.finalize(wait=False)     # the search streams behind the uploads still in flight
        job2 = engine.CorrelationJob(db2, cfg, inp)
        if world > 1:
            from strugatzki_b200.distributed import sharded_search
            r = sharded_search(job2, device, None)
        else:
            r = job2.run()
        n = job2.num_offsets
        job2.close()
        db2.close()
        return r, n

    e2e_step()
    ctx.synchronize()
    if world > 1:
        dist.barrier()
    times = []
    n_off = 0
    for _ in range(max(args.e2e_steps, 1)):
        ctx.synchronize()
        t = time.perf_counter()
        r, n_off = e2e_step()
        ctx.synchronize()
        times.append(time.perf_counter() - t)
    t_loc = torch.tensor([float(np.mean(times)), float(n_off)], dtype=torch.float64, device=device)
    if world > 1:
        mx = t_loc.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = t_loc.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        sec, tot = float(mx[0]), float(sm[1])
    else:
        sec, tot = float(t_loc[0]), float(t_loc[1])
    del host
    return {"value": tot / sec, "unit": "offsets/s", "ms_per_step": sec * 1e3,
            "h2d_bytes_per_step": int(e2e_files * bytes_per_file + inp.nbytes + 14 * 172 * 4),
            "d2h_bytes_per_step": int(100 * 32 + e2e_files * 8 + 100 * 32),
            "files_per_gpu": e2e_files,
            "note": "DB upload from pinned host memory (planar float32) + normalise + search + result download, "
                    "through sgz_db_add_file / sgz_db_finalize_async / sgz_corr_run (K1 streams behind the upload); PCIe-bound"}


def run_secondary(ctx, engine, N, synth, norm):
    """SelfSimilarity cells/s (tensor-core Gram kernel at two sizes, the FP64 replay) and segmentation offsets/s."""
    out = {}
    try:
        try:
            tensor_peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("bf16_tflops_sustained", 1399.0))
        except Exception:
            tensor_peak = 1399.0
        cases = (("selfsimilarity_gram_tc", 30000, 1, 0, "30 000-frame synthetic feature file"),
                 ("selfsimilarity_gram_tc_config3", 155000, 1, 0,
                  "155 000-frame synthetic feature file (BASELINE.json configs[3]; decimation raised to 4 by the 0xB504 rule)"),
                 ("selfsimilarity_exact_fp64", 30000, 3, 1, "30 000-frame synthetic feature file"))
        files = {}
        for name, frames, decim, precise, what in cases:
            if frames not in files:
                files[frames] = synth.regime_file(synth.BASE_SEED, 4, frames, 14, max(4, frames // 2000))[0]
            f = files[frames]
            cfg = N.SelfConfig(STEP, 0, 0, 0, 0, 44100, decim, 0.5, 0, 1.0, 1.0, None, 0, precise)
            engine.self_run(ctx, cfg, f, None, norm, download=False)
            _, g = engine.self_run(ctx, cfg, f, None, norm, download=False)
            ms, launches = ctx.last_timing()
            cells = g["numCells"]
            flop = cells * 2408.0 / (ms * 1e-3) / 1e12      # 2*(C+1)*H flop per cell at H = 86
            out[name] = {"metric": "SelfSimilarity cells/sec", "value": cells / (ms * 1e-3), "unit": "cells/s",
                         "cells": cells, "imgExt": g["imgExt"], "decim": g["decim"], "kernel_ms": ms,
                         "kernel": engine.self_last_kernel(ctx), "algorithmic_tflops": flop,
                         "workload": what + ", corrLen 44100 (H = 86), temporalWeight 0.5, GrayScale; matrix only "
                                            "(PNG encode and image download excluded)"}
            if not precise:
                # executed: 3 split-FP16 products x 14 channels x 6 K steps of M128 x N128 x K16 per 128 x 128 tile
                tiles = (g["imgExt"] + 127) // 128
                tiles = tiles * (tiles + 1) // 2
                exec_flop = tiles * 252 * 2.0 * 128 * 128 * 16
                out[name]["roofline"] = {
                    "bound": "tensor", "achieved": flop, "peak": tensor_peak, "unit": "TFLOP/s", "frac": flop / tensor_peak,
                    "peak_source": "dense bf16 (= fp16) matmul, sustained, MEASURED_PEAKS.json",
                    "executed_tflops": exec_flop / (ms * 1e-3) / 1e12, "executed_frac": exec_flop / (ms * 1e-3) / 1e12 / tensor_peak,
                    "mma_floor_ms": tiles / 148.0 * 252 * 64 / 1.965e6,
                    "frac_of_mma_floor": tiles / 148.0 * 252 * 64 / 1.965e6 / ms,
                    "note": "FP32-grade result from three FP16 products (a1 b1 + a2 b1 + a1 b2): executed flops are 3.35x "
                            "the algorithmic ones (x3 products, K padded 86 -> 96); M128 x N128 x K16 runs at its 64-cycle "
                            "floor in isolation (tools/umma_rate_probe.cu)"}
        seg, _ = synth.regime_file(synth.BASE_SEED, 31, FRAMES_PER_FILE, 14, 26)
        scfg = N.SegmConfig(STEP, 0, 0, 0, 0, 22050, 0.5, 20, 22050)
        engine.segm_run(ctx, scfg, seg, norm)
        _, _, noff = engine.segm_run(ctx, scfg, seg, norm, want_curve=True)
        ms, _ = ctx.last_timing()
        out["segmentation_exact_fp64"] = {"metric": "FeatureSegmentation offsets/sec", "value": noff / (ms * 1e-3),
                                          "unit": "offsets/s", "offsets": int(noff), "kernel_ms": ms,
                                          "workload": "10 min synthetic file, corrLen 0.5 s, 20 breaks "
                                                      "(BASELINE.json configs[1])"}
    except Exception as e:   # secondary numbers must never break the headline line
        out["error"] = repr(e)
    return out


def run_secondary_sharded(torch, dist, device, ctx, engine, N, synth, norm, rank, world):
    """SelfSimilarity cells/s with the matrix sharded by column blocks over the ranks (BASELINE.json configs[3], the feature
    file replicated on every GPU): every rank renders its block, the time is the max over ranks of the device time."""
    from strugatzki_b200.distributed import selfsim_column_blocks
    ms, cells, ext, err = 0.0, 0, 0, None
    try:
        frames = 155000
        f = synth.regime_file(synth.BASE_SEED, 4, frames, 14, max(4, frames // 2000))[0]
        cfg = N.SelfConfig(STEP, 0, 0, 0, 0, 44100, 1, 0.5, 0, 1.0, 1.0, None, 0, 0)
        ext = engine.self_geometry(cfg, frames, frames)["imgExt"]
        b, e = selfsim_column_blocks(ext, world)[rank]
        if e > b:
            engine.self_run(ctx, cfg, f, None, norm, b, e, download=False)
            engine.self_run(ctx, cfg, f, None, norm, b, e, download=False)
            ms, _ = ctx.last_timing()
        cells = sum(ext - a for a in range(b, e))
    except Exception as ex:   # secondary numbers must never break the headline line -- nor leave a rank out of the collective
        err = repr(ex)
    t = torch.tensor([ms, float(cells), 1.0 if err else 0.0], dtype=torch.float64, device=device)
    mx, sm = t.clone(), t.clone()
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    dist.all_reduce(sm, op=dist.ReduceOp.SUM)
    if rank != 0:
        return None
    if float(mx[2]) > 0 or float(mx[0]) <= 0:
        return {"error": err or "a rank failed"}
    return {"selfsimilarity_gram_tc_config3_sharded": {
        "metric": "SelfSimilarity cells/sec", "value": float(sm[1]) / (float(mx[0]) * 1e-3), "unit": "cells/s",
        "cells": int(sm[1]), "imgExt": ext, "n_gpus": world, "kernel_ms_max_over_ranks": float(mx[0]),
        "kernel": engine.self_last_kernel(ctx),
        "workload": "155 000-frame synthetic feature file (BASELINE.json configs[3]), column blocks balanced by cell "
                    "count, one per GPU; matrix only (image gather and PNG encode excluded)"}}


def cpu_baseline(args, synth, norm, inp):
    from oracle import oracle as O
    mu, sigma, floor0, _ = synth.default_profile(14)
    p = O.CorrParams(step_size=STEP, input=inp, punch_in=(0, W_IN_SAMPLES), norm=norm, num_matches=100,
                     num_per_file=1, min_spacing=22050)
    f0 = synth.synth_file(synth.BASE_SEED, 1, FRAMES_PER_FILE, mu, sigma, floor0)
    t = time.perf_counter()
    O.corr_search(p, [f0])
    per_file = time.perf_counter() - t
    n_files = int(max(2, min(60, round(args.cpu_seconds / max(per_file, 1e-3)))))
    files = [f0] + [synth.synth_file(synth.BASE_SEED, 1 + i, FRAMES_PER_FILE, mu, sigma, floor0)
                    for i in range(1, n_files)]
    n = O.corr_num_offsets(p, [f.shape[0] for f in files])
    t = time.perf_counter()
    O.corr_search(p, files)
    dt = time.perf_counter() - t
    return {"value": n / dt, "unit": "offsets/s", "cores": 1, "kind": "port",
            "sample": f"first {n_files} of the 6000 DB files ({n} offsets, {dt:.1f} s), C restatement of the "
                      f"reference's single-threaded Double loops, features pre-loaded in RAM"}


if __name__ == "__main__":
    sys.exit(main())
