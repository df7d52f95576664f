#!/usr/bin/env python3
"""bench.py -- headline benchmark of the wavelet hot path.

Metric (BASELINE.json): Mpixel/s of forward + inverse separable 2D DWT, db2, 3 levels, 8192x8192 fp32.
A "step" = one forward() + inverse() over the per-GPU batch (one 8192^2 image per GPU; weak scaling:
every rank owns its own image(s), no data-path collective).  Mpixel/s = pixels transformed / time.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

Keys of the JSON line (one line, rank 0):
  value        device-resident throughput: inputs already in HBM, CUDA events on the plan's stream,
               barrier + synchronize on both sides, max over ranks
  e2e          same metric through the public pycudwt.Wavelets API with HOST buffers: every step does
               forward(img_host) [H2D of the image from pinned memory] + inverse() + image_into(pinned) [D2H]
  roofline     dominant kernel (level-1 forward), algorithmic bytes (8 B per level-1 pixel) / its
               CUDA-event duration, against MEASURED_PEAKS.json's HBM copy bandwidth
  roofline_step  whole step at 16 B/px (the BASELINE.md headline fraction)
  cpu_baseline the plain-C/OpenMP oracle port of the same transform timed on the host cores (bounded sample)
  pdwt_cuda    the reference's own CUDA kernels (oracle/_ref, recompiled for sm_100a) on the same GPU,
               same workload, device-resident -- reported beside, not part of `value`
--impl reference: the CPU path of the reference workflow (pywt-equivalent C/OpenMP restatement
oracle/dwt_cpu.c; pywt itself is not installable here) on all host cores, same metric/config.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "Mpixel/s fwd+inv 2D DWT (db2, 3 lvl, 8192^2 fp32)"
WNAME, LEVELS, SIDE = "db2", 3, 8192


def synth(shape, seed):
    rng = np.random.default_rng(seed)
    x = rng.standard_normal(shape, dtype=np.float32) * 50 + 128
    i = np.arange(shape[-2], dtype=np.float32)[:, None]
    j = np.arange(shape[-1], dtype=np.float32)[None, :]
    x += 64 * np.sin(2 * np.pi * i / shape[-2] * 3) * np.cos(2 * np.pi * j / shape[-1] * 5)
    return x


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md: 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.rows, self.proc = device, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx = float(r[2])
                for n, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx,
                "reasons": sorted(reasons), "samples": len(sm)}


_CPU_PLANS = {}


def cpu_port_fwd_inv(img):
    """The CPU restatement of the reference path (oracle/dwt_cpu.c, plain C + OpenMP on all host threads):
    3-level db2 forward + inverse, fp32 -- stands in for pywt.wavedec2/waverec2(mode="periodization")."""
    from oracle import dwt_cpu
    P = _CPU_PLANS.get(img.shape)
    if P is None:
        P = _CPU_PLANS[img.shape] = dwt_cpu.CpuDwt2(img.shape, WNAME, LEVELS)
    P.forward(img)
    return P.inverse()


def cpu_cores():
    from oracle import dwt_cpu
    return dwt_cpu.threads()


def time_cpu_port(budget_s=float(os.environ.get("PWT_BENCH_CPU_BUDGET", "10")), side=4096):
    img = synth((side, side), 99)
    cpu_port_fwd_inv(img)                  # warm up (page faults, thread pool)
    n, t0 = 0, time.perf_counter()
    while True:
        cpu_port_fwd_inv(img)
        n += 1
        dt = time.perf_counter() - t0
        if dt >= budget_s or n >= 200:
            break
    return n * side * side / dt / 1e6, "%d x (%dx%d db2 3-level fwd+inv), oracle/dwt_cpu.c fp32, %d OpenMP threads, %.1f s" % (
        n, side, side, cpu_cores(), dt)


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def run_reference(args):
    rank, world, _ = dist_env()
    if rank != 0:
        return
    # bounded sample of the workload: one step = one SxS image on the host, S the largest of
    # 4096/2048/1024/512 for which warmup + K steps stay within ~2.5 minutes
    probe = synth((1024, 1024), 98)
    cpu_port_fwd_inv(probe)
    t0 = time.perf_counter()
    cpu_port_fwd_inv(probe)
    t_px = (time.perf_counter() - t0) / probe.size
    side = 512
    for cand in (4096, 2048, 1024):
        if (args.steps + max(args.warmup, 1)) * t_px * cand * cand * 1.3 <= 150.0:
            side = cand
            break
    img = synth((side, side), 99)
    for _ in range(max(args.warmup, 1)):
        cpu_port_fwd_inv(img)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_port_fwd_inv(img)
    dt = time.perf_counter() - t0
    val = args.steps * side * side / dt / 1e6
    cores = cpu_cores()
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "Mpixel/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "gpu_launches": 0,
        "config": {"workload": "8192x8192 fp32 db2 3-level separable DWT forward+inverse",
                   "sample": "each step = one %dx%d image (bounded sample of the 8192^2 workload)" % (side, side),
                   "l2": "n/a (CPU)"},
        "cpu_baseline": {"value": val, "unit": "Mpixel/s", "cores": cores, "kind": "port",
                         "sample": "%d steps x %dx%d, plain-C/OpenMP restatement (oracle/dwt_cpu.c) of pywt "
                                   "mode=periodization (pywt not installable here); host has %d cores" % (args.steps, side, side, os.cpu_count() or 0)},
        "e2e": {"value": val, "unit": "Mpixel/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def time_pdwt(img, steps, warmup):
    """PDWT's own CUDA build on the same image (device-resident, explicit sync through a tiny D2H)."""
    p = os.path.join(ROOT, "oracle", "_ref")
    if not os.path.isdir(p):
        return None
    sys.path.insert(0, p)
    try:
        import pycudwt_ref
    except Exception as e:      # noqa: BLE001
        return {"unavailable": str(e)[:200]}
    try:
        R = pycudwt_ref.Wavelets(img, WNAME, LEVELS)

        def sync():
            R.norm1()   # blocking cuBLAS call: the only sync point the reference API offers

        for _ in range(max(warmup, 1)):
            R.forward(); R.inverse()
        sync()
        t_sync0 = time.perf_counter(); sync(); t_sync = time.perf_counter() - t_sync0
        t0 = time.perf_counter()
        for _ in range(steps):
            R.forward(); R.inverse()
        sync()
        dt = time.perf_counter() - t0 - t_sync
        rec_err = float(np.abs(R.image - img).max())
        del R
        return {"value": steps * img.size / dt / 1e6, "unit": "Mpixel/s", "ms_per_step": dt / steps * 1e3,
                "how": "oracle/_ref (unmodified PDWT, sm_100a), wall clock around %d fwd+inv, sync via norm1()" % steps,
                "reconstruction_max_err": rec_err}
    except Exception as e:      # noqa: BLE001
        return {"unavailable": str(e)[:200]}


def run_ours(args):
    rank, world, local = dist_env()
    import pypwt_b200
    import pycudwt

    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if pypwt_b200.device_count() < 1:
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback)")
    pypwt_b200.set_device(local)

    def barrier():
        if dist is not None:
            dist.barrier()

    def max_over_ranks(x):
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    B = args.batch
    shape = (SIDE, SIDE) if B == 1 else (B, SIDE, SIDE)
    pix = B * SIDE * SIDE
    img = pypwt_b200.pinned_empty(shape)
    img[...] = synth(shape, 1234 + rank)
    out = pypwt_b200.pinned_empty(shape)
    W = pycudwt.Wavelets(img, WNAME, LEVELS)

    # ---- device-resident timing ---------------------------------------------------------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()      # runs through warm-up, the timed region and the per-kernel pass (all GPU-loaded)
    for _ in range(max(args.warmup, 3)):
        W.forward(); W.inverse()
    W.sync()
    l0 = W.launch_count
    barrier(); W.sync()
    W.timer_start()
    for _ in range(args.steps):
        W.forward(); W.inverse()
    ms = W.timer_stop()
    W.sync(); barrier()
    launches = W.launch_count - l0
    ms = max_over_ranks(ms)
    value = world * pix * args.steps / (ms * 1e-3) / 1e6

    # ---- per-kernel durations (same loop, every launch bracketed by events) --------------------
    W.profile_enable(1)
    for _ in range(min(args.steps, 40)):
        W.forward(); W.inverse()
    recs = W.profile_read()
    W.profile_enable(0)
    clocks = sampler.stop() if rank == 0 else None
    by_tag = {}
    for tag, t in recs:
        by_tag.setdefault(tag, []).append(t)
    avg = {tag: float(np.mean(v)) for tag, v in by_tag.items()}
    peak, peak_src = measured_peak()
    fused = 311 in avg
    k_ms = avg.get(311) or avg.get(101)
    # dominant kernel: the fused 3-level forward (k_fwd3) reads the image once and writes all N coefficients:
    # 8 B per pixel.  Without the fused path the level-1 kernel moves the same 8 B per pixel.
    alg_bytes = 8.0 * pix
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "r01_ncu_traffic.json")) as fh:
            traffic = json.load(fh).get("k_fwd3" if fused else "k_fwd_reg", {}).get("dram_bytes_per_launch")
    except Exception:
        pass
    roof = {"bound": "hbm", "kernel": "k_fwd3: forward levels 1-3, one launch" if fused else "level-1 forward",
            "achieved": alg_bytes / (k_ms * 1e-3) / 1e9 if k_ms else None, "peak": peak, "unit": "GB/s",
            "frac": (alg_bytes / (k_ms * 1e-3) / 1e9 / peak) if k_ms else None, "traffic": traffic,
            "traffic_source": "ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum of one launch at the 8192^2 "
                              "workload (profiles/r01_ncu_traffic.json)" if traffic else None,
            "algorithmic_bytes": alg_bytes,
            "peak_source": peak_src, "kernel_ms": k_ms,
            "kernel_ms_by_level": {{1: "fwd", 2: "inv", 11: "fwd1-", 12: "inv1-"}.get(t % 100, "k") + str(t // 100): round(v, 5) for t, v in sorted(avg.items())},
            "share_of_step": (k_ms / sum(avg.values())) if k_ms else None}
    step_gbs = 16.0 * pix / (ms / args.steps * 1e-3) / 1e9
    roof_step = {"bytes_per_px": 16, "achieved": step_gbs, "peak": peak, "unit": "GB/s", "frac": step_gbs / peak}

    # ---- end to end through the public API with host buffers -----------------------------------
    # serial: one plan; every step waits for its own D2H before the next H2D starts
    for _ in range(2):
        W.forward(img); W.inverse(); W.image_into(out)
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(4, min(args.steps, 10))
    for _ in range(e2e_steps):
        W.forward(img)          # H2D (pinned -> device) + forward
        W.inverse()
        W.image_into(out)       # D2H of the reconstructed image (synchronises)
    dt_serial = time.perf_counter() - t0
    barrier()
    dt_serial = max_over_ranks(dt_serial)
    # pipelined: two host threads, each with its own plan (own stream) and its own pinned output buffer, process
    # alternate frames; the library calls release the GIL, so the D2H of one frame overlaps the H2D of the next on
    # the full-duplex PCIe link.  Every step still copies its input from pinned host memory and its reconstructed
    # image back to pinned host memory, all inside the timed region.
    W2 = pycudwt.Wavelets(img, WNAME, LEVELS)
    out2 = pypwt_b200.pinned_empty(shape)
    plans, outs = (W, W2), (out, out2)
    for k in range(2):
        plans[k].forward(img); plans[k].inverse(); plans[k].image_into(outs[k])
    e2e_steps -= e2e_steps & 1

    def worker(k):
        P, o = plans[k], outs[k]
        for _ in range(e2e_steps // 2):
            P.forward(img)      # H2D (pinned -> device) + forward
            P.inverse()
            P.image_into(o)     # D2H of the reconstructed image (synchronises this plan's stream)

    barrier()
    threads = [threading.Thread(target=worker, args=(k,)) for k in range(2)]
    t0 = time.perf_counter()
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    dt = time.perf_counter() - t0
    barrier()
    dt = max_over_ranks(dt)
    e2e = {"value": world * pix * e2e_steps / dt / 1e6, "unit": "Mpixel/s",
           "h2d_bytes_per_step": int(img.nbytes), "d2h_bytes_per_step": int(out.nbytes),
           "steps": e2e_steps, "ms_per_step": dt / e2e_steps * 1e3,
           "how": "pycudwt.Wavelets.forward(host image) + inverse() + image_into(pinned host buffer) per step; two host "
                  "threads / plans process alternate frames so the D2H of one overlaps the H2D of the next "
                  "(PCIe-bound: %.1f GB/s per direction)" % (img.nbytes * e2e_steps / dt / 1e9),
           "serial_value": world * pix * e2e_steps / dt_serial / 1e6,
           "max_abs_reconstruction_err": float(max(np.abs(out - img).max(), np.abs(out2 - img).max()))}
    del W2

    # ---- multi-GPU: global norms through the fused reduction + NCCL all-reduce ------------------
    extra = {}
    if world > 1:
        uid = [pypwt_b200.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        W.forward()
        W.comm_init(world, rank, uid[0])
        g1, g2 = W.norms_allreduce()
        W.sync(); barrier()
        t0 = time.perf_counter()
        for _ in range(20):
            g1, g2 = W.norms_allreduce()
        extra["global_norms"] = {"norm1": g1, "norm2sq": g2,
                                 "us_per_call": (time.perf_counter() - t0) / 20 * 1e6,
                                 "how": "fused |c|,c^2 reduction kernel + ncclAllReduce(2 x f64) on the plan's stream"}
        l1, l2 = W.norms()
        extra["global_norms"]["local_norm1_rank0"] = l1
        W.comm_destroy()

    if rank == 0:
        cpu_v, cpu_s = time_cpu_port()
        pd = time_pdwt(np.asarray(img if B == 1 else img[0]), min(args.steps, 10), 2) if not args.no_pdwt else None
        line = {
            "metric": METRIC, "value": value, "unit": "Mpixel/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "8192x8192 fp32 db2 3-level separable DWT forward+inverse",
                       "per_gpu_batch": B, "l2": "inputs larger than L2 (256 MiB image + 256 MiB coefficients per image vs 126 MB L2); no flush",
                       "parallelism": "independent images per GPU, no data-path collective"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
            "roofline": roof, "roofline_step": roof_step,
            "cpu_baseline": {"value": cpu_v, "unit": "Mpixel/s", "cores": cpu_cores(), "kind": "port", "sample": cpu_s},
            "pdwt_cuda": pd,
        }
        line.update(extra)
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=1, help="8192^2 images per GPU per step")
    ap.add_argument("--no-pdwt", action="store_true", help="skip the side-by-side timing of the reference's CUDA build")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
