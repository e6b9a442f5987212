#!/usr/bin/env python3
"""bench.py — throughput of the BioEM likelihood hot path on B200.

  python bench.py --gpus N --steps K --warmup W            (our arm, C ABI -> sm_100a kernels)
  python bench.py --impl reference --gpus N --steps K --warmup W   (reference CPU code on host cores)

Workload (BASELINE.json configs[1], "cfg2"): synthetic ~1k-point model vs 1000 synthetic
224x224 particles, QUATERNION_LIST_4608_Orient, production CTF grid (4 x 8 x 1 = 32 kernels),
DISPLACE_CENTER 40 1  ->  4608 * 32 * 1000 = 147,456,000 likelihoods per step.
A step = one full pass of the hot path (projection .. log-sum-exp/arg-max merge) over the whole
orientation grid.  With N GPUs the orientation grid is sharded in contiguous blocks (one process
per GPU), partial results are merged with one NCCL all-gather per step (strong scaling).

metric  likelihoods/s  = image x orientation x CTF triples per second (whole job, all GPUs)
value   inputs resident in HBM before the timed region (CUDA events on the library's stream)
e2e     the same job through the public API from pinned HOST buffers: upload (H2D) + device
        particle FFTs + run + download (D2H) inside the timed region
roofline  dominant kernel = likelihood_kernel<224>; algorithmic bytes per likelihood = 8*F =
        202,496 B (one streamed read of the particle half-spectrum) over the kernel's
        CUDA-event time, against the measured HBM copy bandwidth (MEASURED_PEAKS.json)
cpu_baseline  the unmodified reference (oracle/_ref/bioEM_ref = reference sources + FFTW-API
        shim) on the host cores, on a bounded slice of the same workload
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

WORKLOADS = {
    # name: (case name, particles, orientations)
    "cfg2": ("cfg2", 1000, 4608),
    "cfg1": ("cfg1", 10, 576),
}


def _traffic(kernel: str, likelihoods_per_launch: float):
    """dram bytes per (average) launch of the dominant kernel, from the committed ncu capture
    (bytes per likelihood x likelihoods per launch), or None."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        return int(float(json.load(open(p))[kernel]["dram_bytes_per_likelihood"]) * likelihoods_per_launch)
    except Exception:
        return None


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""

    def __init__(self, gpu_index: int):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                for k, nme in enumerate(names):
                    if r[3 + k].lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                pass
        # samples under load = upper half (the region is bracketed by idle moments)
        sm_sorted = sorted(sm)
        load = sm_sorted[len(sm_sorted) // 2:] if sm_sorted else []
        med = load[len(load) // 2] if load else None
        return {"sm_mhz": med, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def build_workload(name: str):
    from bioem_b200 import api
    from bioem_b200.cases import build_case
    cname, m, o = WORKLOADS[name]
    cd = build_case(cname)
    hi, parts = api.inputs_for_case(cd)
    return cd, hi, parts


def run_ours(args):
    import torch
    import torch.distributed as dist
    from bioem_b200 import api

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the hot path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        # stdout carries exactly one JSON line: NCCL's own output (the version banner it prints at
        # NCCL_DEBUG=VERSION / WARN, warnings) goes to stderr
        # (NCCL honours NCCL_DEBUG_FILE only above the VERSION level)
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() in ("VERSION", "NONE", ""):
            os.environ["NCCL_DEBUG"] = "WARN"
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    cd, hi, parts = build_workload(args.workload)
    case = cd.case
    O, Cn, M, N = hi.O, hi.C, parts.shape[0], hi.N
    o_lo, o_hi = rank * O // world, (rank + 1) * O // world  # reference bioem.cpp:748-753
    likelihoods_step = O * Cn * M

    eng = api.Engine(hi.cfg, local)
    eng.upload_all(hi, parts)
    stream = torch.cuda.ExternalStream(api.lib().bioem_b200_stream(eng._h), device=local)
    pbytes = eng.partial_bytes()
    mine = torch.empty(pbytes, dtype=torch.uint8, device="cuda")
    gathered = torch.empty(pbytes * world, dtype=torch.uint8, device="cuda")

    def step():
        eng.reset()
        eng.run(o_lo, o_hi)
        if world > 1:
            eng.export_partial(mine.data_ptr())
            dist.all_gather_into_tensor(gathered, mine)
            torch.cuda.current_stream().synchronize()
            eng.import_partials(gathered.data_ptr(), world)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    eng.reset()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    t_wall0 = time.perf_counter()
    lik_ms = 0.0
    lik_launches = 0
    total_launches = 0
    for _ in range(args.steps):
        step()
        if args.steps <= 4:  # kernel-only time of each step (events were recorded inside run())
            ms, nl = eng.kernel_time()
            lik_ms += ms
            lik_launches += nl
        total_launches += eng.stats()[0]
    e1.record(stream)
    barrier()
    t_wall = time.perf_counter() - t_wall0
    ms_dev = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    t = torch.tensor([ms_dev], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    pm, _ = eng.download()

    # ---- end to end through the public API, host buffers, copies inside the timed region
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()  # noqa: E731
    h_parts, h_ctf, h_ang = pin(parts), pin(hi.refCTF), pin(hi.angles)
    h2d = h_parts.nbytes + h_ctf.nbytes + h_ang.nbytes + hi.points.nbytes + hi.CtfParam.nbytes
    out_maps = np.zeros(M, dtype=api.PROB_MAP_DTYPE)
    e2e_steps = max(1, min(2, args.steps))

    def e2e_step():
        e = api.Engine(hi.cfg, local)
        e.upload_model(hi.points, hi.NormDen)
        e.upload_orientations(h_ang)
        e.upload_ctf(h_ctf, hi.CtfParam)
        e.upload_particles(h_parts)
        e.reset()
        e.run(o_lo, o_hi)
        if world > 1:
            e.export_partial(mine.data_ptr())
            dist.all_gather_into_tensor(gathered, mine)
            torch.cuda.current_stream().synchronize()
            e.import_partials(gathered.data_ptr(), world)
        e.download(out_maps)
        e.close()

    e2e_val = None
    if not args.no_e2e:
        e2e_step()  # warm
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        barrier()
        te = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e2e_val = likelihoods_step * e2e_steps / float(te.item())

    if rank == 0:
        hbm_peak, peak_src = _peaks()
        F = N * (N // 2 + 1)
        value = likelihoods_step * args.steps / (ms_total / 1e3)
        roof = None
        if lik_launches:
            # per-rank kernel time covers this rank's share of the likelihoods
            per_rank_lik = (o_hi - o_lo) * Cn * M * args.steps
            ach = per_rank_lik * 8.0 * F / (lik_ms / 1e3) / 1e9
            flop = 2.5 * N * N * np.log2(N * N)
            roof = {"bound": "hbm", "achieved": round(ach, 1), "peak": hbm_peak, "unit": "GB/s",
                    "frac": round(ach / hbm_peak, 4), "traffic": _traffic(f"likelihood_kernel<{N}>", per_rank_lik / lik_launches),
                    "peak_source": peak_src,
                    "kernel": f"likelihood_kernel<{N}>", "launches": lik_launches,
                    "avg_launch_ms": round(lik_ms / lik_launches, 3),
                    "algorithmic_bytes_per_likelihood": 8 * F,
                    "fp32_algorithmic_tflops": round(per_rank_lik * flop / (lik_ms / 1e3) / 1e12, 2),
                    # compute side of the two-sided FFT roofline: 148 SMs x 128 FP32 lanes x 2 x SM clock
                    "fp32_peak_tflops": round(148 * 128 * 2 * (clocks.get("sm_mhz") or clocks.get("sm_max_mhz") or 1965.0) / 1e6, 1),
                    "fp32_frac": round(per_rank_lik * flop / (lik_ms / 1e3) / 1e12
                                       / (148 * 128 * 2 * (clocks.get("sm_mhz") or clocks.get("sm_max_mhz") or 1965.0) / 1e6), 4),
                    "kernel_share_of_step": round(lik_ms / (ms_total), 4)}
        line = {
            "metric": "likelihoods/s", "value": round(value, 1), "unit": "likelihoods/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(ms_total / args.steps, 3), "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.workload}: {O} orientations x {Cn} CTF x {M} particles "
                                   f"{N}x{N}, DISPLACE_CENTER {case.max_disp} {case.grid_space}",
                       "likelihoods_per_step": likelihoods_step,
                       "parallelism": f"orientation-sharded x{world}" if world > 1 else "single GPU",
                       "l2": "inputs larger than L2 (particle spectra 202 MB re-streamed per orientation group + 1.07 GB of conv spectra per launch)"},
            "clocks": clocks,
            "e2e": {"value": round(e2e_val, 1) if e2e_val else None, "unit": "likelihoods/s", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(out_maps.nbytes), "steps": e2e_steps},
            "gpu_launches": int(total_launches),
            "roofline": roof,
            "wall_s_timed_region": round(t_wall, 3),
        }
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline(args.workload, budget_s=args.cpu_seconds)
            line["reference_gpu"] = reference_gpu(args.workload)
        print(json.dumps(line), flush=True)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


def _ref_slice(workload: str, n_orient: int, n_part: int, workdir: str, threads: int):
    """Run oracle/_ref/bioEM_ref on the first n_orient orientations x all CTFs x n_part particles
    of the workload; returns (likelihoods, seconds of the reference's own run() timer)."""
    from bioem_b200.cases import build_case, reference_cli
    cname, _, _ = WORKLOADS[workload]
    cd = build_case(cname, workdir, n_particles=n_part, n_orient=n_orient)
    refbin = os.path.join(ROOT, "oracle", "_ref", "bioEM_ref")
    env = {**os.environ, "OMP_NUM_THREADS": str(threads)}
    env.pop("GPU", None)
    r = subprocess.run([refbin] + reference_cli(cd), cwd=workdir, env=env, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("reference binary failed: " + r.stdout[-500:] + r.stderr[-500:])
    sec = None
    for ln in r.stdout.splitlines():
        if "The code ran for" in ln:
            sec = float(ln.split("for")[1].split("seconds")[0])
    return cd.case.likelihoods, sec


def cpu_baseline(workload: str, budget_s: float = 15.0) -> dict:
    refbin = os.path.join(ROOT, "oracle", "_ref", "bioEM_ref")
    threads = os.cpu_count() or 1
    if not os.path.exists(refbin):
        return {"value": None, "unit": "likelihoods/s", "cores": threads, "kind": "reference",
                "sample": "oracle/_ref/bioEM_ref not present"}
    with tempfile.TemporaryDirectory() as d:
        cname, m_full, o_full = WORKLOADS[workload]
        # the reference parallelises over images (bioem.cpp:1392): give every thread some
        n_part = min(m_full, max(64, 4 * threads))
        n, s = _ref_slice(workload, 2, n_part, os.path.join(d, "probe"), threads)
        rate = n / max(s, 1e-6)
        # scale the slice to ~budget_s of CPU work (orientations first, then particles)
        per_orient = rate and (n / 2) / rate
        n_or = int(max(2, min(o_full, budget_s / max(per_orient, 1e-9))))
        if n_or > 64 and n_part < m_full:
            n_part = min(m_full, n_part * max(1, n_or // 64))
            n_or = 64
        n, s = _ref_slice(workload, n_or, n_part, os.path.join(d, "run"), threads)
    return {"value": round(n / s, 1), "unit": "likelihoods/s", "cores": threads, "kind": "reference",
            "sample": f"unmodified reference (FFTW-API shim FFT, Algo 1, OpenMP {threads} threads) on the first "
                      f"{n_or} orientations x all CTFs x {n_part} particles of {workload} = {n} likelihoods in {s:.2f} s "
                      f"(reference's own run() timer)"}


def reference_gpu(workload: str, n_orient: int = 8) -> dict:
    """The reference's own CUDA path (bioem_cuda.cu + cuFFT rebuilt for sm_100a, oracle/_ref/bioEM_ref_cuda,
    GPU=1 GPUWORKLOAD=100) on the first n_orient orientations of the workload on this box's GPU 0:
    likelihoods/s from its own per-orientation timer ("reference GPU on the same box").
    Reported beside the CPU baseline; never part of a timed region of ours."""
    refbin = os.path.join(ROOT, "oracle", "_ref", "bioEM_ref_cuda")
    if not os.path.exists(refbin):
        return {"value": None, "unit": "likelihoods/s", "sample": "oracle/_ref/bioEM_ref_cuda not present"}
    from bioem_b200.cases import build_case, reference_cli
    cname, m_full, _ = WORKLOADS[workload]
    try:
        with tempfile.TemporaryDirectory() as d:
            cd = build_case(cname, d, n_particles=m_full, n_orient=n_orient)
            env = {**os.environ, "GPU": "1", "GPUWORKLOAD": "100", "GPUDEVICE": "0", "BIOEM_DEBUG_OUTPUT": "1",
                   "OMP_NUM_THREADS": str(os.cpu_count() or 1)}
            r = subprocess.run([refbin] + reference_cli(cd), cwd=d, env=env, capture_output=True, text=True,
                               timeout=300)
        if r.returncode != 0:
            raise RuntimeError((r.stdout[-300:] + r.stderr[-300:]).replace("\n", " | "))
        mean = None
        for ln in r.stdout.splitlines():
            if "Total time of projection" in ln:
                mean = float(ln.split("Mean")[1].split("sec")[0])
        per_or = cd.case.n_ctf * cd.case.n_particles
        return {"value": round(per_or / mean, 1), "unit": "likelihoods/s",
                "sample": f"unmodified reference CUDA path (bioem_cuda.cu + cuFFT, nvcc sm_100a, GPU=1 GPUWORKLOAD=100) "
                          f"on 1 GPU: mean {mean * 1e3:.1f} ms per orientation over the first {n_orient} orientations x "
                          f"{cd.case.n_ctf} CTFs x {cd.case.n_particles} particles of {workload} (its own timer)"}
    except Exception as e:  # a reported side number: never fail the bench for it
        return {"value": None, "unit": "likelihoods/s", "sample": f"failed: {e}"[:300]}


def run_reference(args):
    """Reference arm: the reference's own CPU implementation of the path on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    refbin = os.path.join(ROOT, "oracle", "_ref", "bioEM_ref")
    if not os.path.exists(refbin):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/bioEM_ref was not built"}))
        return
    cname, m_full, o_full = WORKLOADS[args.workload]
    from bioem_b200.cases import CASES
    case = CASES[cname]
    with tempfile.TemporaryDirectory() as d:
        n_part = min(m_full, max(64, 4 * threads))
        n, s = _ref_slice(args.workload, 2, n_part, os.path.join(d, "probe"), threads)
        per_orient = s / 2
        n_or = int(max(2, min(o_full, args.cpu_seconds / max(per_orient, 1e-9))))
        times = []
        for k in range(args.warmup + args.steps):
            n, s = _ref_slice(args.workload, n_or, n_part, os.path.join(d, f"s{k}"), threads)
            if k >= args.warmup:
                times.append(s)
    total = sum(times)
    value = n * len(times) / total
    sample = (f"each step = unmodified reference (oracle/_ref/bioEM_ref: reference sources + FFTW-API shim FFT, "
              f"Algo 1, OpenMP {threads} threads) on the first {n_or} orientations x {case.n_ctf} CTFs x {n_part} "
              f"particles of {args.workload} = {n} likelihoods; time = the reference's own run() timer")
    line = {
        "impl": "reference", "metric": "likelihoods/s", "value": round(value, 1), "unit": "likelihoods/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(1e3 * total / len(times), 3), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {o_full} orientations x {case.n_ctf} CTF x {m_full} particles "
                               f"{case.n_pixels}x{case.n_pixels}, DISPLACE_CENTER {case.max_disp} {case.grid_space}",
                   "likelihoods_per_step": n, "parallelism": f"host OpenMP x{threads}"},
        "cpu_baseline": {"value": round(value, 1), "unit": "likelihoods/s", "cores": threads, "kind": "reference",
                         "sample": sample},
        "e2e": {"value": round(value, 1), "unit": "likelihoods/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="CPU work budget of one reference sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end leg (profiling runs only)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
