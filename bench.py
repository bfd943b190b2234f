#!/usr/bin/env python
"""Benchmark of the nano_gicp hot path on B200 (BASELINE.json config C2).

One STEP = one scan-to-map registration of a 64x1024 scan against a resident 500k-point keyframe
submap, i.e. what OdomNode does per LiDAR frame on this path (odom.cc:518-532, 745-793) once the
submap is unchanged:
    setInputSource(scan)            -> kNN index over the scan
    calculateSourceCovariances()    -> k=20 neighbourhood covariances (PLANE)
    align(guess)                    -> device-resident LM loop (1-NN + Mahalanobis + H/b, error trials)
The submap's points, index and covariances stay on the device between steps (the reference keeps
its target kd-tree and covariance vector while `submap_hasChanged_` is false, odom.cc:777-784).

  value      registrations/s over all ranks, scan already in HBM, L2 flushed before every step,
             CUDA events on the library's stream.  ms_per_step is the ms/scan of BASELINE's metric.
  e2e        the same step driven through the public API from pinned HOST memory: upload, index,
             covariances, align, result + residual read-back; host wall clock.
  roofline   the align kernel (k_align): algorithmic bytes / its measured duration.
  cpu_baseline / --impl reference
             the same step on the host cores with all OpenMP threads: the reference's OWN nano_gicp engine
             (oracle/_ref/libnano_gicp_ref.so, its headers compiled unmodified over Eigen/PCL/Boost stand-ins);
             if that library is missing, the oracle's restatement over the reference's nanoflann.

Multi-GPU (torchrun, one rank per GPU): every rank registers its own scans against its own copy of
the submap; no data-path collective (independent registrations), weak scaling.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

NS_BEAMS, NS_COLS, NT = 64, 1024, 500_000
K_COV = 20
METRIC = "gicp_s2m_registrations_per_s"
UNIT = "registrations/s"
L2_FLUSH_BYTES = 256 << 20

# algorithmic bytes per unit (DESIGN.md "traffic model")
B_INDEX, B_COV, B_LIN, B_ERR = 36, 64, 184, 84


def make_workload(rank: int):
    from dynamic_direct_lidar_odometry_b200 import synth

    src, tgt, guess = synth.workload_c2(NT, NS_BEAMS, NS_COLS, src_frame=50 + rank)
    return src, tgt, guess


def config_dict(n_src: int, extra=None):
    cfg = {
        "workload": "C2: S2M registration, 64x1024 synthetic scan vs 500k-point synthetic keyframe submap",
        "source_points": int(n_src),
        "target_points": NT,
        "k_correspondences": K_COV,
        "step": "source index build + source covariances (k=20, PLANE) + LM align; target index/covariances resident",
        "engine_params": "reference defaults (max_iter 64, trans_eps 5e-4, rot_eps 2e-3, LM)",
        "l2": f"flushed before every timed step ({L2_FLUSH_BYTES >> 20} MiB write)",
    }
    if extra:
        cfg.update(extra)
    return cfg


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    QUERY = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "50", "-i", str(self.gpu)],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
            t0 = time.time()  # nvidia-smi needs a moment to attach: wait for its first line (bounded)
            while time.time() - t0 < 3.0 and os.path.getsize(self.path) == 0:
                time.sleep(0.02)
        except Exception:
            self.proc = None

    def mark(self):
        """number of samples so far (to select the ones taken during the timed region)"""
        try:
            return sum(1 for _ in open(self.path))
        except Exception:
            return 0

    def stop(self, first: int = 0):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if not self.proc:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for line in list(open(self.path))[first:]:
                f = [x.strip() for x in line.split(",")]
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
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out


# ------------------------------------------------------------------------------------------------ CPU arm
def use_all_host_cores():
    """torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU arm must run on all the cores it can use.  Has to happen
    before the OpenMP runtime is loaded (the oracle / reference libraries bring it in)."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    os.environ["OMP_NUM_THREADS"] = str(n)
    os.environ.pop("OMP_THREAD_LIMIT", None)
    return n


def cpu_reference_setup(src, tgt, prefer_reference_engine=True):
    """Host-core arm.  Preferred: the reference's OWN nano_gicp engine (oracle/_ref/libnano_gicp_ref.so: its headers
    compiled unmodified from /root/reference, over the Eigen/PCL/Boost stand-ins of oracle/stub_include since none of
    the three is installed) - kind "reference".  Fallback: the oracle's restatement over the reference's vendored
    nanoflann - kind "port"."""
    from oracle import pyoracle as po

    mod, kind, what = None, None, None
    if prefer_reference_engine:
        from oracle import refgicp

        if refgicp.available():
            mod, kind = refgicp, "reference"
            what = "the reference's own nano_gicp sources (oracle/_ref/libnano_gicp_ref.so, Eigen/PCL/Boost stand-ins), OpenMP"
            eng = refgicp.NanoGICP()
    if mod is None:
        backend = po.BACKEND_NANOFLANN_REF if po.load_reference_nanoflann() else po.BACKEND_CANONICAL
        mod, kind = po, "port"  # GICP/LM layer is the restatement; the kd-tree under it is the reference's own code when available
        what = "oracle restatement of GICP/LM over " + ("reference nanoflann 1.3.2 (oracle/_ref)" if backend == po.BACKEND_NANOFLANN_REF else "oracle kd-tree")
        eng = po.NanoGICP(backend=backend)
    T = mod.Cloud(tgt)
    eng.setInputTarget(T)          # kd-tree over the submap: built once, outside the timed region (steady state)
    eng.calculateTargetCovariances()
    return mod, eng, T, kind, what


def cpu_step(po, eng, src, guess):
    t0 = time.perf_counter()
    S = po.Cloud(src)
    eng.setInputSource(S)              # kd-tree build over the scan
    eng.calculateSourceCovariances()
    r = eng.align(guess)
    return time.perf_counter() - t0, r


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    use_all_host_cores()
    src, tgt, guess = make_workload(0)
    po, eng, _, kind, knn = cpu_reference_setup(src, tgt)
    from oracle import pyoracle as _po

    cores = _po.max_threads()
    for _ in range(max(args.warmup, 0)):
        cpu_step(po, eng, src, guess)
        eng.clearSource()
    times = []
    for _ in range(args.steps):
        dt, r = cpu_step(po, eng, src, guess)
        times.append(dt)
        eng.clearSource()
    total = sum(times)
    value = args.steps / total
    sample = f"{args.steps} full registrations (scan kd-tree + covariances + LM align, {r.iterations + 1} outer iterations each); {knn}"
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 kNN / f64 GICP", "data": "synthetic", "impl": "reference",
        "config": config_dict(len(src)),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------ GPU arm
def run_ours(args):
    from dynamic_direct_lidar_odometry_b200 import nano_gicp as ng

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist

        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    src, tgt, guess = make_workload(rank)
    rt = ng.Runtime(local_rank)
    eng = ng.NanoGICP(rt)
    eng.setCorrespondenceRandomness(K_COV)
    target = ng.PointCloud(rt, tgt)
    eng.setInputTarget(target)
    eng.calculateTargetCovariances()
    rt.synchronize()

    n_src = len(src)
    src_pinned = ng.pinned_array((n_src, 4), np.float32)
    src_pinned[:] = src
    resident = ng.PointCloud(rt, src)  # the scan, already in HBM, for the device-timed arm
    rt.synchronize()

    def device_step(timed: bool):
        """one step with the scan resident in HBM; returns stage times (ms) and the align info"""
        fresh = resident.transformed(np.eye(4, dtype=np.float32))  # a new, index-less cloud handle (device copy, untimed)
        rt.flush_l2(L2_FLUSH_BYTES)
        rt.event_record(0)
        eng.setInputSource(fresh)            # index build
        rt.event_record(1)
        eng.calculateSourceCovariances()
        rt.event_record(2)
        eng.align_async(guess)
        rt.event_record(3)
        info = eng.align_finish()
        t = (rt.event_elapsed(0, 1), rt.event_elapsed(1, 2), rt.event_elapsed(2, 3), rt.event_elapsed(0, 3))
        eng.clearSource()
        return t, info

    res_pinned = ng.pinned_array((n_src,), np.float64)  # where getResiduals lands (odom.cc:793)

    def e2e_step():
        t0 = time.perf_counter()
        cloud = ng.PointCloud(rt, src_pinned)          # H2D from pinned host memory
        eng.setInputSource(cloud)
        eng.calculateSourceCovariances()
        eng.align_async(guess)
        eng.getResidualsAsync(res_pinned)              # D2H: per-point residuals, enqueued behind the align
        info = eng.align_finish()                      # D2H: result struct; the one host synchronisation of the step
        dt = time.perf_counter() - t0
        eng.clearSource()
        return dt, info, res_pinned

    for _ in range(max(args.warmup, 3)):
        device_step(False)
        e2e_step()
    rt.synchronize()

    def barrier():
        if dist is not None:
            dist.barrier()
        rt.synchronize()

    sampler = ClockSampler(local_rank)
    first_sample = 0
    if rank == 0:
        sampler.start()
        for _ in range(50):  # keep the GPU under load while the sampler spins up, so that its first samples are loaded ones
            device_step(False)
        first_sample = sampler.mark()
    launches0 = rt.launch_count()
    barrier()
    stage = []
    info = None
    for _ in range(args.steps):
        t, info = device_step(True)
        stage.append(t)
    barrier()
    launches1 = rt.launch_count()
    total_ms = sum(t[3] for t in stage)

    barrier()
    e2e_times = []
    for _ in range(args.steps):
        dt, info_e, res = e2e_step()
        e2e_times.append(dt)
    barrier()
    clocks = sampler.stop(first_sample) if rank == 0 else None
    e2e_total = sum(e2e_times)

    # Batched throughput (the second half of BASELINE's metric): the same step for independent scans, S host threads
    # per GPU, each with its own runtime (stream), engine and copy of the submap; every stream's align kernel is limited
    # to 148 // S SMs so that the cooperative launches of all streams are resident side by side.  No L2 flush here: the
    # S working sets together (S x ~50 MB of submap points, covariances and index) exceed the L2.
    batched_s, batched_n = 0.0, 0
    S = args.batched_streams
    if S > 0:
        import threading

        workers = []
        try:  # an extra leg: whatever happens here must not cost the main line (no barrier inside: a failing rank cannot hang the others)
            per_thread = max(8, min(args.steps, 200) // S)
            for _ in range(S):
                brt = ng.Runtime(local_rank)
                brt.set_align_blocks(max(1, 148 // S))
                beng = ng.NanoGICP(brt)
                beng.setCorrespondenceRandomness(K_COV)
                btarget = ng.PointCloud(brt, tgt)
                beng.setInputTarget(btarget)
                beng.calculateTargetCovariances()
                bres = ng.PointCloud(brt, src)
                brt.synchronize()
                workers.append((brt, beng, btarget, bres))
            errors = []

            def batched_worker(w, count):
                try:
                    brt, beng, _, bres = w
                    for _ in range(count):
                        fresh = bres.transformed(np.eye(4, dtype=np.float32))
                        beng.setInputSource(fresh)
                        beng.calculateSourceCovariances()
                        beng.align(guess)
                        beng.clearSource()
                    brt.synchronize()
                except Exception as exc:  # noqa: BLE001
                    errors.append(exc)

            def run_batched(count):
                th = [threading.Thread(target=batched_worker, args=(w, count)) for w in workers]
                for t_ in th:
                    t_.start()
                for t_ in th:
                    t_.join()

            run_batched(3)
            t0 = time.perf_counter()
            run_batched(per_thread)
            if not errors:
                batched_s = time.perf_counter() - t0
                batched_n = per_thread * S
        except Exception:  # noqa: BLE001
            batched_s, batched_n = 0.0, 0
        for brt, beng, btarget, bres in workers:
            del beng, btarget, bres
            brt.close()
        workers = []

    if dist is not None:
        import torch

        t = torch.tensor([total_ms, e2e_total, batched_s, 0.0 if batched_n else 1.0], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, e2e_total, batched_s = float(t[0]), float(t[1]), float(t[2])
        if float(t[3]) > 0.0:  # some rank has no batched result
            batched_n = 0

    if rank == 0:
        align_ms = statistics.mean(t[2] for t in stage)
        L, E = info.n_linearize, info.n_compute_error
        align_bytes = (L * B_LIN + E * B_ERR) * n_src
        try:
            peaks = json.load(open(ROOT / "MEASURED_PEAKS.json"))
            peak, peak_src = float(peaks["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured)"
        except Exception:
            peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
        achieved = align_bytes / (align_ms * 1e-3) / 1e9
        traffic = None
        tfile = ROOT / "profiles" / "align_traffic.json"
        if tfile.exists():
            try:
                traffic = json.load(open(tfile)).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        line = {
            "metric": METRIC, "value": world * args.steps / (total_ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32 kNN / f64 GICP", "data": "synthetic",
            "config": config_dict(n_src, {"parallelism": f"{world} independent replica(s), no collective"}),
            "stages_ms": {"source_index": statistics.mean(t[0] for t in stage), "source_covariances": statistics.mean(t[1] for t in stage),
                          "align": align_ms},
            "align": {"converged": info.converged, "outer_iterations": info.iterations + 1, "n_linearize": L, "n_compute_error": E},
            "roofline": {"bound": "hbm", "kernel": "k_align (1 cooperative launch per align)", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "algorithmic_bytes": align_bytes, "peak_source": peak_src,
                         "note": "latency-bound by construction: 65k points x ~100 MB of algorithmic traffic per align (SURVEY.md finding 6)"},
            "e2e": {"value": world * args.steps / e2e_total, "unit": UNIT, "ms_per_step": 1e3 * e2e_total / args.steps,
                    "h2d_bytes_per_step": int(src_pinned.nbytes), "d2h_bytes_per_step": int(232 + res.nbytes), "timer": "host wall clock"},
            "gpu_launches": int(launches1 - launches0),
            "clocks": clocks,
        }
        if batched_n:
            line["batched"] = {"value": world * batched_n / batched_s, "unit": UNIT, "streams_per_gpu": S, "align_sms_per_stream": max(1, 148 // S),
                               "registrations": world * batched_n, "timer": "host wall clock, max over ranks",
                               "note": "independent registrations of the same workload driven concurrently from S host threads per GPU "
                                       "(own stream, engine and submap copy each); `value` above is one stream, i.e. 1000 / ms_per_scan"}
        if world == 1 and not args.no_cpu_baseline:
            use_all_host_cores()
            po, ceng, _, kind, knn = cpu_reference_setup(src, tgt)
            from oracle import pyoracle as _po
            cores = _po.max_threads()
            cpu_step(po, ceng, src, guess)
            ceng.clearSource()
            ct, n = 0.0, 0
            while ct < 12.0 and n < 400:
                dt, r = cpu_step(po, ceng, src, guess)
                ceng.clearSource()
                ct += dt
                n += 1
            line["cpu_baseline"] = {"value": n / ct, "unit": UNIT, "cores": cores, "kind": kind,
                                    "sample": f"{n} full registrations of the same workload in {ct:.1f} s; {knn}",
                                    "ms_per_step": 1e3 * ct / n}
            if kind == "reference":  # for comparison: the oracle's restatement on the same inputs (a few steps)
                po2, peng, _, _, _ = cpu_reference_setup(src, tgt, prefer_reference_engine=False)
                cpu_step(po2, peng, src, guess)
                peng.clearSource()
                pts = []
                for _ in range(5):
                    pts.append(cpu_step(po2, peng, src, guess)[0])
                    peng.clearSource()
                line["cpu_baseline"]["oracle_port_ms_per_step"] = 1e3 * min(pts)
        print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    del eng, target, resident
    rt.close()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=500)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--batched-streams", type=int, default=4, help="host threads / streams per GPU for the batched-throughput leg (0 = skip)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
