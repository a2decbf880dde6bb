#!/usr/bin/env python
"""Benchmark of the late-interaction hot path (BASELINE.json metric: MaxSim queries/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

A *step* is one pass of the hot path over one batch: 32 queries (N_q = 32) scored against a
1000-document bank (N_d = 1030, dim 128, bf16) = BASELINE.json configs[1].  Under torchrun (N > 1)
every rank scores the same 32 queries against ITS OWN 1000-document shard of an N*1000-document
corpus (weak scaling) and the [32, 1000] score slabs are all-gathered (NCCL) inside the timed region,
so "queries/sec" counts query x 1000-document-shard units across all ranks.

Prints ONE JSON line (rank 0).  Keys: see the contract in DESIGN.md section "Measurement".
"""

from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
# the reference arm imports the real colpali_engine (transformers / huggingface_hub) on a box without network
for _k in ("HF_HUB_OFFLINE", "TRANSFORMERS_OFFLINE", "HF_HUB_DISABLE_TELEMETRY"):
    os.environ.setdefault(_k, "1")

N_QUERIES, N_Q, N_DOCS, N_D, DIM = 32, 32, 1000, 1030, 128
FLOPS_PER_STEP = 2.0 * N_QUERIES * N_Q * N_DOCS * N_D * DIM  # 2*Bq*Nq*Bd*Nd*D
MIN_BYTES_PER_STEP = 2 * (N_DOCS * N_D * DIM + N_QUERIES * N_Q * DIM) + 4 * N_QUERIES * N_DOCS
METRIC = "maxsim_queries_per_sec"
UNIT = "queries/s"
# identical in both arms (the driver compares the `config` objects of the two JSON lines)
CONFIG = {"workload": "score_multi_vector 32q x 1000d x 1030p x 128d bf16 (BASELINE configs[1])",
          "n_queries": N_QUERIES, "query_len": N_Q, "n_docs_per_gpu": N_DOCS, "doc_len": N_D, "dim": DIM}
REF_DIR = os.path.join(ROOT, "baseline", "_ref")


def reference_scorer():
    """The CPU arm: the UNMODIFIED reference from baseline/_ref (installed by baseline/install_ref.py; it travels to
    the GPU box) called exactly as shipped -- ``BaseVisualRetrieverProcessor.score_multi_vector(qs, ps, batch_size=128,
    device="cpu")`` (processing_utils.py:132-187).  Falls back to the oracle port (same ATen calls) when the install is
    absent.  Returns (callable, kind)."""
    if os.path.isdir(os.path.join(REF_DIR, "colpali_engine")):
        if REF_DIR not in sys.path:
            sys.path.insert(0, REF_DIR)
        try:
            from colpali_engine.utils.processing_utils import BaseVisualRetrieverProcessor

            return (lambda q, d: BaseVisualRetrieverProcessor.score_multi_vector(q, d, batch_size=128, device="cpu"),
                    "reference")
        except Exception as e:  # noqa: BLE001 - a broken install must not kill the bench; say so and use the port
            print(f"[bench] baseline/_ref import failed ({e!r}); using the oracle port", file=sys.stderr)
    from oracle import li_oracle as O  # bench.py's CPU-baseline leg is allowed to execute oracle/

    return (lambda q, d: O.score_multi_vector_port(q, d, batch_size=128, device="cpu")), "port"


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"bf16_tflops": p["bf16_tflops"], "bf16_tflops_sustained": p.get("bf16_tflops_sustained"),
                "hbm_gbs": p["hbm_gbs"], "source": "measured (MEASURED_PEAKS.json)"}
    return {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0,
            "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """Samples SM clock / throttle reasons with NVML while the timed region runs."""

    def __init__(self, index: int):
        self.samples, self.reasons, self._stop = [], set(), threading.Event()
        self.ok = False
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:  # noqa: BLE001
            self.err = repr(e)
        self.t = threading.Thread(target=self._run, daemon=True)

    _NAMES = {
        0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
        0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting", 0x10: "sync_boost",
        0x100: "display_clock_setting",
    }

    def _run(self):
        nv = self.nv
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for bit, name in self._NAMES.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.002)

    def __enter__(self):
        if self.ok:
            self.t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self.ok:
            self.t.join(timeout=1)

    def summary(self):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": getattr(self, "max_mhz", None), "reasons": [], "samples": 0}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def host_cores() -> int:
    """CPU cores this process may really use: min(cpu_count, affinity, cgroup quota).  The GPU boxes show 128
    logical CPUs but cap the container at 16 (cpu.max); oversubscribing oneDNN with 128 threads is ~80x slower."""
    n = os.cpu_count() or 1
    try:
        n = min(n, len(os.sched_getaffinity(0)))
    except AttributeError:
        pass
    try:
        with open("/sys/fs/cgroup/cpu.max") as f:  # cgroup v2
            quota, period = f.read().split()
        if quota != "max":
            n = min(n, max(1, int(int(quota) / int(period))))
    except (OSError, ValueError):
        try:  # cgroup v1
            with open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us") as f:
                quota = int(f.read())
            with open("/sys/fs/cgroup/cpu/cpu.cfs_period_us") as f:
                period = int(f.read())
            if quota > 0 and period > 0:
                n = min(n, max(1, quota // period))
        except (OSError, ValueError):
            pass
    return n


def pick_cpu_threads(fn) -> int:
    """Give the reference arm its best thread count: calibrate {cores, 2*cores} on a 1/16-size sample."""
    from oracle import li_oracle as O  # input generator only

    cores = host_cores()
    q, d = O.cfg2_inputs(64)
    best, best_t = cores, float("inf")
    for t in sorted({cores, min(2 * cores, os.cpu_count() or cores)}):
        torch.set_num_threads(t)
        fn(q, d)
        t0 = time.perf_counter()
        fn(q, d)
        dt = time.perf_counter() - t0
        if dt < best_t:
            best, best_t = t, dt
        if dt > 5.0:  # a starved host (the sample takes ~50 ms on a healthy one): do not spend the run calibrating
            break
    return best


def cpu_reference_arm(fn, steps: int, warmup: int, threads: int, budget_s: float = 60.0):
    """Times the reference's score_multi_vector (torch.einsum CPU path, processing_utils.py:170-187) on the host
    cores, inputs resident in host memory as bf16.  Returns (queries/s, seconds/step, sample, timed steps)."""
    from oracle import li_oracle as O  # input generator only (seeded cfg2 tensors)

    torch.set_num_threads(threads)
    q, d = O.cfg2_inputs()
    for _ in range(max(1, min(warmup, 1))):
        fn(q, d)
    times, t_start = [], time.perf_counter()
    for _ in range(steps):
        t = time.perf_counter()
        fn(q, d)
        times.append(time.perf_counter() - t)
        if time.perf_counter() - t_start > budget_s:  # bounded sample: a healthy host does a step in < 1 s
            break
    times.sort()
    med = times[len(times) // 2]
    return N_QUERIES / med, med, f"full step (32 queries x 1000 docs x 1030 x 128 bf16) x {len(times)}, median", len(times)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    fn, kind = reference_scorer()
    threads = pick_cpu_threads(fn)
    steps = max(3, min(args.steps, 9))
    qps, sec, sample, steps = cpu_reference_arm(fn, steps, args.warmup, threads)
    path = ("UNMODIFIED colpali_engine from baseline/_ref: BaseVisualRetrieverProcessor.score_multi_vector(qs, ps, "
            "batch_size=128, device='cpu')" if kind == "reference"
            else "oracle port of colpali_engine score_multi_vector (torch.einsum, CPU, batch_size=128)")
    line = {
        "impl": "reference", "metric": METRIC, "value": qps, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
        "warmup": 1, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": dict(CONFIG), "details": {"path": path, "timing": "time.perf_counter around each call, median"},
        "cpu_baseline": {"value": qps, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample,
                         "host_cores": host_cores()},
        "e2e": {"value": qps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def bind_to_gpu_numa_node(local: int):
    """Run this rank (and therefore first-touch its pinned host buffers) on the NUMA node its GPU hangs off.  Round 1:
    eight ranks pushing 264 MB each from buffers placed on one node made the 8-GPU end-to-end leg slower than the
    4-GPU one.  Best effort; returns a short record for the JSON line."""
    try:
        import pynvml

        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(local)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        bus = bus.lower()
        if len(bus.split(":")[0]) == 8:  # NVML pads the PCI domain to 8 hex digits, sysfs uses 4
            bus = bus[4:]
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return {"node": None, "why": "no NUMA information for this device"}
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = set()
            for part in f.read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = cpus & set(os.sched_getaffinity(0))
        if not allowed:
            return {"node": node, "why": "none of the node's CPUs are in this process's affinity mask"}
        os.sched_setaffinity(0, allowed)
        return {"node": node, "cpus": len(allowed)}
    except Exception as e:  # noqa: BLE001
        return {"node": None, "why": repr(e)[:120]}


def run_b200(args):
    import torch.distributed as dist

    import colpali_b200 as cb
    from colpali_b200 import _lib

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # every rank runs on (and first-touches its pinned buffers on) its GPU's NUMA node -- at N = 1 too: an unbound
    # process may sit on the other socket and push its 264 MB across the inter-socket link (5.8-6.3 ms per cold step
    # against 5.0 ms for a bound rank of the N = 2 run).  The affinity is restored before the CPU baseline is timed.
    orig_affinity = os.sched_getaffinity(0)
    numa = bind_to_gpu_numa_node(local)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    # synthetic unit-norm bf16 embeddings generated on device; each rank owns a different shard
    g = torch.Generator(device=dev).manual_seed(1000 + rank)
    gq = torch.Generator(device=dev).manual_seed(0)
    q = torch.nn.functional.normalize(torch.randn(N_QUERIES, N_Q, DIM, device=dev, generator=gq), dim=-1).bfloat16()
    d = torch.nn.functional.normalize(torch.randn(N_DOCS, N_D, DIM, device=dev, generator=g), dim=-1).bfloat16()
    bank = cb.DocBank.from_passages(d, dev)
    qb = cb.QueryBlock(q, dev)
    gathered = torch.empty(world, N_QUERIES, N_DOCS, dtype=torch.float32, device=dev) if world > 1 else None
    fused = None
    if world > 1 and os.environ.get("COLPALI_B200_NCCL_GATHER") != "1":
        from colpali_b200.sharded import FusedGatherScorer

        if FusedGatherScorer.available(dev):  # agreed across ranks
            fused = FusedGatherScorer(N_QUERIES, N_DOCS, dev)
    independent = os.environ.get("COLPALI_B200_DEPENDENT") != "1"  # every step scores the same resident inputs

    def step():
        if fused is not None:
            # all-gather fused into the kernel epilogue (NVSwitch multicast / NVLink peer stores); completion is a
            # per-CTA release-add on every rank's counters, the write-after-read guard an in-kernel wait on them
            return fused.score(qb, bank, independent=independent)
        s = cb.maxsim(qb, bank, independent=independent)
        if world > 1:
            dist.all_gather_into_tensor(gathered.view(world * N_QUERIES, N_DOCS), s)
        return s

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    sync_all()

    # ---- kernel-resident throughput: K steps, CUDA events on the launching stream ----------------
    l0 = _lib.gpu_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local) as clk:
        sync_all()
        e0.record()
        for _ in range(args.steps):
            step()
        if fused is not None:
            fused.wait()  # the region ends when every rank's last slab has arrived here
        e1.record()
        sync_all()
    launches = _lib.gpu_launches() - l0
    ms_total = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms_total, op=dist.ReduceOp.MAX)
    ms_step = float(ms_total) / args.steps
    value = world * N_QUERIES / (ms_step * 1e-3)

    # ---- the dominant kernel alone: roofline numerator -------------------------------------------------------------
    # At N = 1 a step IS one launch of the kernel, so the timed region above is the measurement; with the exchange in
    # the step (N > 1) the kernel is timed again on its own.
    if world == 1:
        ms_kernel = ms_step
    else:
        k_steps = max(10, min(args.steps, 200))
        sync_all()
        e0.record()
        for _ in range(k_steps):
            cb.maxsim(qb, bank, independent=independent)
        e1.record()
        torch.cuda.synchronize()
        ms_kernel = e0.elapsed_time(e1) / k_steps
    # the exchange makes all ranks run at the pace of the slowest GPU (power-limited clocks differ by a few per cent):
    # report that GPU's kernel-alone time next to rank 0's, so the cost of the exchange itself can be read off
    ms_kernel_all = torch.tensor([ms_kernel], device=dev)
    if world > 1:
        dist.all_reduce(ms_kernel_all, op=dist.ReduceOp.MAX)
    ms_kernel_slowest = float(ms_kernel_all)

    # ---- end to end through the public API: host buffers in, host scores out ----------------------------------------
    # N = 1: colpali_b200.score_multi_vector (the reference's scorer signature).  N > 1: the sharded API -- every rank
    # uploads ITS shard and the queries, scores with the fused gather and reads the WHOLE [world, 32, 1000] result back.
    # "cold" re-uploads the bank every step (what the reference does for every call); "warm" keeps the DocBank resident.
    q_host = q.cpu().pin_memory()
    d_host = d.cpu().pin_memory()
    e2e_steps = max(3, min(args.steps, 10))

    def e2e_step(cold: bool):
        if world == 1:
            return cb.score_multi_vector(q_host, d_host if cold else bank, device=dev)
        b = cb.DocBank.from_passages(d_host, dev) if cold else bank
        qq = cb.QueryBlock(q_host, dev)
        if fused is not None:
            view = fused.score(qq, b)
            fused.wait()
            return view.cpu()
        s = cb.maxsim(qq, b)
        dist.all_gather_into_tensor(gathered.view(world * N_QUERIES, N_DOCS), s)
        return gathered.cpu()

    e2e = {}
    for name, cold in (("cold", True), ("warm", False)):
        out = e2e_step(cold)
        sync_all()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            out = e2e_step(cold)
        torch.cuda.synchronize()
        sec = torch.tensor([(time.perf_counter() - t0) / e2e_steps], device=dev)
        if world > 1:
            dist.all_reduce(sec, op=dist.ReduceOp.MAX)
        e2e[name] = world * N_QUERIES / float(sec)
    assert out.device.type == "cpu" and out.shape[-2:] == (N_QUERIES, N_DOCS)

    # ---- self-test carried by the bench run (the driver's GPU test box has one GPU; this is where N > 1 is checked) ---
    selftest = None
    if world > 1:
        n_check = 20
        qbs = [cb.QueryBlock(torch.nn.functional.normalize(
            torch.randn(N_QUERIES, N_Q, DIM, device=dev, generator=gq), dim=-1).bfloat16(), dev) for _ in range(n_check)]
        want = torch.empty(n_check, world, N_QUERIES, N_DOCS, device=dev)
        for i, x in enumerate(qbs):
            dist.all_gather_into_tensor(want[i].view(world * N_QUERIES, N_DOCS), cb.maxsim(x, bank))
        if fused is not None:
            got = torch.empty_like(want)
            for i, x in enumerate(qbs):  # back to back, different queries, no host synchronisation in between
                view = fused.score(x, bank)
                fused.wait()
                got[i].copy_(view)
            torch.cuda.synchronize()
            fused.check_status()
            ok = torch.tensor([int(torch.equal(got, want))], device=dev)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            assert int(ok), "fused all-gather disagrees with NCCL on back-to-back launches"
            selftest = {"fused_gather_equals_nccl_allgather": f"{n_check}/{n_check} back-to-back launches, all ranks",
                        "store_path": "multimem.st (NVSwitch multicast)" if fused.mc_base else "per-peer st.global"}

    if rank == 0:
        pk = peaks()
        achieved = FLOPS_PER_STEP / (ms_kernel * 1e-3) / 1e12
        traffic, traffic_src = None, None
        tpath = os.path.join(ROOT, "profiles", "maxsim_traffic.json")
        if os.path.exists(tpath):
            with open(tpath) as f:
                traffic = json.load(f).get("dram_bytes_per_launch")
            traffic_src = "profiles/maxsim_traffic.json (static: one ncu --set full capture, not measured in this run)"
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": dict(CONFIG),
            "details": {
                "parallelism": (f"corpus-sharded x{world}: 1000 docs/rank, score slabs all-gathered by "
                                + (("NVSwitch multicast stores" if fused.mc_base else "NVLink peer stores")
                                   + " fused into the kernel epilogue, per-CTA completion counters" if fused is not None
                                   else "NCCL all_gather_into_tensor")) if world > 1 else "single GPU",
                "launches": "independent (CPB_FLAG_INDEPENDENT: a step reads nothing the previous step wrote)"
                            if independent else "stream-ordered",
                "l2": "document bank (264 MB) exceeds L2 (126 MB); no explicit flush",
                "timing": "CUDA events on the launching stream, max over ranks",
                "numa": numa,
            },
            "roofline": {
                "bound": "tensor", "achieved": achieved, "peak": pk["bf16_tflops"], "unit": "TFLOP/s",
                "frac": achieved / pk["bf16_tflops"], "traffic": traffic, "traffic_source": traffic_src,
                "peak_source": pk["source"] + " burst",
                "kernel_ms": ms_kernel, "kernel_ms_slowest_rank": ms_kernel_slowest,
                "algorithmic_flops_per_launch": FLOPS_PER_STEP,
                "algorithmic_bytes_per_launch": MIN_BYTES_PER_STEP,
                "hbm_gbs_achieved": MIN_BYTES_PER_STEP / (ms_kernel * 1e-3) / 1e9,
            },
            "e2e": {"value": e2e["cold"], "unit": UNIT,
                    "h2d_bytes_per_step": q_host.numel() * 2 + d_host.numel() * 2,
                    "d2h_bytes_per_step": world * N_QUERIES * N_DOCS * 4, "steps": e2e_steps,
                    "warm_bank_value": e2e["warm"], "warm_bank_h2d_bytes_per_step": q_host.numel() * 2,
                    "api": ("colpali_b200.score_multi_vector(pinned host q, pinned host docs | resident DocBank) -> CPU fp32"
                            if world == 1 else
                            "per rank: DocBank.from_passages(pinned host shard) + QueryBlock(pinned host q) -> "
                            "FusedGatherScorer.score/wait -> gathered [world, 32, 1000] .cpu()")},
            "gpu_launches": launches,
            "clocks": clk.summary(),
        }
        if selftest:
            line["selftest"] = selftest
        if not args.no_cpu and world == 1:
            os.sched_setaffinity(0, orig_affinity)  # the CPU arm gets every core the container may use, not one node's
            fn, kind = reference_scorer()
            threads = pick_cpu_threads(fn)
            qps, sec, sample, _ = cpu_reference_arm(fn, 5, 1, threads)
            line["cpu_baseline"] = {"value": qps, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample,
                                    "seconds_per_step": sec, "host_cores": host_cores()}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
