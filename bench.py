#!/usr/bin/env python3
"""
bench.py -- reads/sec of the adapter-trimming hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W]          this repo's CUDA path
    python bench.py --impl reference ...                         the reference's CPU path

Workload (config.workload): BASELINE.json configs[1]: 100 M x 150 bp single-end synthetic reads,
one 3' adapter AGATCGGAAGAGC, e=0.1 -- per GPU (weak scaling: every rank trims its own shard,
no data-path collective; the trim statistics are all-reduced once per step).

One "step" = one pass of the hot path over the whole per-GPU batch.
  value  : whole-job reads/s with the batch resident in HBM (CUDA events, max over ranks).
  e2e    : the same through the host-facing C ABI (cg_process_batch): pinned host buffers, the
           library's host-side packing (reads travel partly as a 3-bases-per-byte stream), H2D
           copies, kernels, D2H copy of the match records all inside the timed region; the byte
           counts come from the library (cg_ctx_transfer_bytes), raw_transfer_value is the same
           call with CUTADAPT_B200_H2D_PACK=0.
  roofline.achieved : algorithmic bytes (190 B/read: 150 sequence + 8 offset + 32 result)
           x reads per pass / mean device time of ALL kernels of the trimming pass (scan,
           plan, DP rounds), measured with CUDA events on the launching stream inside the
           library -- i.e. the whole hot path, not just its largest kernel.
  cpu_baseline : the reference's own compiled hot path (oracle/_ref: Adapter.match_to +
           Match.trimmed) on the host cores, bounded sample, rank 0 at N=1 only.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ADAPTER = "AGATCGGAAGAGC"
ERROR_RATE = 0.1
READ_LEN = 150
ALGO_BYTES_PER_READ = READ_LEN + 8 + 32
NCU_DRAM_BYTES_PER_READ = 380    # measured once per round with ncu (profiles/README.md), not in the timed run
DEFAULT_READS = 100_000_000
HOST_WINDOW_READS = 20_000_000           # pinned host window streamed repeatedly in the e2e leg


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference (oracle/_ref) or, if it did not travel, the C oracle port
# ------------------------------------------------------------------------------------------------

_WORKER = {}


def _cpu_init(kind, path, index, n_workers):
    """Pool initializer: every worker loads its slice of the sample once and builds the adapter."""
    with open(path) as f:
        reads = f.read().split("\n")
    reads = [r for r in reads if r]
    per = max(1, len(reads) // n_workers)
    with index.get_lock():
        me = index.value
        index.value += 1
    _WORKER["reads"] = reads[me * per:(me + 1) * per] or reads[:per]
    if kind == "reference":
        sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
        from cutadapt.adapters import BackAdapter

        _WORKER["adapter"] = BackAdapter(ADAPTER, max_errors=ERROR_RATE, min_overlap=3)
    else:
        from oracle import oracle as O
        from cutadapt_b200.kmer_heuristic import create_positions_and_kmers

        _WORKER["oracle"] = O
        _WORKER["tables"] = O.KmerTables(create_positions_and_kmers(ADAPTER, 3, ERROR_RATE, True, False))
    _WORKER["kind"] = kind


def _cpu_worker(repeat):
    """The reference hot path per read: Adapter.match_to (prefilter + locate) and Match.trimmed."""
    import time as _t

    reads = _WORKER["reads"]
    kept = 0
    t0 = _t.perf_counter()
    if _WORKER["kind"] == "reference":
        adapter = _WORKER["adapter"]
        for _ in range(repeat):
            for read in reads:
                m = adapter.match_to(read)
                if m is not None:
                    kept += len(m.trimmed(read))
    else:
        O, kt = _WORKER["oracle"], _WORKER["tables"]
        for _ in range(repeat):
            for read in reads:
                if kt.present(read):
                    r = O.locate(ADAPTER, read, ERROR_RATE, 14, min_overlap=3)
                    if r is not None:
                        kept += r[2]
    return _t.perf_counter() - t0, len(reads) * repeat, kept


def reference_kind():
    ref_dir = os.path.join(ROOT, "oracle", "_ref", "cutadapt")
    if os.path.isdir(ref_dir) and any(f.startswith("_align") and f.endswith(".so") for f in os.listdir(ref_dir)):
        return "reference"
    return "port"


def cpus_available():
    """CPUs this process may use: the affinity mask, cut down by a cgroup CPU quota (containers on a shared host)."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        if quota != "max":
            n = min(n, max(1, round(int(quota) / int(period))))
    except Exception:
        try:
            q = int(open("/sys/fs/cgroup/cpu/cpu.cfs_quota_us").read())
            p = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
            if q > 0 and p > 0:
                n = min(n, max(1, round(q / p)))
        except Exception:
            pass
    return n


class CpuArm:
    """All host cores running the reference hot path on pre-parsed reads (spawned workers)."""

    def __init__(self, sample_reads, cores=None):
        import multiprocessing as mp
        import tempfile

        self.kind = reference_kind()
        self.cores = cores or len(os.sched_getaffinity(0))
        self.cores = max(1, min(self.cores, len(sample_reads)))
        fd, self.path = tempfile.mkstemp(suffix=".reads")
        with os.fdopen(fd, "w") as f:
            f.write("\n".join(sample_reads))
        ctx = mp.get_context("spawn")
        index = ctx.Value("i", 0)
        self.pool = ctx.Pool(self.cores, initializer=_cpu_init, initargs=(self.kind, self.path, index, self.cores))
        # warm up + calibrate: one pass over each worker's slice
        t0 = time.perf_counter()
        res = self.pool.map(_cpu_worker, [1] * self.cores, chunksize=1)
        self.pass_reads = sum(r[1] for r in res)
        self.pass_seconds = max(r[0] for r in res)

    def run(self, seconds_target):
        # calibrate on a warm, multi-pass run (a single pass is dominated by stragglers / dispatch)
        if not getattr(self, "rate", None):
            t0 = time.perf_counter()
            res = self.pool.map(_cpu_worker, [4] * self.cores, chunksize=1)
            self.rate = sum(r[1] for r in res) / (time.perf_counter() - t0)
        # timed rounds until the sample is long enough (the rate estimate improves with every round)
        total_n, total_wall = 0, 0.0
        while total_wall < 0.8 * seconds_target:
            remaining = seconds_target - total_wall
            repeat = max(1, int(round(remaining * self.rate / max(self.pass_reads, 1))))
            t0 = time.perf_counter()
            res = self.pool.map(_cpu_worker, [repeat] * self.cores, chunksize=1)
            wall = time.perf_counter() - t0
            n = sum(r[1] for r in res)
            self.rate = n / wall
            total_n += n
            total_wall += wall
        return total_n / total_wall, total_n, total_wall

    def close(self):
        self.pool.close()
        self.pool.join()
        try:
            os.unlink(self.path)
        except OSError:
            pass


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------

class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.samples = []
        self._stop = threading.Event()
        self._thread = None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(
                    ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits"],
                    capture_output=True, text=True, timeout=5).stdout.strip().split("\n")[0]
                parts = [p.strip() for p in out.split(",")]
                if len(parts) >= 6:
                    self.samples.append(parts)
            except Exception:
                pass
            self._stop.wait(0.2)

    def start(self):
        self._thread = threading.Thread(target=self._run, daemon=True)
        self._thread.start()

    def stop(self):
        self._stop.set()
        if self._thread:
            self._thread.join(timeout=6)

    def summary(self):
        sm, mx, reasons = [], 0, set()
        for p in self.samples:
            try:
                sm.append(float(p[0])); mx = max(mx, float(p[1]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------

def numa_nodes():
    """{node: set of CPUs this process may use} for the NUMA nodes of the host."""
    nodes = {}
    try:
        allowed = os.sched_getaffinity(0)
        for entry in sorted(os.listdir("/sys/devices/system/node")):
            if not entry.startswith("node") or not entry[4:].isdigit():
                continue
            cpus = set()
            for part in open(f"/sys/devices/system/node/{entry}/cpulist").read().strip().split(","):
                if part:
                    a, _, b = part.partition("-")
                    cpus.update(range(int(a), int(b or a) + 1))
            if cpus & allowed:
                nodes[int(entry[4:])] = cpus & allowed
    except Exception:
        return {}
    return nodes


def best_numa_node_for(dev, torch):
    """
    The NUMA node whose memory the GPU reads fastest, MEASURED: pinned buffers are allocated by a thread running on
    each node in turn (first touch puts them there) and copied to the device.  A deployment would place its read
    buffers like this; the topology files are not reliable inside containers.  Returns (node or None, {node: GB/s}).
    """
    nodes = numa_nodes()
    if len(nodes) < 2:
        return None, {}
    saved = os.sched_getaffinity(0)
    rates = {}
    try:
        dst = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        for node, cpus in nodes.items():
            os.sched_setaffinity(0, cpus)
            src = torch.empty(256 << 20, dtype=torch.uint8, pin_memory=True)
            src.fill_(65)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            dst.copy_(src, non_blocking=True)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(4):
                dst.copy_(src, non_blocking=True)
            e1.record()
            torch.cuda.synchronize()
            rates[node] = 4 * src.numel() / (e0.elapsed_time(e1) / 1000.0) / 1e9
            del src
    except Exception:
        rates = {}
    finally:
        os.sched_setaffinity(0, saved)
    if not rates:
        return None, {}
    return max(rates, key=rates.get), {k: round(v, 1) for k, v in rates.items()}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--reads", type=int, default=int(os.environ.get("BENCH_READS", DEFAULT_READS)),
                    help="reads per GPU (default: the 100 M of BASELINE configs[1])")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    workload = f"{args.reads}x{READ_LEN}bp SE synthetic per GPU, one 3' adapter {ADAPTER}, e={ERROR_RATE}"

    if args.impl == "reference":
        if rank != 0:
            return 0
        from cutadapt_b200.synth import make_reads

        sample, _ = make_reads(400_000, config=2)
        arm = CpuArm(sample)
        kind, cores = arm.kind, arm.cores
        vals = []
        for i in range(args.warmup + args.steps):
            v, n, wall = arm.run(5.0)
            if i >= args.warmup:
                vals.append((v, n, wall))
        arm.close()
        value = sum(n for _, n, _ in vals) / sum(w for _, _, w in vals)
        line = {
            "impl": "reference", "metric": "reads/sec (150bp SE, 1 adapter, e=0.1)", "value": value, "unit": "reads/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1000.0 * sum(w for _, _, w in vals) / len(vals), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": workload, "sample": f"{vals[0][1]} reads per step over {cores} host cores"},
            "cpu_baseline": {"value": value, "unit": "reads/s", "cores": cores, "kind": kind,
                             "cpus_available": cpus_available(),
                             "sample": f"{vals[0][1]} pre-parsed reads per step, Adapter.match_to + Match.trimmed"},
            "e2e": {"value": value, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }
        print(json.dumps(line))
        return 0

    import numpy as np
    import torch
    import torch.distributed as dist

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    from cutadapt_b200 import _lib
    from cutadapt_b200.adapters import BackAdapter, MultipleAdapters
    from cutadapt_b200.pipeline import DeviceBatch, allreduce_statistics
    from cutadapt_b200.synth import make_read_tensor

    n = args.reads
    # CPU baseline first (spawned workers; rank 0, N=1 only), before the GPU is busy
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        sample_t, _ = make_read_tensor(400_000, config=2, shard=0, device="cpu")
        raw = sample_t.numpy().tobytes()
        sample = [raw[i * READ_LEN:(i + 1) * READ_LEN].decode() for i in range(sample_t.shape[0])]
        arm = CpuArm(sample)
        v, nproc, wall = arm.run(12.0)
        arm.close()
        cpu = {"value": v, "unit": "reads/s", "cores": arm.cores, "kind": arm.kind, "cpus_available": cpus_available(),
               "sample": f"{nproc} pre-parsed reads in {wall:.1f}s: Adapter.match_to + Match.trimmed over {arm.cores} processes"}

    seq, _ = make_read_tensor(n, config=2, shard=rank, device=str(dev))
    seq = seq.reshape(-1)
    offsets = torch.arange(n + 1, device=dev, dtype=torch.int64) * READ_LEN
    adapters = MultipleAdapters([BackAdapter(ADAPTER, max_errors=ERROR_RATE, min_overlap=3, name="adapter")])
    batch = DeviceBatch(adapters, device=local_rank)
    out = torch.empty((n, 8), dtype=torch.int32, device=dev)
    stats = torch.zeros(int(_lib.lib().cg_stats_size(1, READ_LEN, 3)), dtype=torch.int64, device=dev)

    def step():
        res = batch.run(seq, offsets, None, max_read_len=READ_LEN, out=out)
        stats.zero_()
        batch.statistics(res, READ_LEN, 3, into=stats)
        allreduce_statistics(stats)
        return res

    for _ in range(args.warmup):
        step()
    barrier()
    batch.ctx.kernel_time(reset=True)
    if os.environ.get("CUTADAPT_B200_STAGE_TIMES"):
        batch.ctx.stage_times(reset=True)
    launches0 = batch.ctx.launch_count()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    barrier()
    elapsed_ms = ev0.elapsed_time(ev1)
    launches = batch.ctx.launch_count() - launches0
    kern_ms, kern_n = batch.ctx.kernel_time(reset=True)
    stage_ms = batch.ctx.stage_times(reset=True) if os.environ.get("CUTADAPT_B200_STAGE_TIMES") else None
    t = torch.tensor([elapsed_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms = float(t.item())
    with_adapters = int(stats[2].item())
    total_reads_stat = int(stats[0].item())
    value = n * world * args.steps / (elapsed_ms / 1000.0)

    # ---- end-to-end through the host-facing C ABI ------------------------------------------------
    e2e = None
    if not args.no_e2e:
        hw = min(n, HOST_WINDOW_READS)
        passes = (n + hw - 1) // hw
        # host buffers on the NUMA node the GPU reads fastest (measured), like a deployment would place them
        gpu_node, node_rates = best_numa_node_for(dev, torch)
        saved_affinity = os.sched_getaffinity(0)
        if gpu_node is not None:
            os.sched_setaffinity(0, numa_nodes()[gpu_node])
        h_seq = torch.empty(hw * READ_LEN, dtype=torch.uint8, pin_memory=True)
        h_seq.copy_(seq[: hw * READ_LEN])
        h_off = torch.empty(hw + 1, dtype=torch.int64, pin_memory=True)
        h_off.copy_(offsets[: hw + 1])
        h_out = torch.empty((hw, 8), dtype=torch.int32, pin_memory=True)
        h_out.zero_()
        os.sched_setaffinity(0, saved_affinity)
        # worker threads of the library's host side (packing for the compressed transfer): this rank's
        # share of the host cores
        if world > 1:
            os.environ.setdefault("CUTADAPT_B200_HOST_THREADS",
                                  str(max(2, min(32, _lib.lib().cg_host_cpus_available() // world))))
        host_threads = int(_lib.lib().cg_host_threads())
        host_ctx = _lib.Context(local_rank)
        host_set = _lib.AdapterSet(batch.spec, host_ctx)
        params = batch.params
        import ctypes as C

        def e2e_step():
            done = 0
            for p in range(passes):
                cnt = min(hw, n - done)
                _lib.check(_lib.lib().cg_process_batch(
                    host_ctx.handle, host_set.handle, h_seq.data_ptr(), None, h_off.data_ptr(), cnt,
                    C.byref(params), h_out.data_ptr(), None))
                done += cnt

        e2e_step()
        # for the record: one step with the raw (uncompressed) transfer
        os.environ["CUTADAPT_B200_H2D_PACK"] = "0"
        e2e_step()
        barrier()
        t0 = time.perf_counter()
        e2e_step()
        barrier()
        raw_wall = time.perf_counter() - t0
        os.environ["CUTADAPT_B200_H2D_PACK"] = "1"
        e2e_step()             # lets the compressed share of the transfer settle
        e2e_step()
        barrier()
        host_ctx.transfer_bytes(reset=True)
        host_ctx.host_profile(reset=True)
        l0 = host_ctx.launch_count()
        t0 = time.perf_counter()
        e2e_steps = max(1, min(args.steps, 3))
        for _ in range(e2e_steps):
            e2e_step()
        barrier()
        wall = time.perf_counter() - t0
        tw = torch.tensor([wall], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tw, op=dist.ReduceOp.MAX)
        wall = float(tw.item())
        h2d_bytes, d2h_bytes = host_ctx.transfer_bytes()
        e2e = {"value": n * world * e2e_steps / wall, "unit": "reads/s",
               # counted by the library from the copies it issues (cg_ctx_transfer_bytes): the reads travel as a
               # base-6 stream, three characters per byte, packed by the library's host threads inside the timed
               # region and expanded to the caller's bytes on the device; the offsets of equally long reads are
               # regenerated on the device from (first offset, length)
               "h2d_bytes_per_step": h2d_bytes // e2e_steps, "d2h_bytes_per_step": d2h_bytes // e2e_steps,
               "steps": e2e_steps, "launches": host_ctx.launch_count() - l0,
               "host_threads": host_threads, "host_cpus_available": int(_lib.lib().cg_host_cpus_available()),
               "host_numa_node": int(_lib.lib().cg_ctx_numa_node(host_ctx.handle)), "host_buffer_node": gpu_node, "h2d_GBps_by_node": node_rates,
               "host_profile": {k: (round(v, 4) if isinstance(v, float) else v)
                                for k, v in host_ctx.host_profile().items()},
               "raw_transfer_value": n * world / raw_wall,
               "how": f"cg_process_batch on pinned host buffers (sequences + int64 offsets in, 32-byte records out), "
                      f"compressed host-to-device transfer (raw_transfer_value: the same with "
                      f"CUTADAPT_B200_H2D_PACK=0, {n * READ_LEN} B in per step); "
                      f"the {n}-read step streams a {hw}-read pinned window {passes}x"}
    sampler.stop()

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        per_launch_ms = kern_ms / max(kern_n, 1)
        achieved = ALGO_BYTES_PER_READ * n / (per_launch_ms / 1000.0) / 1e9 if kern_n else None
        line = {
            "metric": "reads/sec (150bp SE, 1 adapter, e=0.1)", "value": value, "unit": "reads/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": workload, "reads_per_gpu": n, "l2": "inputs (15 GB/GPU) larger than L2",
                       "step": "trim pipeline (scan, plan, DP rounds) + statistics reduction + int64 all-reduce of the statistics",
                       "with_adapters": with_adapters, "reads_counted": total_reads_stat},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": (achieved / peak) if achieved else None,
                         # ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum of the pipeline's kernels at
                         # 4 M reads (scan 749 MB, plan 615 MB, run rounds ~155 MB; profiles/README.md), scaled to
                         # the reads of one pass: bytes per launch.  The plan and run kernels re-read the windows
                         # of the reads that pass the prefilter, hence ~2x the algorithmic bytes.
                         "traffic": NCU_DRAM_BYTES_PER_READ * n,
                         "traffic_unit": "bytes per pass (ncu DRAM bytes per read x reads)",
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s",
                         "kernel": "trim pipeline: cg_scan_kernel -> cg_list_kernel<plan> -> 4x cg_list_kernel<run> "
                                   "(all kernels of one pass; per-kernel shares in profiles/)",
                         "kernel_ms_per_launch": per_launch_ms,
                         "stage_ms_per_launch": ({k: v / max(kern_n, 1) for k, v in stage_ms.items()} if stage_ms else None),
                         "algorithmic_bytes_per_read": ALGO_BYTES_PER_READ},
            "cpu_baseline": cpu,
            "e2e": e2e,
            "gpu_launches": launches,
            "clocks": sampler.summary(),
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
