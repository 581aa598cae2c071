#!/usr/bin/env python
"""bench.py -- encode/decode throughput of the GPUAR hot path on B200 (contract in the task brief).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

A step is one pass of the hot path over one batch of synthetic input.  K encode steps
(model+coder kernel, size scan + compaction) and K decode steps (device packet-chain discovery +
decode kernel) are timed separately with CUDA events on the launching stream, inputs resident in
HBM; L2 is flushed between timed iterations.  `value` is encode GB/s (uncompressed bytes / time,
GB = 1e9 B, whole job over all ranks); decode GB/s is reported beside it under "decode".
`e2e` is the same encode metric through gpuar_b200_compress_host from pinned HOST buffers
(H2D + kernels + D2H inside the timed region); "e2e_decode" likewise.

Workloads: u64m = uniform(0x64, 64 MiB) (BASELINE config 2, the default),
s1g = and3(2, 1 GiB) (config 3), m2g = mixed(3, 2 GiB per rank) (config 4's per-GPU shard at 8 GPUs).
With N > 1 (torchrun) every rank encodes its own packet range of the N-times-larger input
(weak scaling), the ranks' payload totals are exclusive-scanned, and the streams are concatenated
with peer stores over NVLink into W equal segments (segment g on GPU g); the gather of the whole
stream onto one GPU and the no-exchange case are timed beside it.

--impl reference times the reference's own CPU codec (oracle/_ref, compiled from the reference
sources; falls back to the oracle port) on the host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GB = 1e9
WORKLOADS = {
    # name: (generator, seed, bytes per rank, description)
    "u64m": ("uniform", 0x64, 64 << 20, "uniform(0x64, 64 MiB) = data/random_64m.dat stand-in (BASELINE config 2)"),
    "s1g": ("and3", 2, 1 << 30, "and3(2, 1 GiB) low-entropy (BASELINE config 3)"),
    "m2g": ("mixed", 3, 2 << 30, "mixed(3, 2 GiB per rank) (BASELINE config 4 shard)"),
    "m4g": ("mixed", 7, 4 << 30, "mixed(7, 4 GiB) (BASELINE config 5)"),
    "m16g": ("mixed", 3, 16 << 30, "mixed(3, 16 GiB) (BASELINE config 4, whole job on one rank)"),
    # tuning only (where the warp-specialised encoder hands over to the lane=packet one)
    "u128m": ("uniform", 0x65, 128 << 20, "uniform(0x65, 128 MiB) (tuning)"),
    "u192m": ("uniform", 0x66, 192 << 20, "uniform(0x66, 192 MiB) (tuning)"),
    "u256m": ("uniform", 0x67, 256 << 20, "uniform(0x67, 256 MiB) (tuning)"),
    "u384m": ("uniform", 0x68, 384 << 20, "uniform(0x68, 384 MiB) (tuning)"),
}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def measured_traffic(workload, kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full capture of
    this workload (profiles/traffic.json), or None if that workload has not been captured."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))[workload][kernel]
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks/throttle reasons while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def wait_first(self, timeout=3.0):
        """Blocks until nvidia-smi has printed its first row; rows up to here (idle GPU) are dropped."""
        t_end = time.perf_counter() + timeout
        while self.proc and not self.rows and time.perf_counter() < t_end:
            time.sleep(0.005)
        self.t_load = time.perf_counter()

    def mark(self):
        """Start of the timed region: samples before it (warm-up, same kernels) are kept apart."""
        self.t_mark = time.perf_counter()

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, inside = [], None, set(), 0
        for t, r in self.rows:
            if len(r) < 6 or t < getattr(self, "t_load", 0.0):
                continue
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
            except ValueError:
                continue
            inside += t >= getattr(self, "t_mark", 0.0)
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        # the sampler runs from the warm-up on (the same kernels, back to back); `samples_timed` of them
        # fall inside the timed region, which at 64 MiB lasts only a few tens of milliseconds
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm), "samples_timed": inside}


# =============================================================== reference arm
def run_reference(args):
    """The reference's own CPU codec on the host cores (rank 0 only)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import _oracle as O
    from gpuar_b200 import datagen as D

    gen, seed, nbytes, desc = WORKLOADS[args.workload]
    cores = os.cpu_count() or 1
    # bounded sample of the workload: the reference codes ~9 MB/s per core
    sample = min(nbytes, 64 << 20)
    data = D.GENERATORS[gen](seed, sample)
    if O.have_ref():
        kind = "reference"
        enc = lambda: O.ref_encode(data, cores)
        dec = lambda pay: O.ref_decode(pay, data.size, cores)
    else:
        kind = "port"
        cores = 1
        enc = lambda: O.encode(data)
        dec = lambda pay: O.decode(pay)
    pay = enc()
    for _ in range(max(0, min(args.warmup, 1))):
        enc()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        pay = enc()
    t_enc = (time.perf_counter() - t0) / args.steps
    t0 = time.perf_counter()
    for _ in range(args.steps):
        back = dec(pay)
    t_dec = (time.perf_counter() - t0) / args.steps
    assert np.array_equal(back, data)
    v = sample / t_enc / GB
    line = {
        "impl": "reference", "metric": "encode_GBps", "value": v, "unit": "GB/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_enc * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u16", "data": "synthetic",
        "config": {"workload": args.workload, "desc": desc, "packet_bytes": 8192},
        "decode": {"value": sample / t_dec / GB, "unit": "GB/s", "ms_per_step": t_dec * 1e3},
        "cpu_baseline": {"value": v, "unit": "GB/s", "cores": cores, "kind": kind,
                         "sample": f"first {sample >> 20} MiB of the workload per step, packets partitioned over "
                                   f"{cores} host threads (the reference itself is single-threaded)"},
        "e2e": {"value": v, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ==================================================================== our arm
def cpu_baseline(workload):
    """Reference CPU codec on this box's host cores, bounded sample (rank 0, N=1 only)."""
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import _oracle as O
    from gpuar_b200 import datagen as D
    gen, seed, nbytes, _ = WORKLOADS[workload]
    cores = os.cpu_count() or 1
    sample = min(nbytes, 64 << 20)
    data = D.GENERATORS[gen](seed, sample)
    if O.have_ref():
        kind = "reference"
        t0 = time.perf_counter(); pay = O.ref_encode(data, cores); t_enc = time.perf_counter() - t0
        t0 = time.perf_counter(); back = O.ref_decode(pay, data.size, cores); t_dec = time.perf_counter() - t0
    else:
        kind, cores = "port", 1
        sample = min(sample, 8 << 20)
        data = data[:sample]
        t0 = time.perf_counter(); pay = O.encode(data); t_enc = time.perf_counter() - t0
        t0 = time.perf_counter(); back = O.decode(pay); t_dec = time.perf_counter() - t0
    assert np.array_equal(back, data)
    out = {"value": sample / t_enc / GB, "unit": "GB/s", "cores": cores, "kind": kind,
           "decode_value": sample / t_dec / GB,
           "sample": f"first {sample >> 20} MiB of the workload, one pass, packets partitioned over {cores} host "
                     f"threads ({(t_enc + t_dec) * cores:.1f} core-seconds)"}
    if kind == "reference":
        out["reference_gpu_kernel"] = reference_gpu_kernel(data, pay)
    return out, pay


def reference_gpu_kernel(data, pay):
    """The reference's ORIGINAL CUDA kernels (garCompress / garDecompress, gpuar_kernel.cu:894-949, compiled
    unmodified for sm_100 into oracle/_ref) on this GPU, device-resident slots, timed from launch to
    cudaDeviceSynchronize as the reference times itself.  Reported beside our numbers as north_star asks."""
    import numpy as np
    import torch
    import _oracle as O
    try:
        n = data.size
        packets = (n + 8191) // 8192
        src = torch.from_numpy(data).cuda()
        slots = torch.zeros(packets * 8704, dtype=torch.uint8, device="cuda")
        back = torch.zeros(packets * 8192, dtype=torch.uint8, device="cuda")
        torch.cuda.synchronize()
        ref = O.ref()
        ref.gpuar_ref_gpu_init()
        times = []
        for fn, args in ((ref.gpuar_ref_gpu_encode, (src.data_ptr(), n, slots.data_ptr())),
                         (ref.gpuar_ref_gpu_decode, (slots.data_ptr(), packets, back.data_ptr()))):
            fn(*args)                                           # warm-up
            assert ref.gpuar_ref_gpu_sync() == 0
            best = 1e9
            for _ in range(3):                                  # launch .. cudaDeviceSynchronize, like the reference's
                t0 = time.perf_counter()                        # own process_timer (gpu_compressor.cpp:184-194)
                fn(*args)
                assert ref.gpuar_ref_gpu_sync() == 0
                best = min(best, time.perf_counter() - t0)
            times.append(best)
        ok = bool(torch.equal(back[:n], src))
        return {"encode_GBps": n / times[0] / GB, "decode_GBps": n / times[1] / GB, "round_trip": ok,
                "what": "reference garCompress/garDecompress kernels, 8704-byte slots, no compaction, no index"}
    except Exception as e:                                      # never let the baseline break the bench line
        return {"error": repr(e)}


def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from gpuar_b200 import _lib, codec, datagen as D
    from gpuar_b200.shard import ShardedCodec

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run"
    torch.cuda.set_device(local)
    if world > 1:
        # NCCL announces its version on stdout at communicator creation: keep stdout for the JSON line
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    dev = codec.DeviceCodec(local)
    _lib.set_option(_lib.OPT_ENCODE_PATH, {"auto": 0, "fused": 1, "ws": 2}[args.encode_path])
    if args.work_unit:
        _lib.set_option(_lib.OPT_COMPACT_TILE, args.work_unit)
    sharded = ShardedCodec(dev, rank, world, "segments")
    gathered = ShardedCodec(dev, rank, world, "gather") if world > 1 else None

    gen, seed, nbytes, desc = WORKLOADS[args.workload]
    # weak scaling: the job is `world` times the per-rank workload; rank r owns packets [r*P, (r+1)*P)
    start = rank * nbytes
    x = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    gen_dev = {"uniform": D.uniform_device, "and3": D.and3_device, "mixed": D.mixed_device}[gen]
    for a in range(0, nbytes, 1 << 29):                            # 512 MiB at a time: bounded temporaries
        b = min(nbytes, a + (1 << 29))
        x[a:b] = gen_dev(seed, b - a, start + a)
    packet = args.packet
    packets = (nbytes + packet - 1) // packet
    cap = codec.payload_bound(nbytes, packet)
    payload = torch.empty(cap + 16, dtype=torch.uint8, device="cuda")
    total = torch.zeros(1, dtype=torch.int64, device="cuda")
    offsets = torch.empty(packets + 1, dtype=torch.int64, device="cuda")
    result = torch.zeros(4, dtype=torch.int64, device="cuda")
    out = torch.empty(packets * packet, dtype=torch.uint8, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")       # > 126 MB L2
    sharded.reserve(cap)
    if gathered:
        gathered.reserve(cap)

    def enc_step():
        dev.encode(x, payload, total, packet=packet)
        sharded.concat(payload, total)            # N > 1: totals scan + NVLink peer copies; N = 1: nothing

    c_holder = [0]

    def dec_step():
        dev.index(payload, c_holder[0], packets, offsets, result, packet=packet)
        dev.decode(payload, c_holder[0], offsets, packets, out, packet=packet)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(step, k):
        """K steps, each bracketed by CUDA events on the launching stream, L2 flushed in between."""
        evs = []
        for _ in range(k):
            flush.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            step()
            b.record()
            evs.append((a, b))
        torch.cuda.synchronize()
        return sum(a.elapsed_time(b) for a, b in evs) / 1e3

    # ---- warm-up (>= 3) and correctness gate: nothing is reported unless the round trip holds
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        sampler.wait_first()                      # nvidia-smi is looping from here on
    for _ in range(max(3, args.warmup)):
        enc_step()
    torch.cuda.synchronize()
    c = int(total.item())
    c_holder[0] = c
    for _ in range(max(3, args.warmup)):
        dec_step()
    # a 64 MiB step lasts ~3 ms: keep the same kernels running (untimed) for ~0.2 s so that the sampler
    # sees the load; the count depends on the size only, so every rank does the same (enc_step is collective)
    for _ in range(max(3, min(100, int(0.2 / (nbytes / 30e9))))):
        enc_step()
        dec_step()
    torch.cuda.synchronize()
    torch.cuda.synchronize()
    assert [int(v) for v in result.tolist()[:3]] == [packets, nbytes, 0], result.tolist()
    assert torch.equal(out[:nbytes], x), "decode(encode(x)) != x"

    sampler.mark()
    launches0 = _lib.launch_count()
    _lib.profile(True)
    _lib.profile_read()
    barrier()
    t_enc = timed(enc_step, args.steps)
    barrier()
    t_local = timed(lambda: dev.encode(x, payload, total, packet=packet), args.steps) if world > 1 else t_enc
    barrier()

    def gather_step():
        dev.encode(x, payload, total, packet=packet)
        gathered.concat(payload, total)

    t_gather = t_enc
    if gathered:
        for _ in range(3):
            gather_step()
        barrier()
        t_gather = timed(gather_step, args.steps)
        barrier()
    t_dec = timed(dec_step, args.steps)
    barrier()
    spans = _lib.profile_read()
    _lib.profile(False)
    launches = _lib.launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None

    # ---- e2e through the host-buffer entry points, pinned host memory
    host_in = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    host_in.copy_(x)
    host_gip = torch.empty(20 + codec.payload_bound(nbytes), dtype=torch.uint8).pin_memory()   # e2e: 8192-byte packets
    host_out = torch.empty(nbytes + 8192, dtype=torch.uint8).pin_memory()
    np_in, np_gip, np_out = host_in.numpy(), host_gip.numpy(), host_out.numpy()
    e2e_steps = max(1, min(args.steps, 5))
    g = codec.compress(np_in, out=np_gip)
    codec.decompress(g, out=np_out)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        g = codec.compress(np_in, out=np_gip)
    torch.cuda.synchronize()
    t_e2e_enc = (time.perf_counter() - t0) / e2e_steps
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        back = codec.decompress(g, out=np_out)
    torch.cuda.synchronize()
    t_e2e_dec = (time.perf_counter() - t0) / e2e_steps
    assert back.size == nbytes and np.array_equal(back[:4096], np_in[:4096]) and np.array_equal(back[-4096:], np_in[-4096:])
    gip_bytes = int(g.size)

    # ---- max over ranks
    times = torch.tensor([t_enc, t_dec, t_e2e_enc, t_e2e_dec, t_local, t_gather], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    t_enc, t_dec, t_e2e_enc, t_e2e_dec, t_local, t_gather = (float(v) for v in times.tolist())

    if rank == 0:
        job = nbytes * world
        peak, peak_src = measured_peaks()
        enc_ms, enc_calls = spans["encode"]
        alg_bytes = nbytes + c                                     # N read + C written per launch
        achieved = alg_bytes / (enc_ms / 1e3 / max(1, enc_calls)) / GB
        line = {
            "metric": "encode_GBps", "value": job / (t_enc / args.steps) / GB, "unit": "GB/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
            "ms_per_step": t_enc / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u16", "data": "synthetic",
            "config": {"workload": args.workload, "desc": desc, "bytes_per_rank": nbytes, "packet_bytes": packet,
                       "payload_bytes_rank0": c, "l2": "flushed between timed iterations (256 MiB write)",
                       "parallelism": f"packet-range shards x{world}" if world > 1 else "single GPU",
                       "encode_path": args.encode_path, "work_unit_packets": args.work_unit or "auto"},
            "encode_shards_in_place": {"value": job / (t_local / args.steps) / GB, "unit": "GB/s",
                                       "note": "encode only, every shard's payload left on its own GPU (SURVEY 8e)"},
            "encode_gather_one_gpu": {"value": job / (t_gather / args.steps) / GB, "unit": "GB/s",
                                      "note": "concatenation of the whole stream into rank 0's memory: ingress-limited "
                                              "on that GPU (SURVEY 8e); `value` concatenates into W equal segments, "
                                              "segment g on GPU g, so every GPU receives total/W bytes"},
            "decode": {"value": job / (t_dec / args.steps) / GB, "unit": "GB/s",
                       "ms_per_step": t_dec / args.steps * 1e3,
                       "includes": "device packet-chain discovery + decode kernel"},
            "e2e": {"value": job / t_e2e_enc / GB, "unit": "GB/s", "h2d_bytes_per_step": nbytes,
                    "d2h_bytes_per_step": gip_bytes - 20, "api": "gpuar_b200_compress_host (pinned host buffers)"},
            "e2e_decode": {"value": job / t_e2e_dec / GB, "unit": "GB/s", "h2d_bytes_per_step": gip_bytes - 20,
                           "d2h_bytes_per_step": nbytes, "api": "gpuar_b200_decompress_host"},
            "gpu_launches": launches,
            "kernels_ms_per_step": {k: (v[0] / max(1, v[1])) for k, v in spans.items()},
            "roofline": {"bound": "hbm", "kernel": "encode kernel (encode_ws_kernel up to one wave, else encode_kernel)",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": measured_traffic(args.workload, "encode"), "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg_bytes,
                         "note": "integer-latency-bound path: see profiles/ for issue-slot and occupancy evidence"},
            "roofline_decode": {"bound": "hbm", "kernel": "decode_kernel",
                                "achieved": alg_bytes / (spans["decode"][0] / 1e3 / max(1, spans["decode"][1])) / GB,
                                "peak": peak, "unit": "GB/s",
                                "frac": alg_bytes / (spans["decode"][0] / 1e3 / max(1, spans["decode"][1])) / GB / peak,
                                "traffic": measured_traffic(args.workload, "decode"),
                                "algorithmic_bytes_per_launch": alg_bytes},
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu_baseline:
            base, _ = cpu_baseline(args.workload)
            line["cpu_baseline"] = base
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="u64m", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--packet", type=int, default=8192,
                    help="raw bytes per packet for the device-resident numbers (8192 = the reference format; "
                         "4096/12288/16112 for the BASELINE config-5 sweep; e2e and CPU baseline stay at 8192)")
    ap.add_argument("--encode-path", default="auto", choices=["auto", "fused", "ws"],
                    help="encoder kernel: auto (by packet count), fused lane=packet, warp-specialised")
    ap.add_argument("--work-unit", type=int, default=0,
                    help="packets per CTA of the scan + compaction kernel (0 = auto; 4..128 = 32 KiB..1 MiB of "
                         "input: the work-unit half of the BASELINE config-5 sweep)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
