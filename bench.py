#!/usr/bin/env python
"""bench.py -- encode/decode throughput of the GPUAR hot path on B200 (contract in the task brief).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME] [--sub A,B]

A step is one pass of the hot path over one batch of synthetic input.  K encode steps
(model+coder kernel, size scan + compaction) and K decode steps (device packet-chain discovery +
decode kernel) are timed separately with CUDA events on the launching stream, inputs resident in
HBM; L2 is flushed between timed iterations.  `value` is encode GB/s (uncompressed bytes / time,
GB = 1e9 B, whole job over all ranks); decode GB/s is reported beside it under "decode".
`e2e` is the same encode metric from pinned HOST buffers through the library's host entry points
(H2D + kernels + D2H inside the timed region); "e2e_decode" likewise.

Workloads.  The headline (default) is BASELINE config 4, m16g = mixed(3, 16 GiB): ONE
16 GiB input at every N, sharded by packet range over the N GPUs (strong scaling; it fits one B200).
The other configurations ride in the same JSON line under "sub": u64m = uniform(0x64, 64 MiB)
(config 2) and s1g = and3(2, 1 GiB) (config 3), each per rank (weak scaling at N > 1).
With N > 1 (torchrun) every rank encodes its packet range with gpuar_b200_encode_sharded: the ranks'
totals travel through peer-mapped mailboxes and the compaction kernel writes every packet straight to
its place in the concatenated stream, laid out in N equal segments, segment g on GPU g (peer stores
over NVLink; no NCCL call and no host round trip inside the timed region).  Decode at N > 1 is
gpuar_b200_decode_sharded: every rank discovers the packet chain of its own segment and decodes the
packets that start there.  The gather of the whole stream onto one GPU and the no-exchange case are
timed beside it.

Nothing is printed unless parity holds ("parity" in the line): every rank's decoded slice equals the
regenerated input, the head of the concatenated stream has the md5 of the reference's own output for
it (tests/golden), and at N > 1 rank 0 decodes the whole gathered stream with the single-GPU path.

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
    # name: (generator, seed, bytes, description); bytes = the whole job for STRONG workloads, per rank otherwise
    "m16g": ("mixed", 3, 16 << 30, "mixed(3, 16 GiB), one input sharded by packet range over the GPUs (BASELINE config 4)"),
    "u64m": ("uniform", 0x64, 64 << 20, "uniform(0x64, 64 MiB) = data/random_64m.dat stand-in (BASELINE config 2)"),
    "s1g": ("and3", 2, 1 << 30, "and3(2, 1 GiB) low-entropy (BASELINE config 3)"),
    "m2g": ("mixed", 3, 2 << 30, "mixed(3, 2 GiB per rank) (BASELINE config 4 shard, weak)"),
    "m4g": ("mixed", 7, 4 << 30, "mixed(7, 4 GiB) (BASELINE config 5)"),
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


STRONG = {"m16g"}             # one fixed job split over the ranks; everything else is per rank
# md5 of the head of the reference's payload for each generator/seed (tests/golden/vectors.json: m1m, s1m, u64m)
PREFIX_GOLDEN = {("mixed", 3): "m1m", ("and3", 2): "s1m", ("uniform", 0x64): "u64m"}


def kernel_source_hash():
    """sha256 over the kernel sources with comments and blank space removed: ties profiles/traffic.json (ncu DRAM
    bytes) to the code it was captured from, not to its commentary."""
    import hashlib
    import re
    h = hashlib.sha256()
    d = os.path.join(ROOT, "gpuar_b200", "csrc")
    for name in ("coder_math.h", "encode_math.h", "decode_math.h", "common.cuh", "lookback.cuh", "shard.cuh", "encode.cu",
                 "encode_ws.cu", "decode.cu"):
        text = open(os.path.join(d, name), encoding="utf-8").read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)            # block comments
        text = re.sub(r"//[^\n]*", "", text)                         # line comments (no string of these files holds //)
        h.update(re.sub(r"\s+", " ", text).encode())
    return h.hexdigest()[:16]


def measured_traffic(workload, kernel):
    """(dram__bytes_read.sum + dram__bytes_write.sum per launch, note) from the committed ncu --set full capture of
    this workload (profiles/traffic.json).  The capture records the hash of the kernel sources it was taken from;
    a number from other sources is not reported."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        entry = t[workload]
    except Exception:
        return None, "no ncu --set full capture of this workload under profiles/"
    if entry.get("kernel_sources_sha") != kernel_source_hash():
        return None, "profiles/traffic.json was captured from other kernel sources (stale): not reported"
    return entry.get(kernel), f"ncu --set full capture {entry.get('capture', '?')} of these kernel sources"


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



def job_config(args, world):
    """The `config` object of the JSON line: the same for both arms (the reference arm times the reference's CPU
    codec on a bounded sample of the same workload)."""
    gen, seed, nbytes, desc = WORKLOADS[args.workload]
    strong = args.workload in STRONG
    return {"workload": args.workload, "desc": desc, "job_bytes": nbytes if strong else nbytes * world,
            "packet_bytes": args.packet, "l2": "flushed between timed iterations (256 MiB write)",
            "parallelism": f"packet-range shards x{world}" if world > 1 else "single GPU",
            "encode_path": args.encode_path, "work_unit_packets": args.work_unit or "auto",
            "sub_workloads": [w for w in args.sub.split(",") if w]}


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
        "scaling": "strong" if args.workload in STRONG else "weak", "vs_baseline": None, "dtype": "u16",
        "data": "synthetic", "config": job_config(args, args.gpus),
        "decode": {"value": sample / t_dec / GB, "unit": "GB/s", "ms_per_step": t_dec * 1e3},
        "cpu_baseline": {"value": v, "value_per_core": v / cores, "unit": "GB/s", "cores": cores, "kind": kind,
                         "sample": f"first {sample >> 20} MiB of the workload per step, packets partitioned over "
                                   f"{cores} host threads (the reference itself is single-threaded)"},
        "e2e": {"value": v, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ==================================================================== our arm
def cpu_baseline(workload):
    """Reference CPU codec on this box's host cores, bounded sample (rank 0, N=1 only).  Returns the JSON object,
    the sample and the reference's payload for it (the parity check compares our stream's head with it)."""
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
    out = {"value": sample / t_enc / GB, "value_per_core": sample / t_enc / GB / cores, "unit": "GB/s", "cores": cores,
           "kind": kind, "decode_value": sample / t_dec / GB, "decode_value_per_core": sample / t_dec / GB / cores,
           "sample": f"first {sample >> 20} MiB of the workload, one pass, packets partitioned over {cores} host "
                     f"threads ({(t_enc + t_dec) * cores:.1f} core-seconds)"}
    if kind == "reference":
        out["reference_gpu_kernel"] = reference_gpu_kernel(data, pay)
    return out, data, pay


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


def gen_device(gen, seed, start, n):
    """Bytes [start, start+n) of the synthetic stream on the current device (start a multiple of 8192)."""
    import torch
    from gpuar_b200 import datagen as D
    fn = {"uniform": D.uniform_device, "and3": D.and3_device, "mixed": D.mixed_device}[gen]
    out = torch.empty(n, dtype=torch.uint8, device="cuda")
    pos, end = start, start + n
    while pos < end:                                               # 512 MiB at a time: bounded temporaries
        a0 = pos // 65536 * 65536
        b = min(end, a0 + (1 << 29))
        b1 = (b + 65535) // 65536 * 65536
        piece = fn(seed, b1 - a0, a0)
        out[pos - start: b - start] = piece[pos - a0: b - a0]
        del piece
        pos = b
    return out


def equals_stream(gen, seed, start, got):
    """got == bytes [start, start + len(got)) of the synthetic stream, compared 512 MiB at a time."""
    import torch
    n = got.numel()
    for a in range(0, n, 1 << 29):
        b = min(n, a + (1 << 29))
        if not torch.equal(got[a:b], gen_device(gen, seed, start + a, b - a)):
            return False
    return True


class Rig:
    """What every measurement shares: the rank, its codec, the two sharded layouts, the L2 flush buffer."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        from gpuar_b200 import _lib, codec
        from gpuar_b200.shard import ShardedCodec
        self.args = args
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        assert self.world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={self.world}: launch with torch.distributed.run"
        torch.cuda.set_device(self.local)
        if self.world > 1:
            # NCCL announces its version on stdout at communicator creation: keep stdout for the JSON line.
            # NCCL is plumbing here (setup, barriers, the max over ranks): no call inside a timed region.
            sys.stdout.flush()
            saved = os.dup(1)
            os.dup2(2, 1)
            try:
                dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
                dist.barrier()
                torch.cuda.synchronize()
            finally:
                sys.stdout.flush()
                os.dup2(saved, 1)
                os.close(saved)
        # a second group on the host (gloo): a rank that waits in an NCCL barrier spins in a kernel on its GPU, and
        # that GPU is needed by rank 0's process during the single-process multi-GPU end-to-end leg
        self.host_pg = dist.new_group(backend="gloo") if self.world > 1 else None
        self.dev = codec.DeviceCodec(self.local)
        _lib.set_option(_lib.OPT_ENCODE_PATH, {"auto": 0, "fused": 1, "ws": 2}[args.encode_path])
        if args.work_unit:
            _lib.set_option(_lib.OPT_COMPACT_TILE, args.work_unit)
        self.sharded = ShardedCodec(self.dev, self.rank, self.world, "segments")
        self.gathered = ShardedCodec(self.dev, self.rank, self.world, "gather") if self.world > 1 else None
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")       # > 126 MB L2

    def barrier(self):
        import torch
        import torch.distributed as dist
        torch.cuda.synchronize()
        if self.world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def host_barrier(self):
        """All ranks meet on the HOST: the GPUs stay idle while they wait."""
        import torch
        import torch.distributed as dist
        torch.cuda.synchronize()
        if self.world > 1:
            dist.barrier(group=self.host_pg)

    def timed(self, step, k):
        """K steps, each bracketed by CUDA events on the launching stream, L2 flushed in between."""
        import torch
        evs = []
        for _ in range(k):
            self.flush.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            step()
            b.record()
            evs.append((a, b))
        torch.cuda.synchronize()
        return sum(a.elapsed_time(b) for a, b in evs) / 1e3

    def all_true(self, ok):
        import torch
        import torch.distributed as dist
        if self.world == 1:
            return bool(ok)
        f = torch.tensor([1 if ok else 0], device="cuda")
        dist.all_reduce(f, op=dist.ReduceOp.MIN)
        return bool(f.item())


def golden_prefix(gen, seed):
    """(bytes, md5) of the reference's payload for the head of this stream, from tests/golden/vectors.json."""
    name = PREFIX_GOLDEN.get((gen, seed))
    if not name:
        return None
    try:
        v = json.load(open(os.path.join(ROOT, "tests", "golden", "vectors.json")))
        v = v["vectors"] if "vectors" in v else v
        e = v[name] if isinstance(v, dict) else next(x for x in v if x.get("name") == name)
        return int(e["payload_bytes"]), e["payload_md5"], name
    except Exception:
        return None


def measure(rig, workload, steps, warmup, sampler=None, with_cpu_baseline=False):
    """One workload: parity first, then the timed encode / decode / end-to-end legs.  Returns the result object."""
    import hashlib
    import numpy as np
    import torch
    import torch.distributed as dist
    from gpuar_b200 import _lib, codec
    from gpuar_b200.shard import byte_range

    args, rank, world, dev = rig.args, rig.rank, rig.world, rig.dev
    gen, seed, nbytes, desc = WORKLOADS[workload]
    strong = workload in STRONG
    job = nbytes if strong else nbytes * world
    b0, b1 = byte_range(job, rank, world) if strong else (rank * nbytes, (rank + 1) * nbytes)
    n = b1 - b0
    packet = args.packet if world == 1 else 8192                    # the sharded entry points speak the 8192-byte dialect
    x = gen_device(gen, seed, b0, n)
    packets = (n + packet - 1) // packet
    job_packets = (job + packet - 1) // packet
    result = torch.zeros(8, dtype=torch.int64, device="cuda")
    if world == 1:
        cap = codec.payload_bound(n, packet)
        payload = torch.empty(cap + 16, dtype=torch.uint8, device="cuda")
        total = torch.zeros(1, dtype=torch.int64, device="cuda")
        offsets = torch.empty(packets + 1, dtype=torch.int64, device="cuda")
        out = torch.empty(packets * packet, dtype=torch.uint8, device="cuda")
        c_holder = [0]

        def enc_step():
            dev.encode(x, payload, total, packet=packet)

        def dec_step():
            dev.index(payload, c_holder[0], packets, offsets, result, packet=packet)
            dev.decode(payload, c_holder[0], offsets, packets, out, packet=packet)
    else:
        rig.sharded.reserve(codec.payload_bound(max(n, 8192)))
        rig.gathered.reserve(codec.payload_bound(max(n, 8192)))
        # a rank decodes the packets that START in its segment: about job_packets / world of them
        out = torch.empty((job_packets * 5 // (4 * world) + 64) * 8192, dtype=torch.uint8, device="cuda")
        c_holder = [0]

        def enc_step():
            rig.sharded.encode(x)          # totals through the mailboxes, packets straight to their segment

        def dec_step():
            rig.sharded.decode(c_holder[0], out, result)

    def stream_total():
        torch.cuda.synchronize()
        if world == 1:
            return int(total.item()), int(total.item())
        lay = [int(v) for v in rig.sharded.group.layout_out.tolist()[:5]]
        assert lay[4] == 0, f"gpuar_b200_encode_sharded status {lay[4]}"
        return lay[0], lay[3]

    # ---- warm-up (>= 3) and the parity gate: nothing is reported unless it holds
    w = max(3, warmup)
    for _ in range(w):
        enc_step()
    c_all, c_mine = stream_total()
    c_holder[0] = c_all
    rig.barrier()                                   # the concatenated stream is complete on every GPU
    for _ in range(w):
        dec_step()
    if sampler is not None:
        # a 64 MiB step lasts ~3 ms: keep the same kernels running (untimed) for ~0.2 s so that the sampler
        # sees the load; the count depends on the size only, so every rank does the same (the steps are collective)
        for _ in range(max(3, min(100, int(0.2 / (n / 30e9))))):
            enc_step()
            dec_step()
    torch.cuda.synchronize()
    parity = {}
    if world == 1:
        res = [int(v) for v in result.tolist()[:3]]
        parity["round_trip"] = res == [packets, n, 0] and bool(torch.equal(out[:n], x))
        head = payload
    else:
        mine, raw, status, before, raw_before = (int(v) for v in result.tolist()[:5])
        ok = status == 0 and raw_before == min(before * 8192, job) and equals_stream(gen, seed, raw_before, out[:raw])
        counts = torch.tensor([mine, raw], dtype=torch.int64, device="cuda")
        dist.all_reduce(counts)
        covered = [int(v) for v in counts.tolist()] == [job_packets, job]
        # every rank's decoded slice of the CONCATENATED stream equals the input, and the slices cover the job
        parity["concat_round_trip"] = rig.all_true(ok and covered)
        # the whole stream gathered onto rank 0 and decoded there by the single-GPU path
        for _ in range(2):
            rig.gathered.encode(x)
        rig.barrier()
        good = True
        head = None
        if rank == 0:
            view, valid = rig.gathered.group.my_segment()
            head = view
            good = valid == c_all
            if good:
                padded = torch.as_tensor(rig.gathered.group.segment_view(valid + 64))
                padded[valid:] = 0
                offs, res = dev.index(padded, valid, job_packets)
                full = dev.decode(padded, valid, offs, job_packets)
                torch.cuda.synchronize()
                good = [int(v) for v in res.tolist()[:3]] == [job_packets, job, 0] and equals_stream(gen, seed, 0, full[:job])
                del full, offs
        parity["gathered_round_trip_on_rank0"] = rig.all_true(good)
    gp = golden_prefix(gen, seed) if packet == 8192 else None
    if gp and rank == 0 and c_all >= gp[0]:
        got = hashlib.md5(head[: gp[0]].cpu().numpy().tobytes()).hexdigest()
        parity["prefix_md5"] = got == gp[1]
        parity["prefix_md5_of"] = f"first {gp[0]} stream bytes == the reference's payload for the first " \
                                  f"{'64 MiB' if gp[2] == 'u64m' else '1 MiB'} of this input (tests/golden {gp[2]})"
    flags = [v for k, v in parity.items() if isinstance(v, bool)]
    assert rig.all_true(all(flags)), f"parity failed on rank {rank}: {parity}"
    head = None

    if sampler is not None:
        sampler.mark()
    launches0 = _lib.launch_count()
    _lib.profile(True)
    _lib.profile_read()
    rig.barrier()
    t_enc = rig.timed(enc_step, steps)
    rig.barrier()
    t_local = t_gather = t_enc
    if world > 1:
        payload = torch.empty(codec.payload_bound(n) + 16, dtype=torch.uint8, device="cuda")
        total = torch.zeros(1, dtype=torch.int64, device="cuda")
        local_step = lambda: dev.encode(x, payload, total)
        for _ in range(2):
            local_step()
        rig.barrier()
        t_local = rig.timed(local_step, steps)
        del payload
        rig.barrier()
        gather_step = lambda: rig.gathered.encode(x)
        for _ in range(2):
            gather_step()
        rig.barrier()
        t_gather = rig.timed(gather_step, steps)
        rig.barrier()
        for _ in range(2):                         # the stream the decode reads: back to the segments layout
            enc_step()
        rig.barrier()
    t_dec = rig.timed(dec_step, steps)
    rig.barrier()
    spans = _lib.profile_read()
    _lib.profile(False)
    launches = _lib.launch_count() - launches0

    # ---- e2e: ONE input in pinned host memory -> ONE .gip image in pinned host memory through the library's
    # host entry point (H2D + kernels + D2H inside the timed region).  N > 1: rank 0's process drives all N GPUs
    # with gpuar_b200_compress_host_multi (chunks rotate over the devices); the other ranks wait.
    if world > 1:
        del out
        out = None
    torch.cuda.empty_cache()
    t_e2e_enc = t_e2e_dec = t_link = t_link_dec = 0.0
    gip_bytes = 0
    host_n = n if world == 1 else job
    devices = None if world == 1 else list(range(world))
    rig.host_barrier()
    if rank == 0:
        host_in = torch.empty(host_n, dtype=torch.uint8, pin_memory=True)
        if world == 1:
            host_in.copy_(x)
        else:
            for a in range(0, job, 1 << 29):
                b = min(job, a + (1 << 29))
                host_in[a:b].copy_(x[a - b0: b - b0] if (a >= b0 and b <= b1) else gen_device(gen, seed, a, b - a))
        host_gip = torch.empty(20 + codec.payload_bound(host_n), dtype=torch.uint8, pin_memory=True)   # e2e: 8192-byte packets
        host_out = torch.empty(host_n + 8192, dtype=torch.uint8, pin_memory=True)
        np_in, np_gip, np_out = host_in.numpy(), host_gip.numpy(), host_out.numpy()
        e2e_steps = max(1, min(steps, 3 if host_n >= (1 << 30) else 5))
        g = codec.compress(np_in, out=np_gip, devices=devices)
        back = codec.decompress(g, out=np_out, devices=devices)
        assert back.size == host_n and np.array_equal(back, np_in), "host path does not round-trip"
        gip_bytes = int(g.size)
        if world > 1:
            # the image assembled from N devices' chunks is the reference's stream: head against the golden md5
            gp2 = golden_prefix(gen, seed)
            if gp2 and gip_bytes - 20 >= gp2[0]:
                parity["e2e_image_prefix_md5"] = hashlib.md5(g[20: 20 + gp2[0]].tobytes()).hexdigest() == gp2[1]
                assert parity["e2e_image_prefix_md5"]
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            g = codec.compress(np_in, out=np_gip, devices=devices)
        t_e2e_enc = (time.perf_counter() - t0) / e2e_steps
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            back = codec.decompress(g, out=np_out, devices=devices)
        t_e2e_dec = (time.perf_counter() - t0) / e2e_steps
        # the same transfers without the kernels: the ceiling of the links + host memory for this pattern
        codec.link_probe(np_in, np_out, gip_bytes - 20, devices=devices)
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            codec.link_probe(np_in, np_out, gip_bytes - 20, devices=devices)
        t_link = (time.perf_counter() - t0) / e2e_steps
        # ... and of the decode direction: the image up, the raw bytes down
        codec.link_probe(g, np_out, host_n, devices=devices)
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            codec.link_probe(g, np_out, host_n, devices=devices)
        t_link_dec = (time.perf_counter() - t0) / e2e_steps
        del host_in, host_gip, host_out, np_in, np_gip, np_out, g, back
    rig.host_barrier()

    # ---- max over ranks (per-step seconds)
    times = torch.tensor([t_enc / steps, t_dec / steps, t_local / steps, t_gather / steps], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    t_enc, t_dec, t_local, t_gather = (float(v) for v in times.tolist())

    res = None
    if rank == 0:
        peak, peak_src = measured_peaks()
        per = lambda k: spans[k][0] / max(1, spans[k][1])                       # ms per launch group
        alg = n + c_mine                                                        # N read + C written per launch
        enc_traffic, enc_note = measured_traffic(workload, "encode")
        dec_traffic, _ = measured_traffic(workload, "decode")
        res = {
            "workload": workload, "desc": desc, "scaling": "strong" if strong else "weak", "job_bytes": job,
            "bytes_rank0": n, "payload_bytes": c_all, "payload_bytes_rank0": c_mine,
            "value": job / t_enc / GB, "unit": "GB/s", "ms_per_step": t_enc * 1e3,
            "decode": {"value": job / t_dec / GB, "unit": "GB/s", "ms_per_step": t_dec * 1e3,
                       "includes": "device packet-chain discovery + decode kernel" +
                                   (" (every rank: its own segment of the concatenated stream)" if world > 1 else "")},
            "e2e": {"value": job / t_e2e_enc / GB, "unit": "GB/s", "h2d_bytes_per_step": host_n,
                    "d2h_bytes_per_step": gip_bytes - 20,
                    "api": "gpuar_b200_compress_host (pinned host buffers)" if world == 1 else
                           f"gpuar_b200_compress_host_multi: one process, one host thread, {world} GPUs, one input -> one .gip image",
                    "link_ceiling": {"value": job / t_link / GB, "unit": "GB/s",
                                     "what": "gpuar_b200_host_link_probe: the same chunks up and the same bytes down on the "
                                             "same streams, no kernels (PCIe links + host memory for this pattern)"}},
            "e2e_decode": {"value": job / t_e2e_dec / GB, "unit": "GB/s", "h2d_bytes_per_step": gip_bytes - 20,
                           "d2h_bytes_per_step": host_n,
                           "api": "gpuar_b200_decompress_host" if world == 1 else "gpuar_b200_decompress_host_multi",
                           "link_ceiling": {"value": job / t_link_dec / GB, "unit": "GB/s",
                                            "what": "gpuar_b200_host_link_probe with the image going up and the raw bytes coming down"}},
            "parity": parity,
            "gpu_launches": launches,
            "kernels_ms_per_step": {k: per(k) for k in spans},
            "roofline": {"bound": "hbm", "kernel": "encode kernel (encode_ws_kernel up to two waves of its CTAs, else encode_kernel)",
                         "achieved": alg / (per("encode") / 1e3) / GB, "peak": peak, "unit": "GB/s",
                         "frac": alg / (per("encode") / 1e3) / GB / peak, "traffic": enc_traffic, "traffic_source": enc_note,
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": alg,
                         "note": "integer-latency-bound path: see profiles/ for issue-slot and occupancy evidence"},
            "roofline_decode": {"bound": "hbm", "kernel": "decode_kernel",
                                "achieved": alg / (per("decode") / 1e3) / GB, "peak": peak, "unit": "GB/s",
                                "frac": alg / (per("decode") / 1e3) / GB / peak, "traffic": dec_traffic,
                                "algorithmic_bytes_per_launch": alg},
        }
        if world > 1:
            res["encode_shards_in_place"] = {"value": job / t_local / GB, "unit": "GB/s",
                                             "note": "encode only, every shard's payload left on its own GPU (SURVEY 8e)"}
            res["encode_gather_one_gpu"] = {"value": job / t_gather / GB, "unit": "GB/s",
                                            "note": "the whole stream written into rank 0's memory: ingress-limited on that "
                                                    "GPU (SURVEY 8e); `value` writes W equal segments, segment g on GPU g"}
            res["exchange"] = "per-rank totals through peer-mapped mailboxes, packets stored straight into the " \
                              "segments over NVLink; no NCCL call and no host round trip in the timed region"
        if with_cpu_baseline:
            base, sample, ref_pay = cpu_baseline(workload)
            res["cpu_baseline"] = base
            if packet == 8192 and base["kind"] == "reference":
                # the head of our stream against the reference's own output for the same bytes
                k = min(int(ref_pay.size), c_all)
                src = payload if world == 1 else None
                if src is not None:
                    res["parity"]["head_equals_reference_cpu_output"] = bool(
                        np.array_equal(src[:k].cpu().numpy(), ref_pay[:k]))
                    res["parity"]["head_bytes_compared"] = k
                    assert res["parity"]["head_equals_reference_cpu_output"], "payload differs from the reference's"
    return res


def run_ours(args):
    import torch
    import torch.distributed as dist

    rig = Rig(args)
    rank, world = rig.rank, rig.world
    sampler = ClockSampler(rig.local)
    if rank == 0:
        sampler.start()
        sampler.wait_first()                      # nvidia-smi is looping from here on
    main = measure(rig, args.workload, args.steps, args.warmup, sampler=sampler,
                   with_cpu_baseline=(world == 1 and not args.no_cpu_baseline))
    clocks = sampler.stop() if rank == 0 else None
    subs = {}
    for name in [w for w in args.sub.split(",") if w and w != args.workload]:
        torch.cuda.empty_cache()
        subs[name] = measure(rig, name, args.steps, args.warmup)
    if rank == 0:
        line = {
            "metric": "encode_GBps", "value": main["value"], "unit": "GB/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": main["ms_per_step"], "higher_is_better": True,
            "scaling": main["scaling"], "vs_baseline": None, "dtype": "u16", "data": "synthetic",
            "config": job_config(args, world),
        }
        for k in ("decode", "e2e", "e2e_decode", "parity", "roofline", "roofline_decode", "kernels_ms_per_step",
                  "encode_shards_in_place", "encode_gather_one_gpu", "exchange", "cpu_baseline", "payload_bytes",
                  "bytes_rank0", "payload_bytes_rank0"):
            if k in main:
                line[k] = main[k]
        line["gpu_launches"] = main["gpu_launches"] + sum(v["gpu_launches"] for v in subs.values())
        line["clocks"] = clocks
        line["sub"] = subs
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
    ap.add_argument("--workload", default="m16g", choices=sorted(WORKLOADS))
    ap.add_argument("--sub", default="u64m,s1g",
                    help="comma-separated workloads measured after the main one and reported under \"sub\" ('' = none)")
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
