#!/usr/bin/env python
"""bench.py -- the hot path of cute-nucleotides on B200: 2-bit encode (n_to_bits) + decode (bits_to_n).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One STEP = one encode pass (ASCII -> packed u64) followed by one decode pass (packed -> ASCII) over one
batch of L synthetic nucleotides per GPU (BASELINE config 3+4 shape: 10 GiB, mixed case with U), buffers
resident in HBM.  `value` = nucleotides carried through the round trip per second, summed over GPUs
(weak scaling: every rank owns its own L-nucleotide shard of one logical sequence; the path needs no
data-path collective).  `e2e` = the same round trip through the host-slice C-ABI entry points
(cn_n_to_bits_host / cn_bits_to_n_host) with pinned HOST buffers, H2D and D2H inside the timed region.

`--impl reference` times the reference's own fastest CPU variants (n_to_bits_movemask + bits_to_n_shuffle,
restated in C because the Rust crate cannot be built here: no rustc/cargo) on all host cores.
The oracle is loaded ONLY for the cpu_baseline / reference legs; the post-run verification of the timed outputs
is an independent torch computation.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

GIB = 1 << 30
METRIC = "nucleotides/sec (encode+decode round trip)"
UNIT = "nucleotides/s"
BYTES_PER_NT = 1.25            # per direction: encode 1 B read + 0.25 B written; decode 0.25 B + 1 B
SEED = 0xC0FFEE


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, torch copy_ read+write)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s; MEASURED_PEAKS.json absent)"


def ncu_traffic():
    """Per-launch DRAM bytes of the dominant kernel from the committed ncu --set full capture, if any."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f)
    except Exception:
        return None


class ClockSampler(threading.Thread):
    """Samples SM clock / throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index, period=0.02):
        super().__init__(daemon=True)
        self.index, self.period = index, period
        self.samples, self.reasons, self.power = [], set(), []
        self.max_mhz = None
        self._halt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(visible.split(",")[index]) if visible and visible.split(",")[index].isdigit() else index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8)),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40)),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20)),
            "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)),
            "hw_power_brake": getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80)),
        }
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or getattr(nv, "nvmlDeviceGetCurrentClocksThrottleReasons")
        while not self._halt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0)
                mask = get_reasons(self.h)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._halt.wait(self.period)

    def stop(self):
        self._halt.set()
        if self.ok:
            self.join(timeout=2)

    def summary(self):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0, "note": "NVML unavailable"}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples), "power_w_max": round(max(self.power), 1) if self.power else None}


# -----------------------------------------------------------------------------------------------------
# CPU arm: the reference's fastest variants (restated), all host threads, bounded sample
# -----------------------------------------------------------------------------------------------------
def cpu_roundtrip(sample_nt, min_seconds, min_reps, alphabet):
    """Times oracle n_to_bits_movemask + bits_to_n_shuffle (multi-threaded by sequence offset) on a
    `sample_nt` sample of the workload.  Returns dict with nt/s (round trip), per-direction rates."""
    import numpy as np
    import _oracle
    orc = _oracle.Oracle()
    threads = os.cpu_count() or 1
    n = orc.generate(sample_nt, SEED, alphabet)
    words = np.zeros(orc.words_for_len(sample_nt), dtype=np.uint64)      # pre-faulted outputs
    out = np.zeros(sample_nt, dtype=np.uint8)
    enc_v, dec_v = ("movemask", "shuffle") if orc.simd_ok else ("lut", "lut")
    orc.encode_mt(n, enc_v, threads, out=words)                          # warm-up
    orc.decode_mt(words, sample_nt, dec_v, threads, out=out)
    assert out[: 1 << 20].tobytes() == orc.canonical(n[: 1 << 20])
    t_enc, t_dec = [], []
    t_begin = time.perf_counter()
    while len(t_enc) < min_reps or time.perf_counter() - t_begin < min_seconds:
        t0 = time.perf_counter()
        orc.encode_mt(n, enc_v, threads, out=words)
        t1 = time.perf_counter()
        orc.decode_mt(words, sample_nt, dec_v, threads, out=out)
        t2 = time.perf_counter()
        t_enc.append(t1 - t0)
        t_dec.append(t2 - t1)
    step = [a + b for a, b in zip(t_enc, t_dec)]
    # the reference as shipped is single-threaded: the same two variants on ONE thread, for context
    one_enc, one_dec = [], []
    for _ in range(2):
        t0 = time.perf_counter()
        orc.encode_mt(n, enc_v, 1, out=words)
        t1 = time.perf_counter()
        orc.decode_mt(words, sample_nt, dec_v, 1, out=out)
        t2 = time.perf_counter()
        one_enc.append(t1 - t0)
        one_dec.append(t2 - t1)
    return {
        "single_thread": {"encode_nt_per_s": sample_nt / min(one_enc), "decode_nt_per_s": sample_nt / min(one_dec),
                          "value": sample_nt / (min(one_enc) + min(one_dec))},
        "value": sample_nt / statistics.mean(step), "best": sample_nt / min(step), "reps": len(step),
        "encode_nt_per_s": sample_nt / statistics.mean(t_enc), "decode_nt_per_s": sample_nt / statistics.mean(t_dec),
        "cores": threads, "variants": f"n_to_bits_{enc_v} + bits_to_n_{dec_v}", "step_s": step,
    }


def cpu_baseline_leg(args, L):
    """The `cpu_baseline` object of the GPU arm's line (guarded: never fatal)."""
    sample = min(L, args.cpu_sample)
    try:
        r = cpu_roundtrip(sample, args.cpu_seconds, 5, args.alphabet)
    except Exception as e:                  # noqa: BLE001
        return {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": "failed", "error": repr(e)}
    return {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port",
            "sample": f"{sample} nt ({sample / GIB:.2f} GiB) of the workload x {r['reps']} reps, {r['variants']} "
                      f"(AVX2 restatement of the reference's fastest variants; Rust crate not buildable here), "
                      f"sharded by offset over {r['cores']} host threads, outputs pre-faulted",
            "encode_nt_per_s": r["encode_nt_per_s"], "decode_nt_per_s": r["decode_nt_per_s"], "best": r["best"],
            "single_thread": r["single_thread"]}


def run_reference(args, rank):
    if rank != 0:
        return
    sample = min(args.nucleotides, args.cpu_sample)
    steps, warmup = args.steps, args.warmup
    import numpy as np
    import _oracle
    orc = _oracle.Oracle()
    threads = os.cpu_count() or 1
    n = orc.generate(sample, SEED, args.alphabet)
    words = np.zeros(orc.words_for_len(sample), dtype=np.uint64)
    out = np.zeros(sample, dtype=np.uint8)
    enc_v, dec_v = ("movemask", "shuffle") if orc.simd_ok else ("lut", "lut")
    # bound the whole run to a few minutes: cap steps by a time budget measured on the first step
    t_enc = t_dec = 0.0
    done = 0
    budget_s = 120.0
    t_all = time.perf_counter()
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        orc.encode_mt(n, enc_v, threads, out=words)
        t1 = time.perf_counter()
        orc.decode_mt(words, sample, dec_v, threads, out=out)
        t2 = time.perf_counter()
        if i >= warmup:
            t_enc += t1 - t0
            t_dec += t2 - t1
            done += 1
        if time.perf_counter() - t_all > budget_s and done >= 3:
            break
    ms_step = (t_enc + t_dec) / done * 1e3
    value = sample / (ms_step / 1e3)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": done,
        "warmup": warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": {"workload": workload_name(args), "nucleotides_per_step": sample, "alphabet": args.alphabet,
                   "note": "CPU arm: each step is a bounded sample of the workload; rank 0 only"},
        "encode_nt_per_s": sample * done / t_enc, "decode_nt_per_s": sample * done / t_dec,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{sample} nt ({sample / GIB:.2f} GiB) of the workload per step, {enc_v}+{dec_v} AVX2 restatement "
                                   "of the reference (Rust crate not buildable here: no rustc/cargo), sharded by offset over all host threads"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_name(args):
    alpha = "{A,C,G,T,U,a,c,g,t,u}" if args.alphabet == 10 else "{A,C,G,T}"
    return (f"{args.nucleotides / GIB:g} GiB random {alpha} per GPU: n_to_bits encode then bits_to_n decode "
            f"(BASELINE configs 3+4 shape)")


# -----------------------------------------------------------------------------------------------------
# GPU arm
# -----------------------------------------------------------------------------------------------------
def run_gpu(args, rank, local_rank, world):
    import numpy as np
    import torch
    import torch.distributed as dist
    import cute_nucleotides_b200 as cn
    from cute_nucleotides_b200 import _lib

    lib = _lib.load()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local_rank])
        torch.cuda.synchronize()

    from cute_nucleotides_b200 import sharded
    strong = args.total_gib > 0
    if strong:
        total = int(args.total_gib * GIB)
        offset, end = sharded.shard_bounds(total, world, rank, 1 << 20)
        L = end - offset
    else:
        L = args.nucleotides
        offset = rank * L                                # shard by sequence offset of one logical N*L sequence
    W = cn.words_for_len(L)
    d_n = torch.empty(L, dtype=torch.uint8, device=dev)
    d_bits = torch.empty(W, dtype=torch.int64, device=dev)
    free_b, _ = torch.cuda.mem_get_info(dev)
    in_place = free_b < L + (8 << 30)                    # SURVEY 7 hard part 4: 80 GiB + 20 GiB + 80 GiB > one B200
    d_out = d_n if in_place else torch.empty(L, dtype=torch.uint8, device=dev)
    cn.generate_device(d_n, offset, SEED, args.alphabet)
    if in_place:                                         # the round trip then maps the buffer onto its canonical form:
        cn.encode_device(d_n, out=d_bits)                # start from the canonical form so every step does the same work
        cn.decode_device(d_bits, L, out=d_n)
    stream = torch.cuda.current_stream()

    def step():
        cn.encode_device(d_n, out=d_bits, stream=stream)
        cn.decode_device(d_bits, L, out=d_out, stream=stream)

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()

    K = args.steps
    marks = [torch.cuda.Event(enable_timing=True) for _ in range(2 * K + 1)]
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    launches0 = lib.cn_launch_count()
    barrier()
    marks[0].record(stream)
    for k in range(K):
        cn.encode_device(d_n, out=d_bits, stream=stream)
        marks[2 * k + 1].record(stream)
        cn.decode_device(d_bits, L, out=d_out, stream=stream)
        marks[2 * k + 2].record(stream)
    barrier()
    launches = lib.cn_launch_count() - launches0
    if sampler:
        sampler.stop()
    total_ms = marks[0].elapsed_time(marks[2 * K])
    enc_ms = [marks[2 * k].elapsed_time(marks[2 * k + 1]) for k in range(K)]
    dec_ms = [marks[2 * k + 1].elapsed_time(marks[2 * k + 2]) for k in range(K)]
    t = torch.tensor([total_ms, sum(enc_ms) / K, sum(dec_ms) / K], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, enc_avg, dec_avg = (float(x) for x in t.tolist())
    ms_per_step = total_ms / K

    # ---- verification (outside the timed region): the timed outputs are bit-exact ---------------------
    verified = verify(cn, torch, np, d_n, d_bits, d_out, L, offset, args.alphabet, in_place)

    # ---- optional assemble of the packed shards (the only collective the path can use; not in `value`)
    assemble = None
    if world > 1 and not args.no_assemble:
        full = torch.empty(W * world, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(full, d_bits)
        barrier()
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(2):
            dist.all_gather_into_tensor(full, d_bits)
        a1.record()
        barrier()
        ta = torch.tensor([a0.elapsed_time(a1) / 2], dtype=torch.float64, device=dev)
        dist.all_reduce(ta, op=dist.ReduceOp.MAX)
        assemble = {"op": "nccl all_gather_into_tensor of packed words (every rank ends with all shards)",
                    "ms": float(ta.item()), "bytes_per_rank": W * 8,
                    "busbw_gbs": W * 8 * (world - 1) / (float(ta.item()) * 1e-3) / 1e9}
        del full
        # the fused alternative: ONE kernel per rank reads the shard once and stores the packed words into every
        # rank's buffer over NVLink peer memory (no intermediate shard, no collective)
        try:
            asm = sharded.PeerAssembly(world * L if not strong else int(args.total_gib * GIB), granule=1 << 20)
            asm.encode(d_n); asm.finish()
            barrier()
            f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            f0.record()
            for _ in range(2):
                asm.encode(d_n)
            f1.record()
            barrier()
            tf = torch.tensor([f0.elapsed_time(f1) / 2], dtype=torch.float64, device=dev)
            dist.all_reduce(tf, op=dist.ReduceOp.MAX)
            ws, we = sharded.word_bounds(world * L if not strong else int(args.total_gib * GIB), world, rank, 1 << 20)
            same = bool(torch.equal(asm.full[ws:we], d_bits))
            assemble["fused_encode_assemble_ms"] = float(tf.item())
            assemble["unfused_encode_plus_allgather_ms"] = enc_avg + assemble["ms"]
            assemble["fused_matches"] = same
            assemble["fused_op"] = "cn_encode_multi_device: encode kernel stores packed words into all ranks' buffers over NVLink (CUDA IPC peer mappings)"
            asm.close()
            del asm
        except Exception as e:              # noqa: BLE001  (reported, never fatal for the bench line)
            assemble["fused_error"] = repr(e)

    # ---- e2e: host-slice C-ABI calls with pinned host buffers, copies inside the timed region ----------
    # auxiliary legs are guarded: a failure there (e.g. pinned-memory exhaustion) is reported inside the line, it must
    # never cost the device-resident measurement above
    e2e = None
    if not args.no_e2e:
        try:
            e2e = run_e2e(args, cn, lib, _lib, torch, np, rank, local_rank, world, barrier, dev)
        except Exception as e:              # noqa: BLE001
            e2e = {"value": None, "unit": UNIT, "h2d_bytes_per_step": None, "d2h_bytes_per_step": None, "error": repr(e)}

    # ---- extras (rank 0, N == 1): the widened rows measured the same way (CUDA events, device-resident) --------------
    extras = None
    if rank == 0 and world == 1 and not args.no_extras:
        try:
            extras = run_extras(cn, torch, d_n, d_bits, d_out, L)
        except Exception as e:              # noqa: BLE001
            extras = {"error": repr(e)}

    # ---- CPU baseline beside it (rank 0, N == 1 only) --------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline_leg(args, L)

    if rank != 0:
        return
    total_nt = int(args.total_gib * GIB) if strong else world * L
    peak, peak_src = measured_peak()
    dom = "decode" if dec_avg >= enc_avg else "encode"
    dom_ms = max(dec_avg, enc_avg)
    achieved = BYTES_PER_NT * L / (dom_ms * 1e-3) / 1e9
    traffic = ncu_traffic()
    import ctypes
    vec, unroll, threads = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    tuning = {}
    for name, d in (("encode", 0), ("decode", 1)):
        lib.cn_get_tuning(d, ctypes.byref(vec), ctypes.byref(unroll), ctypes.byref(threads))
        tuning[name] = {"vec_bytes": vec.value, "unroll": unroll.value, "threads": threads.value}
    line = {
        "metric": METRIC, "value": total_nt / (ms_per_step * 1e-3), "unit": UNIT, "n_gpus": world, "steps": K,
        "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if strong else "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": workload_name(args) if not strong else f"{args.total_gib:g} GiB random sequence sharded by offset over {world} GPU(s): encode then decode (BASELINE config 5)",
                   "nucleotides_per_gpu": L, "decode_in_place": bool(in_place), "alphabet": args.alphabet, "seed": SEED,
                   "sharding": "contiguous by sequence offset, one shard per rank, no data-path collective",
                   "l2": "inputs (>= 2.5 GiB per kernel) exceed the 126 MB L2; no flush needed", "tuning": tuning},
        "encode_nt_per_s": total_nt / (enc_avg * 1e-3), "decode_nt_per_s": total_nt / (dec_avg * 1e-3),
        "encode_ms": enc_avg, "decode_ms": dec_avg,
        "roofline": {"bound": "hbm", "kernel": f"{dom}_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak,
                     "traffic": (traffic or {}).get(dom) if L == 10 * GIB else None,   # the ncu capture is of 10 GiB launches
                     "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": BYTES_PER_NT * L,
                     "encode": {"achieved": BYTES_PER_NT * L / (enc_avg * 1e-3) / 1e9, "frac": BYTES_PER_NT * L / (enc_avg * 1e-3) / 1e9 / peak},
                     "decode": {"achieved": BYTES_PER_NT * L / (dec_avg * 1e-3) / 1e9, "frac": BYTES_PER_NT * L / (dec_avg * 1e-3) / 1e9 / peak}},
        "cpu_baseline": cpu,
        "e2e": e2e,
        "gpu_launches": int(launches),
        "clocks": sampler.summary() if sampler else None,
        "verified": verified,
    }
    if assemble:
        line["assemble"] = assemble
    if extras:
        line["extras"] = extras
    print(json.dumps(line), flush=True)


def run_extras(cn, torch, d_n, d_bits, d_out, L, iters=20):
    """Not part of `value`: the rows built after the 2-bit path on the same buffers.
    checked encode = n_to_bits + count of invalid bytes in one pass; base-5 = n_to_bits2 / bits_to_n2 (27 nt per u64,
    35/27 algorithmic bytes per nucleotide)."""
    def timed(fn):
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / iters

    counter = torch.zeros(1, dtype=torch.int64, device=d_n.device)
    ms_checked = timed(lambda: cn.encode_checked_device(d_n, counter, out=d_bits))
    cn.generate2_device(d_n, 0, SEED, 12)                        # 5-letter data for the base-5 codec, same buffer
    W2 = cn.words2_for_len(L)
    d_bits2 = torch.empty(W2, dtype=torch.int64, device=d_n.device)
    ms_e2 = timed(lambda: cn.encode2_device(d_n, out=d_bits2))
    ms_d2 = timed(lambda: cn.decode2_device(d_bits2, L, out=d_out))
    b2 = L + 8 * W2
    return {
        "encode_checked": {"ms": ms_checked, "gbs": BYTES_PER_NT * L / (ms_checked * 1e-3) / 1e9, "invalid_bytes_found": int(counter.item())},
        "base5_encode": {"ms": ms_e2, "nt_per_s": L / (ms_e2 * 1e-3), "gbs": b2 / (ms_e2 * 1e-3) / 1e9},
        "base5_decode": {"ms": ms_d2, "nt_per_s": L / (ms_d2 * 1e-3), "gbs": b2 / (ms_d2 * 1e-3) / 1e9},
        "note": "device-resident, CUDA events, same 10 GiB buffers; not part of `value`",
    }


def verify(cn, torch, np, d_n, d_bits, d_out, L, offset, alphabet, in_place=False):
    """Post-run check of the timed outputs WITHOUT the oracle (the oracle is confined to the cpu_baseline / reference
    legs and to tests/): (i) decode(encode(x)) == canonical(x) over a 256 MiB prefix and the ragged end, (ii) 4 MiB
    windows of packed words equal an independent torch computation of the contract: code = (byte >> 1) & 3 for the
    valid alphabet, nucleotide i at bits 2(i & 31) of word i >> 5 (src/n_to_bits.rs:39-42)."""
    dev = d_n.device
    lut = torch.zeros(256, dtype=torch.uint8, device=dev)
    for ch, canon in zip(b"ACGTUacgtu", b"ACGTTACGTT"):
        lut[ch] = canon
    ok = True
    span = min(L, 1 << 28)

    def source(s, n):                       # the input window, regenerated when the round trip overwrote it in place
        if in_place:
            return cn.generate_device(torch.empty(n, dtype=torch.uint8, device=dev), offset + s, SEED, alphabet)
        return d_n[s:s + n]

    for s in (0, L - span):
        ok &= bool(torch.equal(d_out[s:s + span], lut[source(s, span).long()]))
    win = min(L, 1 << 22) & ~31
    shifts = (torch.arange(32, device=dev, dtype=torch.int64) * 2)
    for s in (0, ((L - win) // 2) & ~31, (L - win) & ~31):
        codes = ((source(s, win).long() >> 1) & 3).view(-1, 32)
        expect = (codes << shifts).sum(dim=1)                       # disjoint bit fields: sum == or; wraps into int64
        ok &= bool(torch.equal(d_bits[s // 32: s // 32 + win // 32], expect))
    return ok


def run_e2e(args, cn, lib, _lib, torch, np, rank, local_rank, world, barrier, dev):
    import psutil
    L = args.e2e_nucleotides or args.nucleotides
    need = 2.3 * L                                        # ASCII in + packed + ASCII out, pinned
    avail = psutil.virtual_memory().available / max(world, 1)
    note = None
    while need > 0.6 * avail and L > (1 << 28):
        L //= 2
        need = 2.3 * L
        note = "e2e batch reduced to fit pinned host memory"
    W = cn.words_for_len(L)
    h_n = torch.empty(L, dtype=torch.uint8).pin_memory()
    h_bits = torch.empty(W, dtype=torch.int64).pin_memory()
    h_out = torch.empty(L, dtype=torch.uint8).pin_memory()
    # fill the host input with the shared synthetic sequence (generated on device, copied once, untimed)
    tmp = torch.empty(min(L, 1 << 30), dtype=torch.uint8, device=dev)
    for s in range(0, L, tmp.numel()):
        e = min(L, s + tmp.numel())
        cn.generate_device(tmp[: e - s], rank * L + s, SEED + 1, args.alphabet)
        h_n[s:e].copy_(tmp[: e - s])
    del tmp
    torch.cuda.synchronize()
    lib.cn_init(local_rank)

    def step():
        _lib.check(lib.cn_n_to_bits_host(h_n.data_ptr(), L, h_bits.data_ptr()))
        _lib.check(lib.cn_bits_to_n_host(h_bits.data_ptr(), W, L, h_out.data_ptr()))

    step()                                                # warm-up (staging allocation, page touching)
    steps = max(2, min(args.steps, args.e2e_steps))
    barrier()
    t0 = time.perf_counter()
    t_enc = t_dec = 0.0
    for _ in range(steps):
        a = time.perf_counter()
        _lib.check(lib.cn_n_to_bits_host(h_n.data_ptr(), L, h_bits.data_ptr()))
        b = time.perf_counter()
        _lib.check(lib.cn_bits_to_n_host(h_bits.data_ptr(), W, L, h_out.data_ptr()))
        c = time.perf_counter()
        t_enc += b - a
        t_dec += c - b
    barrier()
    dt = time.perf_counter() - t0
    tt = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    dt = float(tt.item())
    # Extra, clearly separate figure: the same C-ABI calls from TWO host threads, encode of batch k overlapped with
    # decode of batch k-1, so both PCIe directions carry their dominant stream at once (each call alone leaves one
    # direction 75 % idle).  Every step still performs one full encode and one full decode with all copies timed.
    duplex = None
    if not args.no_duplex:
        h_bits2 = torch.empty(W, dtype=torch.int64).pin_memory()
        bufs = [h_bits, h_bits2]
        errs = []

        def enc_job(k):
            try:
                _lib.check(lib.cn_n_to_bits_host(h_n.data_ptr(), L, bufs[k & 1].data_ptr()))
            except Exception as e:          # noqa: BLE001
                errs.append(e)

        def dec_job(k):
            try:
                _lib.check(lib.cn_bits_to_n_host(bufs[k & 1].data_ptr(), W, L, h_out.data_ptr()))
            except Exception as e:          # noqa: BLE001
                errs.append(e)

        # two persistent host threads (each owns one thread-local staging pipeline), stepped in lock-step
        n_phases = {"n": 0}
        gate = threading.Barrier(3)

        def enc_worker():
            while True:
                gate.wait()
                n = n_phases["n"]
                if n < 0:
                    return
                for p in range(n + 1):
                    if p < n:
                        enc_job(p)
                    gate.wait()

        def dec_worker():
            while True:
                gate.wait()
                n = n_phases["n"]
                if n < 0:
                    return
                for p in range(n + 1):
                    if p > 0:
                        dec_job(p - 1)
                    gate.wait()

        workers = [threading.Thread(target=enc_worker, daemon=True), threading.Thread(target=dec_worker, daemon=True)]
        for wk in workers:
            wk.start()

        def run_phases(n):
            n_phases["n"] = n
            gate.wait()                                       # release both workers
            for _ in range(n + 1):
                gate.wait()                                   # end of each phase

        run_phases(1)                                         # warm-up: creates both thread-local pipelines
        barrier()
        d0 = time.perf_counter()
        run_phases(steps)
        barrier()
        ddt = time.perf_counter() - d0
        n_phases["n"] = -1
        gate.wait()
        for wk in workers:
            wk.join()
        td = torch.tensor([ddt], dtype=torch.float64, device=dev)
        if world > 1:
            import torch.distributed as dist
            dist.all_reduce(td, op=dist.ReduceOp.MAX)
        if not errs:
            duplex = {"value": world * L * steps / float(td.item()), "unit": UNIT, "steps": steps,
                      "schedule": "2 host threads: cn_n_to_bits_host(batch k) overlapped with cn_bits_to_n_host(batch k-1); "
                                  "timed region = steps+1 phases incl. the lone first encode and last decode"}
    # the result read back on the host is the decoded sequence: check it against the canonical input
    lut = np.zeros(256, dtype=np.uint8)
    for ch, canon in zip(b"ACGTUacgtu", b"ACGTTACGTT"):
        lut[ch] = canon
    span = min(L, 1 << 26)
    ok = bool(np.array_equal(h_out[:span].numpy(), lut[h_n[:span].numpy()])) and \
        bool(np.array_equal(h_out[L - span:].numpy(), lut[h_n[L - span:].numpy()]))
    res = {"value": world * L * steps / dt, "unit": UNIT,
           "h2d_bytes_per_step": L + W * 8, "d2h_bytes_per_step": W * 8 + L,
           "nucleotides_per_gpu": L, "steps": steps, "ms_per_step": dt / steps * 1e3,
           "encode_nt_per_s": world * L * steps / t_enc, "decode_nt_per_s": world * L * steps / t_dec,
           "api": "cn_n_to_bits_host + cn_bits_to_n_host on pinned host buffers (C ABI; H2D + kernel + D2H inside the timed region)",
           "pcie_gbs": (2 * L + 2 * W * 8) * steps / dt / 1e9, "verified": ok}
    if duplex:
        res["duplex"] = duplex
    if note:
        res["note"] = note
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--nucleotides", type=int, default=10 * GIB, help="nucleotides per GPU per step (default 10 GiB)")
    ap.add_argument("--total-gib", type=float, default=0.0,
                    help="strong scaling (BASELINE config 5): shard this many GiB of ONE sequence over the ranks instead of "
                         "10 GiB per rank; decode writes back over the input when two ASCII buffers do not fit")
    ap.add_argument("--alphabet", type=int, default=10, choices=[4, 10])
    ap.add_argument("--e2e-nucleotides", type=int, default=0, help="e2e batch (default: same as --nucleotides)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--cpu-sample", type=int, default=1 * GIB, help="nucleotides in the bounded CPU sample")
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-assemble", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-slice leg (profiling runs only)")
    ap.add_argument("--no-duplex", action="store_true", help="skip the two-thread pipelined e2e figure")
    ap.add_argument("--no-extras", action="store_true", help="skip the checked-encode / base-5 extras")
    args = ap.parse_args()

    rank, local_rank, world = env_int("RANK", 0), env_int("LOCAL_RANK", 0), env_int("WORLD_SIZE", 1)
    if args.gpus != world and world == 1 and args.gpus > 1:
        raise SystemExit(f"--gpus {args.gpus} needs torchrun: python -m torch.distributed.run --nproc-per-node {args.gpus} bench.py --gpus {args.gpus}")
    args.gpus = world
    if args.impl == "reference":
        run_reference(args, rank)
        return
    run_gpu(args, rank, local_rank, world)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
