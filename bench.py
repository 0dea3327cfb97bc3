#!/usr/bin/env python
"""bench.py -- the hot path of cute-nucleotides on B200: 2-bit encode (n_to_bits) + decode (bits_to_n).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One STEP = one encode pass (ASCII -> packed u64) followed by one decode pass (packed -> ASCII) over one
batch of L synthetic nucleotides per GPU (BASELINE config 3+4 shape: 10 GiB, mixed case with U), buffers
resident in HBM.  `value` = nucleotides carried through the round trip per second, summed over GPUs
(weak scaling: every rank owns its own L-nucleotide shard of one logical sequence; the path needs no
data-path collective).  `e2e` = the same round trip through the host-slice C-ABI entry points
(cn_n_to_bits_host / cn_bits_to_n_host) with pinned HOST buffers, H2D and D2H inside the timed region: ONE call per
direction from rank 0, fanned out inside the library over the N GPUs of the job (cn_set_devices) -- N PCIe links.
Beside it: the same calls on pageable buffers, the calls made per rank (one process per GPU), and a two-thread duplex schedule.

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
    # the reference's OWN bench shape: one call per sequence on one thread, output allocated inside the timed region
    # (benches/bench_n_to_bits.rs:6-7, 10-19, 38-47), for the two batch shapes of extras.batch_small_sequences
    small = {}
    for name, offs in batch_shapes(np).items():
        buf = n[: int(offs[-1])]
        orc.time_small_calls(buf, offs, 1, True)
        t_rt = orc.time_small_calls(buf, offs, 2, True) / 2
        t_en = orc.time_small_calls(buf, offs, 2, False) / 2
        small[name] = {"sequences": int(offs.size - 1), "nucleotides": int(offs[-1]), "encode_nt_per_s": int(offs[-1]) / t_en,
                       "roundtrip_nt_per_s": int(offs[-1]) / t_rt, "us_per_call_encode": t_en / (offs.size - 1) * 1e6}
    return {
        "small_calls_single_thread": small,
        "single_thread": {"encode_nt_per_s": sample_nt / min(one_enc), "decode_nt_per_s": sample_nt / min(one_dec),
                          "value": sample_nt / (min(one_enc) + min(one_dec))},
        "value": sample_nt / statistics.mean(step), "best": sample_nt / min(step), "reps": len(step),
        "encode_nt_per_s": sample_nt / statistics.mean(t_enc), "decode_nt_per_s": sample_nt / statistics.mean(t_dec),
        "cores": threads, "variants": f"n_to_bits_{enc_v} + bits_to_n_{dec_v}", "step_s": step,
    }


def batch_shapes(np):
    """offsets of the two small-sequence batches: the reference's bench shape (40 000 nt per call, benches/bench_n_to_bits.rs:10)
    and short reads of 150-300 nt"""
    rng = np.random.default_rng(SEED)
    reads = rng.integers(150, 301, size=400000).astype(np.uint64)
    shapes = {"40000nt_x4096": np.arange(4097, dtype=np.uint64) * np.uint64(40000)}
    offs = np.zeros(reads.size + 1, dtype=np.uint64)
    np.cumsum(reads, out=offs[1:])
    shapes["reads_150_300nt_x400000"] = offs
    return shapes


def cpu_baseline_leg(args, L):
    """The `cpu_baseline` object of the GPU arm's line (guarded: never fatal)."""
    sample = min(L, args.cpu_sample if args.cpu_sample > 0 else GIB)
    try:
        r = cpu_roundtrip(sample, args.cpu_seconds, 5, args.alphabet)
    except Exception as e:                  # noqa: BLE001
        return {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": "failed", "error": repr(e)}
    return {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port",
            "sample": f"{sample} nt ({sample / GIB:.2f} GiB) of the workload x {r['reps']} reps, {r['variants']} "
                      f"(AVX2 restatement of the reference's fastest variants; Rust crate not buildable here), "
                      f"sharded by offset over {r['cores']} host threads, outputs pre-faulted",
            "encode_nt_per_s": r["encode_nt_per_s"], "decode_nt_per_s": r["decode_nt_per_s"], "best": r["best"],
            "single_thread": r["single_thread"], "small_calls_single_thread": r["small_calls_single_thread"]}


def run_reference(args, rank):
    """The reference's own fastest CPU variants on all host threads, on the SAME config as the GPU arm: every step is one
    encode + one decode of the full `--nucleotides` batch (10 GiB by default; --cpu-sample N bounds it for small hosts)."""
    if rank != 0:
        return
    sample = args.nucleotides if args.cpu_sample <= 0 else min(args.nucleotides, args.cpu_sample)
    steps, warmup = args.steps, args.warmup
    import numpy as np
    import _oracle
    orc = _oracle.Oracle()
    threads = os.cpu_count() or 1
    n = orc.generate(sample, SEED, args.alphabet)
    words = np.zeros(orc.words_for_len(sample), dtype=np.uint64)
    out = np.zeros(sample, dtype=np.uint8)
    enc_v, dec_v = ("movemask", "shuffle") if orc.simd_ok else ("lut", "lut")
    # bound the whole run to a few minutes: cap steps by a time budget
    t_enc = t_dec = 0.0
    done = 0
    budget_s = 150.0
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
    assert out[: 1 << 20].tobytes() == orc.canonical(n[: 1 << 20])
    # the reference as shipped is single-threaded: the same two variants on ONE thread, for context (1 GiB sample)
    one = min(sample, GIB)
    t0 = time.perf_counter()
    orc.encode_mt(n[:one], enc_v, 1, out=words[: orc.words_for_len(one)])
    t1 = time.perf_counter()
    orc.decode_mt(words[: orc.words_for_len(one)], one, dec_v, 1, out=out[:one])
    t2 = time.perf_counter()
    ms_step = (t_enc + t_dec) / done * 1e3
    value = sample / (ms_step / 1e3)
    note = "full workload per step" if sample == args.nucleotides else f"bounded sample of {sample} nt per step"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": done,
        "warmup": warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic",
        "config": shared_config(args),
        "encode_nt_per_s": sample * done / t_enc, "decode_nt_per_s": sample * done / t_dec,
        "single_thread": {"encode_nt_per_s": one / (t1 - t0), "decode_nt_per_s": one / (t2 - t1), "value": one / (t2 - t0),
                          "note": "the reference as shipped (one thread), 1 GiB sample"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{sample} nt ({sample / GIB:.2f} GiB) per step ({note}), {enc_v}+{dec_v} AVX2 restatement "
                                   "of the reference (Rust crate not buildable here: no rustc/cargo), sharded by offset over all host threads, "
                                   "rank 0 only"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def shared_config(args):
    """The `config` object: identical for the GPU arm and the --impl reference arm (same workload, same step)."""
    return {"workload": workload_name(args), "nucleotides_per_gpu": args.nucleotides, "alphabet": args.alphabet, "seed": SEED,
            "step": "n_to_bits encode of the whole batch, then bits_to_n decode of the encoder's output (a verifiable round trip; "
                    "decode is data-independent -- config 4's random-word decode is timed in extras.decode_random_words)",
            "sharding": "contiguous by sequence offset, one shard per GPU, no data-path collective",
            "l2": "inputs (>= 2.5 GiB per kernel) exceed the 126 MB L2; no flush needed"}


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
    cpu_group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        cpu_group = dist.new_group(backend="gloo")       # a barrier that keeps the GPUs idle while rank 0 drives all of them

    def cpu_barrier():
        if cpu_group is not None:
            dist.barrier(group=cpu_group)

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

    # ---- assemble: the only exchange the path can use (north_star: "a single NCCL broadcast/gather ... to assemble the
    # result").  Not part of `value` (the path itself has no exchange step); reported as value_with_assemble.
    assemble = None
    if world > 1 and not args.no_assemble:
        try:
            assemble = run_assemble(args, cn, torch, dist, sharded, d_n, d_bits, d_out, L, W, world, rank, offset, strong,
                                    enc_avg, dec_avg, barrier, dev)
        except Exception as e:              # noqa: BLE001  (reported, never fatal for the bench line)
            assemble = {"error": repr(e)}

    # ---- e2e: host-slice C-ABI calls with pinned host buffers, copies inside the timed region ----------
    # auxiliary legs are guarded: a failure there (e.g. pinned-memory exhaustion) is reported inside the line, it must
    # never cost the device-resident measurement above
    e2e = None
    if not args.no_e2e:
        try:
            e2e = run_e2e(args, cn, lib, _lib, torch, np, rank, local_rank, world, barrier, cpu_barrier, dev)
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
        "config": shared_config(args) if not strong else dict(shared_config(args), workload=f"{args.total_gib:g} GiB random sequence sharded by offset over {world} GPU(s): encode then decode (BASELINE config 5)", nucleotides_per_gpu=L),
        "decode_in_place": bool(in_place), "tuning": tuning,
        "encode_nt_per_s": total_nt / (enc_avg * 1e-3), "decode_nt_per_s": total_nt / (dec_avg * 1e-3),
        "encode_ms": enc_avg, "decode_ms": dec_avg,
        "roofline": {"bound": "hbm", "kernel": f"{dom}_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak,
                     "traffic": (traffic or {}).get(dom) if L == 10 * GIB else None,   # the ncu capture is of 10 GiB launches
                     "traffic_source": "committed ncu capture of the same 10 GiB launch (profiles/traffic.json: dram__bytes_read.sum + "
                                       "dram__bytes_write.sum per launch), not re-measured in this run",
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
    """Not part of `value`: the rows built after the 2-bit path on the same buffers, device-resident, CUDA
    events.  decode_random_words = BASELINE config 4 literally (335 544 320 random u64 -> 10 GiB ASCII);
    encode_checked / encode_lut_exact = n_to_bits + validation / n_to_bits_lut-exact semantics in one pass;
    base-5 = n_to_bits2 / bits_to_n2 (27 nt per u64, 35/27 algorithmic bytes per nucleotide) and its checked variant;
    packed-word ops = Hamming distance (read-only, 0.5 B/nt), complement, reverse complement (0.5 B/nt)."""
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

    dev = d_n.device
    W = cn.words_for_len(L)
    gbs = lambda bytes_, ms: bytes_ / (ms * 1e-3) / 1e9                       # noqa: E731
    out = {}
    # config 4: decode random words
    d_rand = cn.generate_words_device(torch.empty(W, dtype=torch.int64, device=dev), 0, SEED + 4)
    ms = timed(lambda: cn.decode_device(d_rand, L, out=d_out))
    win = 1 << 22
    codes = (d_rand[: win // 32].view(-1, 1) >> (torch.arange(32, device=dev, dtype=torch.int64) * 2)) & 3
    letters = torch.tensor(list(b"ACTG"), dtype=torch.uint8, device=dev)[codes.reshape(-1)]
    out["decode_random_words"] = {"ms": ms, "nt_per_s": L / (ms * 1e-3), "gbs": gbs(BYTES_PER_NT * L, ms),
                                  "verified": bool(torch.equal(d_out[:win], letters)),
                                  "note": "BASELINE config 4: 335 544 320 random u64 words -> 10 GiB ASCII"}
    # packed-word ops on two word streams (parity unpinned: no reference code, see include/cute_nucleotides_cuda.h)
    d_other = cn.generate_words_device(torch.empty(W, dtype=torch.int64, device=dev), 0, SEED + 5)
    res = torch.zeros(1, dtype=torch.int64, device=dev)
    ms = timed(lambda: cn.hamming_device(d_rand, d_other, L, result=res))
    out["hamming"] = {"ms": ms, "nt_per_s": L / (ms * 1e-3), "gbs": gbs(0.5 * L, ms), "bytes_per_nt": 0.5, "pattern": "pure read"}
    ms = timed(lambda: cn.complement_device(d_rand, L, out=d_other))
    out["complement"] = {"ms": ms, "nt_per_s": L / (ms * 1e-3), "gbs": gbs(0.5 * L, ms), "bytes_per_nt": 0.5, "pattern": "1:1 copy"}
    ms = timed(lambda: cn.reverse_complement_device(d_rand, L, out=d_other))
    out["reverse_complement"] = {"ms": ms, "nt_per_s": L / (ms * 1e-3), "gbs": gbs(0.5 * L, ms), "bytes_per_nt": 0.5, "pattern": "1:1 copy, reversed"}
    del d_rand, d_other
    counter = torch.zeros(1, dtype=torch.int64, device=dev)
    ms_checked = timed(lambda: cn.encode_checked_device(d_n, counter, out=d_bits))
    out["encode_checked"] = {"ms": ms_checked, "gbs": gbs(BYTES_PER_NT * L, ms_checked), "invalid_bytes_found": int(counter.item())}
    ms = timed(lambda: cn.encode_ex_device(d_n, cn.ENC_LUT_EXACT, out=d_bits))
    out["encode_lut_exact"] = {"ms": ms, "gbs": gbs(BYTES_PER_NT * L, ms), "note": "bit-exact n_to_bits_lut on any bytes; valid data here"}
    cn.generate2_device(d_n, 0, SEED, 12)                        # 5-letter data for the base-5 codec, same buffer
    W2 = cn.words2_for_len(L)
    d_bits2 = torch.empty(W2, dtype=torch.int64, device=dev)
    b2 = L + 8 * W2
    ms_e2 = timed(lambda: cn.encode2_device(d_n, out=d_bits2))
    out["base5_encode"] = {"ms": ms_e2, "nt_per_s": L / (ms_e2 * 1e-3), "gbs": gbs(b2, ms_e2)}
    counter.zero_()
    ms = timed(lambda: cn.encode2_ex_device(d_n, cn.ENC_COUNT, counter=counter, out=d_bits2))
    out["base5_encode_checked"] = {"ms": ms, "nt_per_s": L / (ms * 1e-3), "gbs": gbs(b2, ms), "invalid_bytes_found": int(counter.item())}
    ms_d2 = timed(lambda: cn.decode2_device(d_bits2, L, out=d_out))
    out["base5_decode"] = {"ms": ms_d2, "nt_per_s": L / (ms_d2 * 1e-3), "gbs": gbs(b2, ms_d2)}
    out["note"] = "device-resident, CUDA events, same 10 GiB buffers; not part of `value`"
    # many small sequences per call (host buffers, pageable, everything timed): cn_n_to_bits_host_batch / cn_bits_to_n_host_batch
    import numpy as np
    batch = {}
    for name, offs in batch_shapes(np).items():
        total = int(offs[-1])
        buf = cn.generate_device(torch.empty(total, dtype=torch.uint8, device=dev), 0, SEED, 10).cpu().numpy()
        lens = np.diff(offs)
        words, woffs = cn.n_to_bits_batch_cuda(buf, offs)                   # warm-up
        back, _ = cn.bits_to_n_batch_cuda(words, woffs, lens)
        t_en = t_rt = 1e30
        for _ in range(3):                                                  # outputs reused like the warm heap of the CPU loop
            a = time.perf_counter()
            words, woffs = cn.n_to_bits_batch_cuda(buf, offs, out=words)
            b = time.perf_counter()
            back, _ = cn.bits_to_n_batch_cuda(words, woffs, lens, out=back)
            c = time.perf_counter()
            t_en, t_rt = min(t_en, b - a), min(t_rt, c - a)
        i = int(offs.size // 2)
        one = cn.n_to_bits_cuda(buf[int(offs[i]): int(offs[i + 1])])
        batch[name] = {"sequences": int(offs.size - 1), "nucleotides": total, "encode_nt_per_s": total / t_en, "roundtrip_nt_per_s": total / t_rt,
                       "us_per_sequence_encode": t_en / (offs.size - 1) * 1e6,
                       "matches_single_call": bool(np.array_equal(words[int(woffs[i]): int(woffs[i + 1])], one))}
    # the same two shapes RESIDENT IN HBM, tightly concatenated (cn_encode_segmented_device / cn_decode_segmented_device):
    # algorithmic traffic ~1.25 B/nt + 16 B of offsets per sequence per direction
    seg = {}
    for name, (lo, hi, n_seq) in {"40000nt_x25000": (40000, 40001, 25000), "reads_150_300nt_x4000000": (150, 301, 4000000)}.items():
        lens_t = torch.randint(lo, hi, (n_seq,), device=dev, dtype=torch.int64, generator=torch.Generator(device=dev).manual_seed(SEED))
        offs_t = torch.zeros(n_seq + 1, dtype=torch.int64, device=dev)
        torch.cumsum(lens_t, dim=0, out=offs_t[1:])
        total = int(offs_t[-1].item())
        d_seq = cn.generate_device(d_n[:total], 0, SEED, 10)
        woffs_t = cn.segment_word_offsets(offs_t)
        d_w = d_bits[: int(woffs_t[-1].item())]
        ms_e = timed(lambda: cn.encode_segmented_device(d_seq, offs_t, woffs_t, out=d_w))
        ms_d = timed(lambda: cn.decode_segmented_device(d_w, offs_t, woffs_t, out=d_out[:total]))
        i = n_seq // 2
        s_, e_ = int(offs_t[i].item()), int(offs_t[i + 1].item())
        one = cn.encode_device(d_seq[s_:e_].clone())
        ok = bool(torch.equal(d_w[int(woffs_t[i].item()): int(woffs_t[i + 1].item())], one))
        seg[name] = {"sequences": n_seq, "nucleotides": total, "encode_ms": ms_e, "decode_ms": ms_d,
                     "encode_nt_per_s": total / (ms_e * 1e-3), "decode_nt_per_s": total / (ms_d * 1e-3),
                     "encode_gbs": gbs(1.25 * total + 16 * n_seq, ms_e), "decode_gbs": gbs(1.25 * total + 16 * n_seq, ms_d),
                     "matches_single_call": ok}
    out["segmented_device"] = seg
    a = time.perf_counter()
    for _ in range(200):
        cn.n_to_bits_cuda(buf[:40000])
    batch["one_call_per_sequence_40000nt_us"] = (time.perf_counter() - a) / 200 * 1e6
    batch["note"] = ("host-resident pageable sequences; the Python wrapper's pointer arithmetic is inside the timed region, output arrays are "
                     "reused across calls (the CPU loop's malloc/free recycles a warm heap block the same way); "
                     "compare cpu_baseline.small_calls_single_thread (the reference's bench shape on one core)")
    out["batch_small_sequences"] = batch
    return out


def torch_pack(torch, x):
    """Independent torch computation of the contract (src/n_to_bits.rs:39-42): code = (byte >> 1) & 3 on the valid alphabet,
    nucleotide i at bits 2(i & 31) of word i >> 5.  x: uint8 CUDA tensor whose length is a multiple of 32."""
    shifts = torch.arange(32, device=x.device, dtype=torch.int64) * 2
    codes = ((x >> 1) & 3).view(-1, 32).to(torch.int64)
    return (codes << shifts).sum(dim=1)                          # disjoint bit fields: sum == or; wraps into int64


def verify_whole_packed(cn, torch, full, first_nt, total_nt, alphabet):
    """EVERY word of `full` (packed words of nucleotides [first_nt, first_nt + total_nt) of the synthetic sequence) against
    torch_pack of the regenerated ASCII, in 128 Mi-nucleotide pieces.  total_nt must be a multiple of 32."""
    piece = 1 << 27
    tmp = torch.empty(piece, dtype=torch.uint8, device=full.device)
    ok = True
    for s in range(0, total_nt, piece):
        n = min(piece, total_nt - s)
        cn.generate_device(tmp[:n], first_nt + s, SEED, alphabet)
        ok &= bool(torch.equal(full[s // 32: (s + n) // 32], torch_pack(torch, tmp[:n])))
    return ok


def verify_whole_decoded(cn, torch, d_out, first_nt, alphabet):
    """every byte of d_out == canonical form of the regenerated input"""
    lut = torch.zeros(256, dtype=torch.uint8, device=d_out.device)
    for ch, canon in zip(b"ACGTUacgtu", b"ACGTTACGTT"):
        lut[ch] = canon
    piece = 1 << 28
    tmp = torch.empty(piece, dtype=torch.uint8, device=d_out.device)
    ok = True
    for s in range(0, d_out.numel(), piece):
        n = min(piece, d_out.numel() - s)
        cn.generate_device(tmp[:n], first_nt + s, SEED, alphabet)
        ok &= bool(torch.equal(d_out[s:s + n], lut[tmp[:n].long()]))
    return ok


def run_assemble(args, cn, torch, dist, sharded, d_n, d_bits, d_out, L, W, world, rank, offset, strong, enc_avg, dec_avg, barrier, dev):
    """Every way of assembling the packed shards, timed with CUDA events (max over ranks) and verified over the WHOLE
    assembled buffer on every rank that holds one:
      nccl all-gather | nccl gather-to-root | fused encode+all-gather kernel | fused encode+gather-to-root kernel |
      nccl scatter + decode | fused scatter+decode kernel (peer loads from the root)."""
    if strong or L % (1 << 20):
        return {"skipped": "assemble legs run on the weak-scaling shape (equal, 1 Mi-aligned shards)"}
    total = world * L
    root = 0

    def timed(fn, reps=2):
        fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def all_true(flag):
        t = torch.tensor([1 if flag else 0], dtype=torch.int64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return bool(t.item())

    res = {"bytes_per_rank": W * 8, "root": root}
    cn.encode_device(d_n, out=d_bits)
    # 1. NCCL all-gather (every rank ends with all shards)
    full = torch.empty(W * world, dtype=torch.int64, device=dev)
    ms = timed(lambda: dist.all_gather_into_tensor(full, d_bits))
    res["nccl_all_gather"] = {"ms": ms, "busbw_gbs": W * 8 * (world - 1) / (ms * 1e-3) / 1e9,
                              "whole_buffer_verified_on_every_rank": all_true(verify_whole_packed(cn, torch, full, 0, total, args.alphabet))}
    # 2. NCCL gather-to-root (grouped send/recv): only the root ingests
    holder = {}

    def nccl_gather():
        holder["g"] = sharded.gather_packed(d_bits, total, root=root, granule=1 << 20)
    full.fill_(0)
    del full
    ms = timed(nccl_gather)
    ok = verify_whole_packed(cn, torch, holder["g"], 0, total, args.alphabet) if rank == root else True
    res["nccl_gather_to_root"] = {"ms": ms, "root_ingress_gbs": W * 8 * (world - 1) / (ms * 1e-3) / 1e9, "whole_buffer_verified_on_root": all_true(ok)}
    # 3. NCCL scatter from the root + decode
    def nccl_scatter_decode():
        mine = sharded.scatter_packed(holder.get("g"), total, root=root, granule=1 << 20, like=d_bits)
        cn.decode_device(mine, L, out=d_out)
    ms = timed(nccl_scatter_decode)
    res["nccl_scatter_then_decode"] = {"ms": ms, "decoded_verified_on_every_rank": all_true(verify_whole_decoded(cn, torch, d_out, offset, args.alphabet))}
    holder.clear()
    # 4.-6. fused kernels over NVLink peer memory (CUDA IPC mappings of every rank's full-size buffer)
    asm = sharded.PeerAssembly(total, granule=1 << 20)
    asm.full.fill_(-1)
    barrier()
    ms = timed(lambda: asm.encode(d_n))
    barrier()
    res["fused_encode_all_gather"] = {"ms": ms, "egress_gbs_per_gpu": W * 8 * (world - 1) / (ms * 1e-3) / 1e9,
                                      "op": "cn_encode_multi_device: ONE kernel per rank reads its shard and stores the packed words into every rank's buffer",
                                      "whole_buffer_verified_on_every_rank": all_true(verify_whole_packed(cn, torch, asm.full, 0, total, args.alphabet))}
    asm.full.fill_(-1)
    barrier()
    ms_gather = timed(lambda: asm.encode(d_n, dests=[root]))
    barrier()
    ok = verify_whole_packed(cn, torch, asm.full, 0, total, args.alphabet) if rank == root else bool((asm.full == -1).all())
    res["fused_encode_gather_to_root"] = {"ms": ms_gather, "root_ingress_gbs": W * 8 * (world - 1) / (ms_gather * 1e-3) / 1e9,
                                          "op": "cn_encode_multi_device with one destination: the root's buffer",
                                          "whole_buffer_verified_on_root_and_untouched_elsewhere": all_true(ok)}
    d_out.fill_(0)
    ms_scatter = timed(lambda: asm.decode_from(root, out=d_out))
    barrier()
    res["fused_scatter_decode"] = {"ms": ms_scatter, "root_egress_gbs": W * 8 * (world - 1) / (ms_scatter * 1e-3) / 1e9,
                                   "op": "cn_decode_device reading this rank's words straight out of the root's buffer (peer loads)",
                                   "decoded_verified_on_every_rank": all_true(verify_whole_decoded(cn, torch, d_out, offset, args.alphabet))}
    asm.close()
    del asm
    step_ms = ms_gather + ms_scatter
    res["value_with_assemble"] = total / (step_ms * 1e-3)
    res["value_with_assemble_def"] = "nucleotides / (fused encode+gather-to-root + fused scatter+decode): the whole packed sequence passes through ONE GPU"
    res["assemble_efficiency"] = (enc_avg + dec_avg) / step_ms
    link = W * 8 * (world - 1) / 1e9
    res["limiter"] = (f"the root GPU's NVLink port: {link:.1f} GB in and {link:.1f} GB out per step against ~{enc_avg + dec_avg:.1f} ms of kernels; "
                      "gathering a sharded result into ONE GPU is link-bound by construction (ideal at 900 GB/s per direction: "
                      f"{2 * link / 900 * 1e3:.1f} ms)")
    res["unfused_encode_plus_allgather_ms"] = enc_avg + res["nccl_all_gather"]["ms"]
    return res


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


def run_e2e(args, cn, lib, _lib, torch, np, rank, local_rank, world, barrier, cpu_barrier, dev):
    """The round trip through the host-slice C-ABI calls, host buffers, every copy inside the timed region.

    value           ONE cn_n_to_bits_host + ONE cn_bits_to_n_host call per step from rank 0 on PINNED buffers holding the whole
                    N x L batch, fanned out inside the library over the N GPUs of the job (cn_set_devices): N PCIe links.
                    (N = 1: one GPU, one link.)
    per_rank_calls  the r01 shape: every rank calls for its own shard on its own GPU (N processes, N links)   [N > 1]
    pageable        the same calls on pageable (numpy / Vec-like) buffers: the literal drop-in shape
    duplex          ONE host thread, asynchronous calls: encode of batch k+1 in flight while batch k decodes   [N == 1]
    all_visible     fan-out over every visible GPU when more are visible than the job was given               [N == 1]
    """
    import ctypes
    import psutil
    Lg = args.e2e_nucleotides or (args.nucleotides if world == 1 else min(args.nucleotides, (4 if world == 2 else 2) * GIB))
    avail = psutil.virtual_memory().available
    note = None
    while 2.3 * Lg * world > 0.45 * avail and Lg > (1 << 28):
        Lg //= 2
        note = "e2e batch reduced to fit pinned host memory"
    Lg &= ~((1 << 20) - 1)
    steps = max(2, min(args.steps, args.e2e_steps))
    lut = np.zeros(256, dtype=np.uint8)
    for ch, canon in zip(b"ACGTUacgtu", b"ACGTTACGTT"):
        lut[ch] = canon

    def fill(h_n, first_nt):
        """the shared synthetic sequence, generated on the device and copied out once (untimed)"""
        tmp = torch.empty(min(h_n.numel(), 1 << 30), dtype=torch.uint8, device=dev)
        for s in range(0, h_n.numel(), tmp.numel()):
            e = min(h_n.numel(), s + tmp.numel())
            cn.generate_device(tmp[: e - s], first_nt + s, SEED + 1, args.alphabet)
            h_n[s:e].copy_(tmp[: e - s])
        torch.cuda.synchronize()

    def check(h_n, h_out, n_total, marks=()):
        span = min(n_total, 1 << 26)
        a, o = (h_n.numpy(), h_out.numpy()) if hasattr(h_n, "numpy") else (h_n, h_out)
        ok = bool(np.array_equal(o[:span], lut[a[:span]])) and bool(np.array_equal(o[n_total - span:n_total], lut[a[n_total - span:n_total]]))
        for m in marks:                                      # windows straddling the device boundaries of a fanned-out call
            lo, hi = max(0, m - (1 << 20)), min(n_total, m + (1 << 20))
            ok = ok and bool(np.array_equal(o[lo:hi], lut[a[lo:hi]]))
        return ok

    def timed_roundtrip(n_ptr, bits_ptr, out_ptr, nt, everyone):
        """`steps` x (encode call, decode call); returns (seconds, encode seconds, decode seconds).  everyone: all ranks
        run it concurrently (max over ranks); else rank 0 alone while the others wait on a CPU barrier (GPUs idle)."""
        W_ = cn.words_for_len(nt)
        _lib.check(lib.cn_n_to_bits_host(n_ptr, nt, bits_ptr))          # warm-up: staging allocation, page touching, worker threads
        _lib.check(lib.cn_bits_to_n_host(bits_ptr, W_, nt, out_ptr))
        if everyone:
            barrier()
        t0 = time.perf_counter()
        t_enc = t_dec = 0.0
        for _ in range(steps):
            a = time.perf_counter()
            _lib.check(lib.cn_n_to_bits_host(n_ptr, nt, bits_ptr))
            b = time.perf_counter()
            _lib.check(lib.cn_bits_to_n_host(bits_ptr, W_, nt, out_ptr))
            c = time.perf_counter()
            t_enc += b - a
            t_dec += c - b
        if everyone:
            barrier()
        dt = time.perf_counter() - t0
        if everyone and world > 1:
            import torch.distributed as dist
            tt = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dt = float(tt.item())
        return dt, t_enc, t_dec

    def figures(nt, dt, t_enc, t_dec):
        W_ = cn.words_for_len(nt)
        return {"value": nt * steps / dt, "ms_per_step": dt / steps * 1e3, "encode_nt_per_s": nt * steps / t_enc,
                "decode_nt_per_s": nt * steps / t_dec, "host_traffic_gbs": (2 * nt + 2 * W_ * 8) * steps / dt / 1e9}

    res = {"unit": UNIT, "nucleotides_per_gpu": Lg, "steps": steps}
    total = world * Lg
    Wt = cn.words_for_len(total)
    lib.cn_init(local_rank)

    # ---- per-rank calls (N processes, each its own link); at N == 1 this IS the headline -----------------------
    n_mine = total if rank == 0 else Lg                      # rank 0 allocates the whole batch once and reuses it below
    h_n = torch.empty(n_mine, dtype=torch.uint8, pin_memory=True)
    h_bits = torch.empty(cn.words_for_len(n_mine), dtype=torch.int64, pin_memory=True)
    h_out = torch.empty(n_mine, dtype=torch.uint8, pin_memory=True)
    if rank == 0:
        fill(h_n, 0)
    else:
        fill(h_n, rank * Lg)
    cn.set_devices([])
    dt, te, td = timed_roundtrip(h_n.data_ptr(), h_bits.data_ptr(), h_out.data_ptr(), Lg, everyone=True)
    per_rank = figures(world * Lg, dt, te / 1, td / 1)
    per_rank["encode_nt_per_s"] *= 1                         # (already whole-job: every rank ran its shard concurrently)
    ok_rank = check(h_n, h_out, Lg)
    if world > 1:
        import torch.distributed as dist
        t_ok = torch.tensor([1 if ok_rank else 0], dtype=torch.int64, device=dev)
        dist.all_reduce(t_ok, op=dist.ReduceOp.MIN)
        ok_rank = bool(t_ok.item())
    per_rank.update(verified=ok_rank, api="cn_n_to_bits_host + cn_bits_to_n_host per rank on its own GPU, pinned buffers")

    # ---- ONE call from rank 0 over all N GPUs of the job --------------------------------------------------------
    if world == 1:
        head = dict(per_rank)
        head["api"] = "cn_n_to_bits_host + cn_bits_to_n_host on pinned host buffers (C ABI; H2D + kernel + D2H inside the timed region)"
    else:
        if rank != 0:
            del h_n, h_bits, h_out
        cpu_barrier()
        head = None
        if rank == 0:
            cn.set_devices(list(range(world)))
            dt, te, td = timed_roundtrip(h_n.data_ptr(), h_bits.data_ptr(), h_out.data_ptr(), total, everyone=False)
            head = figures(total, dt, te, td)
            head["verified"] = check(h_n, h_out, total, marks=[k * Lg for k in range(1, world)])
            head["api"] = (f"ONE cn_n_to_bits_host + ONE cn_bits_to_n_host call per step from rank 0 on the whole {world} x {Lg / GIB:g} GiB "
                           f"pinned batch, fanned out inside the library over the job's {world} GPUs (cn_set_devices): {world} PCIe links")
        cpu_barrier()
    if rank != 0:
        cn.set_devices([])
        return None
    res.update(head)
    W1 = cn.words_for_len(total)
    res["h2d_bytes_per_step"] = total + W1 * 8
    res["d2h_bytes_per_step"] = W1 * 8 + total
    res["gpus_used"] = world
    res["pcie_gbs"] = res["host_traffic_gbs"]
    if world > 1:
        res["per_rank_calls"] = per_rank

    # ---- pageable buffers: what a Rust &[u8] / Vec is ------------------------------------------------------------
    try:
        pg_n = h_n.numpy().copy()
        pg_bits = np.zeros(Wt, dtype=np.uint64)
        pg_out = np.zeros(total, dtype=np.uint8)                  # pre-faulted, like the CPU arm's outputs
        dt, te, td = timed_roundtrip(pg_n.ctypes.data, pg_bits.ctypes.data, pg_out.ctypes.data, total, everyone=False)
        pg = figures(total, dt, te, td)
        pg["verified"] = check(pg_n, pg_out, total, marks=[k * Lg for k in range(1, world)]) and \
            bool(np.array_equal(pg_bits.view(np.int64), h_bits.numpy()[:Wt]))
        pg["fraction_of_pinned"] = pg["value"] / res["value"]
        pg["note"] = ("pageable caller buffers are copied through pinned staging by the library's copier pool; bounded by host memory "
                      "bandwidth shared with the DMA engines (tools/host_ceiling `mix`, profiles/host_ceiling_r02_n1.jsonl)")
        res["pageable"] = pg
        del pg_n, pg_bits, pg_out
    except Exception as e:                  # noqa: BLE001
        res["pageable"] = {"error": repr(e)}
    cn.set_devices([])

    # ---- every visible GPU, when the job was given fewer (labelled; NOT the headline) ---------------------------------
    visible = torch.cuda.device_count()
    if world == 1 and visible > 1 and not args.no_all_visible:
        try:
            cn.set_devices(list(range(visible)))
            dt, te, td = timed_roundtrip(h_n.data_ptr(), h_bits.data_ptr(), h_out.data_ptr(), total, everyone=False)
            av = figures(total, dt, te, td)
            av.update(gpus_used=visible, verified=check(h_n, h_out, total),
                      note=f"the same two calls with cn_set_devices over all {visible} visible GPUs; the job itself was given 1 GPU, so this is NOT `value`")
            res["all_visible"] = av
        except Exception as e:              # noqa: BLE001
            res["all_visible"] = {"error": repr(e)}
        cn.set_devices([])

    # ---- duplex: the same C-ABI calls from TWO host threads, encode of batch k overlapped with decode of batch k-1, so both
    # PCIe directions carry their dominant stream at once (each call alone leaves one direction 75 % idle) --------------
    if world == 1 and not args.no_duplex:
        try:
            res["duplex"] = run_duplex(args, cn, lib, _lib, torch, np, h_n, h_bits, h_out, total, steps)
        except Exception as e:              # noqa: BLE001
            res["duplex"] = {"error": repr(e)}

    link = "one PCIe Gen5 x16 link: 55 GB/s one way, 45 + 45 GB/s both ways (tools/host_ceiling, profiles/host_ceiling_r02_n1.jsonl)"
    if world == 1:
        res["limiter"] = (f"{link}; a sequential encode-then-decode round trip moves 1.25 B per nucleotide with one direction "
                          "dominant at a time => <= 27 Gnt/s per link; the duplex schedule uses both directions => <= ~36 Gnt/s")
    else:
        per_gpu = res["host_traffic_gbs"] / world
        res["limiter"] = (f"{world} links carried {res['host_traffic_gbs']:.0f} GB/s of host traffic in aggregate ({per_gpu:.0f} GB/s per GPU; "
                          f"a lone link carries ~66 GB/s on this round trip).  Beyond 2 links the HOST side of the box is the bound, not the "
                          "links or the GPUs: pure cudaMemcpyAsync from pinned memory tops out at 71 / 74 / 94 GB/s device-to-host and "
                          "111 / 115 / 186 GB/s host-to-device over 2 / 4 / 8 links of the 8-GPU box (tools/host_ceiling `pcie`, "
                          "profiles/host_ceiling_r02_n8.jsonl), which caps decode at ~90 Gnt/s and the round trip at ~60 Gnt/s")
    if note:
        res["note"] = note
    return res


def run_duplex(args, cn, lib, _lib, torch, np, h_n, h_bits, h_out, L, steps):
    """ONE host thread, the asynchronous entry points: cn_n_to_bits_host_async(batch k+1) is in flight while
    cn_bits_to_n_host(batch k) runs, so both PCIe directions carry their dominant stream at once.  Every step still performs one
    full encode and one full decode with all copies timed; `value` includes the lone first encode."""
    import ctypes
    W = cn.words_for_len(L)
    h_bits2 = torch.empty(W, dtype=torch.int64, pin_memory=True)
    bufs = [h_bits, h_bits2]

    def enc_async(k):
        req = ctypes.c_void_p()
        _lib.check(lib.cn_n_to_bits_host_async(h_n.data_ptr(), L, bufs[k & 1].data_ptr(), ctypes.byref(req)))
        return req

    def run(n):
        phase = []
        req = enc_async(0)
        for k in range(n):
            t = time.perf_counter()
            _lib.check(lib.cn_wait(req))                                    # batch k is encoded
            if k + 1 < n:
                req = enc_async(k + 1)                                      # next encode starts ...
            _lib.check(lib.cn_bits_to_n_host(bufs[k & 1].data_ptr(), W, L, h_out.data_ptr()))   # ... while this decode runs
            phase.append(time.perf_counter() - t)
        return phase

    run(2)                                                # warm-up: creates the library threads' staging
    torch.cuda.synchronize()
    d0 = time.perf_counter()
    phase = run(steps)
    ddt = time.perf_counter() - d0
    steady = phase[1:-1]                                  # phases in which an encode AND a decode ran (the first waits for the lone encode)
    lut = np.zeros(256, dtype=np.uint8)
    for ch, canon in zip(b"ACGTUacgtu", b"ACGTTACGTT"):
        lut[ch] = canon
    span = min(L, 1 << 26)
    ok = bool(np.array_equal(h_out[:span].numpy(), lut[h_n[:span].numpy()])) and bool(np.array_equal(h_bits.numpy(), h_bits2.numpy()))
    return {"value": L * steps / ddt, "unit": UNIT, "steps": steps, "verified": ok,
            "steady_state_value": L * len(steady) / sum(steady) if steady else None,
            "schedule": "one host thread: cn_n_to_bits_host_async(batch k+1) in flight while cn_bits_to_n_host(batch k) runs; `value` is over "
                        "the whole loop incl. the lone first encode and last decode, `steady_state_value` over the overlapped phases only"}


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
    ap.add_argument("--e2e-steps", type=int, default=6)
    ap.add_argument("--cpu-sample", type=int, default=0,
                    help="nucleotides per step of the CPU legs (0: 1 GiB for the GPU arm's cpu_baseline, the full batch for --impl reference)")
    ap.add_argument("--cpu-seconds", type=float, default=10.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-assemble", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-slice leg (profiling runs only)")
    ap.add_argument("--no-duplex", action="store_true", help="skip the two-thread pipelined e2e figure")
    ap.add_argument("--no-all-visible", action="store_true", help="skip the fan-out over GPUs the job was not given (N == 1 only)")
    ap.add_argument("--no-extras", action="store_true", help="skip the checked-encode / base-5 extras")
    args = ap.parse_args()

    rank, local_rank, world = env_int("RANK", 0), env_int("LOCAL_RANK", 0), env_int("WORLD_SIZE", 1)
    if args.impl != "reference" and args.gpus != world and world == 1 and args.gpus > 1:
        raise SystemExit(f"--gpus {args.gpus} needs torchrun: python -m torch.distributed.run --nproc-per-node {args.gpus} bench.py --gpus {args.gpus}")
    if args.impl != "reference":
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
