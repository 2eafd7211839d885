#!/usr/bin/env python
"""Benchmark of the B200-native aoTuV/Vorbis per-block analysis path.

  python bench.py --gpus N --steps K --warmup W            (our arm)
  python bench.py --impl reference --gpus N --steps K ...  (the reference's own CPU path, host cores)

Metric (BASELINE.json): encoded audio-seconds per wall-second at q4 44.1 kHz stereo.
Workload: S independent 60 s synthetic streams (SURVEY.md section 8d signal), all of their blocks batched
per step.  A step = one pass of the hot path over that batch.
  value : inputs (the segmented PCM timeline + block descriptors) already resident in HBM when the
          timed region starts; the region ends when the packets are back on the host.
  e2e   : the public call (Encoder.encode) from HOST PCM: host segmentation (envelope / block-size
          decision), H2D of the PCM, all kernels incl. bit packing, D2H of the packets.
Prints ONE JSON line (rank 0).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")      # before any CUDA context exists (see vorbis_aotuv_lancer_b200/lib.py)

import numpy as np  # noqa: E402

CONFIG = "q4_44k_2ch"
QUALITY_NAME = {"q4_44k_2ch": "q4 44.1 kHz stereo", "q5_44k_2ch": "q5 44.1 kHz stereo", "q10_48k_2ch": "q10 48 kHz stereo",
                "q8_48k_6ch": "q8 48 kHz 5.1", "q2_44k_2ch": "q2 44.1 kHz stereo", "q4_48k_6ch": "q4 48 kHz 5.1",
                "q4_44k_1ch": "q4 44.1 kHz mono", "abr128_44k_2ch": "128 kbit/s ABR 44.1 kHz stereo"}


def workload_config(config, S, T):
    """The `config` object of the JSON line: the workload and nothing else, identical for both arms (the reference arm
    times bounded samples of it, see cpu_baseline.sample)."""
    from vorbis_aotuv_lancer_b200 import blob as vblob
    ch, rate, q, abr = vblob.CATALOGUE[config]
    pcm_mb = S * T * rate * ch * 4 / 1e6
    return {"workload": "%d independent %g s %s synthetic streams per GPU (SURVEY 8d signal, seed per stream), every block of every "
                        "stream through the whole analysis path per step" % (S, T, QUALITY_NAME.get(config, config)),
            "setup": config, "quality": q, "managed_bitrate": int(abr), "channels": ch, "rate": rate,
            "streams_per_gpu": S, "seconds_per_stream": T, "audio_seconds_per_gpu_step": S * T,
            "parallelism": "streams sharded i mod G, no collective",
            "l2": "inputs larger than L2: %.0f MB of float PCM per GPU and step (126 MB L2), no flush needed" % pcm_mb}
# algorithmic HBM bytes per ch-block (long N2=1024 / short N2=128), SURVEY.md section 8(d)
ALGO_BYTES = {
    "K1 transform": (20480, 2560), "K2 noisemask": (16512, 2112), "K3 tonemask": (8192, 1024),
    "K4 offset_and_mix": (32768, 4096), "K5 floor fit": (8308, 1076), "K5b floor encode": (4212, 564),
    "K6 couple": (16512, 2112), "K7 residue": (6200, 800), "K8 bitpack": (2400, 400),
}


def clocks_sampler(stop, out, gpu):
    q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    while not stop.is_set():
        try:
            r = subprocess.run(["nvidia-smi", "-i", str(gpu), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                               capture_output=True, text=True, timeout=5).stdout.strip().split(",")
            out.append([x.strip() for x in r])
        except Exception:
            pass
        stop.wait(0.2)


def summarize_clocks(samples):
    sm = [float(s[0]) for s in samples if len(s) >= 6 and s[0].replace(".", "").isdigit()]
    mx = [float(s[1]) for s in samples if len(s) >= 6 and s[1].replace(".", "").isdigit()]
    reasons = set()
    for s in samples:
        if len(s) >= 6:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
    return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons)}


def synth_streams(L, nstreams, seconds, channels, rate, first_index, threads=8, pinned=None):
    from concurrent.futures import ThreadPoolExecutor
    n = int(rate * seconds)

    def one(i):
        a = pinned[i].array if pinned is not None else np.empty((channels, n), np.float32)
        seed = L.vb200_splitmix64(C.c_uint64(88172645463325252 + first_index + i))
        L.vb200_synth_pcm(C.c_uint64(seed), channels, C.c_long(rate), C.c_long(n), a.ctypes.data_as(C.c_void_p), C.c_long(n))
        return a
    with ThreadPoolExecutor(max_workers=threads) as ex:
        return list(ex.map(one, range(nstreams)))


_REAL_STDOUT = None


def emit(out):
    f = _REAL_STDOUT or sys.stdout
    f.write(json.dumps(out) + "\n")
    f.flush()


# ------------------------------------------------------------------ reference arm / cpu baseline
def _ref_worker(args):
    flavour, channels, rate, quality, seconds, seed_index, reps, abr = args
    from oracle import reflib
    from vorbis_aotuv_lancer_b200.lib import L
    pcm = synth_streams(L, 1, seconds, channels, rate, seed_index, threads=1)[0]
    R = reflib.RefLib(flavour)
    t0 = time.perf_counter()
    nbytes = 0
    for _ in range(reps):
        h = R.open(channels, rate, quality, abr)
        R.write(h, pcm)
        nbytes += sum(len(p[0]) for p in R.packets(h))
        R.close(h)
    return time.perf_counter() - t0, nbytes


def reference_throughput(cores, seconds, reps, flavour="sse", config=None):
    """The UNMODIFIED reference (oracle/_ref, Lancer SSE path, -O3 -ffast-math) one stream per core."""
    import multiprocessing as mp
    from vorbis_aotuv_lancer_b200 import blob as vblob
    ch, rate, q, abr = vblob.CATALOGUE[config or CONFIG]
    ctx = mp.get_context("spawn")
    with ctx.Pool(cores) as pool:
        pool.map(_ref_worker, [(flavour, ch, rate, q, 0.5, i, 1, abr) for i in range(cores)])      # warm the workers
        t0 = time.perf_counter()
        res = pool.map(_ref_worker, [(flavour, ch, rate, q, seconds, i, reps, abr) for i in range(cores)])
        wall = time.perf_counter() - t0
    return cores * reps * seconds / wall, wall, max(r[0] for r in res)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import reflib
    if not reflib.available("sse"):
        reflib.build()
    cores = os.cpu_count() or 1
    seconds = min(float(args.seconds), 60.0)      # a bounded sample of the workload: streams of the workload's kind, at most 60 s each
    vals = []
    for i in range(args.warmup + args.steps):
        v, wall, _ = reference_throughput(cores, seconds, 2, config=args.config)
        if i >= args.warmup:
            vals.append((v, wall))
    value = float(np.mean([v for v, _ in vals]))
    ms = float(np.mean([w for _, w in vals]) * 1e3)
    sample = ("each step: %d processes (one per host core) x two passes over one %g s stream of the workload "
              "(reference build oracle/_ref, Lancer SSE, -O3 -ffast-math, the flags configure.ac:189 ships)" % (cores, seconds))
    out = {"impl": "reference", "metric": "encoded audio-sec/sec @q4 44.1k stereo", "value": value, "unit": "audio-s/s",
           "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
           "config": workload_config(args.config, args.streams, args.seconds),
           "cpu_baseline": {"value": value, "unit": "audio-s/s", "cores": cores, "kind": "reference", "sample": sample},
           "e2e": {"value": value, "unit": "audio-s/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    if args.config != CONFIG:
        out["metric"] = "encoded audio-sec/sec @%s" % QUALITY_NAME.get(args.config, args.config)
    emit(out)


# ------------------------------------------------------------------ our arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from vorbis_aotuv_lancer_b200 import ops, encoder
    from vorbis_aotuv_lancer_b200.lib import L
    from vorbis_aotuv_lancer_b200 import blob as vblob

    ch, rate, q, _ = vblob.CATALOGUE[args.config]
    S, T = args.streams, args.seconds
    setup = ops.Setup(args.config, device=local)
    if args.fast_transform:
        setup.set_exact_transform(False)     # K1: own MDCT/FFT formulation (within 1e-5) instead of the default K1x (bit-identical)
    pins = [encoder.PinnedArray(ch, int(rate * T)) for _ in range(S)]      # the caller's PCM lives in page-locked memory
    pcms = synth_streams(L, S, T, ch, rate, rank * S, pinned=pins)
    blocks_per_stream = int(T * rate / 700) + 64
    if args.chunks <= 0:
        args.chunks = max(1, min(20, S // 3))        # 64 streams: 20 groups (sweep in encoder.group_weights); few streams: few groups
    enc = encoder.Encoder(setup, S * blocks_per_stream, threads=max(2, min(32, (os.cpu_count() or 8) // max(1, world))), chunks=args.chunks, nstreams=S)
    blocks, timeline, counts = enc.segment(pcms)
    # the resident-input measurement runs the whole workload as one batch per step; two batch objects alternate, so
    # the host half of step i+1 (descriptor pre-pass, launches) overlaps the device half of step i
    fulls = [ops.Batch(setup, len(blocks)) for _ in range(2 if args.depth > 1 else 1)]
    full = fulls[0]
    if len(fulls) > 1:
        L.vb200_batch_after(fulls[0].h, fulls[1].h)
        L.vb200_batch_after(fulls[1].h, fulls[0].h)
    nblocks = len(blocks)
    nlong = int(np.frombuffer(blocks, dtype=encoder.BLOCK_DTYPE)["W"].sum())
    nshort = nblocks - nlong
    d_timeline = torch.from_numpy(timeline).cuda()
    audio = S * T
    total_steps = args.warmup + args.steps
    # one fresh set of streams per step and arm (a step then always starts from the reference's initial state); a setup
    # holds 4096 stream slots, so very long runs reuse sets: their streams then simply go on where they stopped
    n_sets = max(4, min(2 * total_steps, 3500 // max(1, S)))
    _sets = [[ops.Stream(setup) for _ in range(S)] for _ in range(n_sets)]

    class _Sets:
        def __getitem__(self, i):
            return _sets[i % n_sets]
    streams = _Sets()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def maxtime(t):
        if world > 1:
            x = torch.tensor([t], device="cuda", dtype=torch.float64)
            dist.all_reduce(x, op=dist.ReduceOp.MAX)
            return float(x.item())
        return t

    # ---- device-resident value ----
    def resident(first, count, record):
        inflight = []
        for i in range(count):
            b = fulls[i % len(fulls)]
            if len(inflight) == len(fulls):
                done = inflight.pop(0)
                done.collect()
                if record:
                    for name, ms in done.kernel_times():
                        kt[name] = kt.get(name, 0.0) + ms
            b.submit(streams[first + i], blocks, d_timeline.data_ptr(), timeline.size)
            inflight.append(b)
        for done in inflight:
            done.collect()
            if record:
                for name, ms in done.kernel_times():
                    kt[name] = kt.get(name, 0.0) + ms

    kt = {}
    resident(0, args.warmup, False)
    stop, samples = threading.Event(), []
    th = threading.Thread(target=clocks_sampler, args=(stop, samples, local), daemon=True)
    barrier()
    L.vb200_launch_count_reset()
    th.start()
    # the K timed steps: device time between a CUDA event recorded on the first batch's (idle) stream before the first
    # submit and one recorded behind the last collect (the packets of the last step are on the host), max over ranks;
    # the host clock around the same region is reported next to it
    from vorbis_aotuv_lancer_b200.lib import check
    check(L.vb200_batch_mark(fulls[0].h), "vb200_batch_mark")          # its stream is idle: the region starts here
    t0 = time.perf_counter()
    resident(args.warmup, args.steps, True)                            # returns when every step has been collected
    if len(fulls) > 1:
        check(L.vb200_batch_mark(fulls[1].h), "vb200_batch_mark")      # all device work of the region is complete
    barrier()
    t_wall = maxtime(time.perf_counter() - t0)
    if len(fulls) > 1:
        ev_ms = C.c_float()
        check(L.vb200_batch_mark_elapsed(fulls[0].h, fulls[1].h, C.byref(ev_ms)), "vb200_batch_mark_elapsed")
        t_dev = maxtime(ev_ms.value * 1e-3)
    else:
        t_dev = t_wall          # a single batch object cannot hold both marks: the synchronised host clock stands in
    launches = L.vb200_launch_count()
    stop.set()
    th.join()
    tab = encoder.table(fulls[(args.steps - 1) % len(fulls)])
    packet_bytes = tab["total"]

    # ---- e2e: host PCM -> packets through the public call ----
    e2e_steps = max(1, min(args.steps, 3))
    t_e2e = None
    if not args.no_e2e:
        for i in range(min(args.warmup, 1)):
            enc.encode(pcms, streams[total_steps + i])
        barrier()
        t0 = time.perf_counter()
        for i in range(e2e_steps):
            enc.encode(pcms, streams[total_steps + 1 + i])
        barrier()
        t_e2e = maxtime(time.perf_counter() - t0)

    # ---- the same call fed with 16-bit PCM (vb200_segmenter_begin_s16): half the host-to-device bytes, the device
    # converts.  An additional number, not the headline: the reference's own API takes float
    t_e2e16 = None
    if not args.no_e2e and not args.no_int16:
        pins16 = [encoder.PinnedArray(ch, int(rate * T), np.int16) for _ in range(S)]
        for a16, a in zip(pins16, pcms):
            a16.array[:] = np.clip(np.rint(a * 32767.0), -32768, 32767).astype(np.int16)
        pcms16 = [a16.array for a16 in pins16]
        enc.encode(pcms16, streams[total_steps + 4])
        barrier()
        t0 = time.perf_counter()
        for i in range(e2e_steps):
            enc.encode(pcms16, streams[total_steps + 5 + i])
        barrier()
        t_e2e16 = maxtime(time.perf_counter() - t0)
        for a16 in pins16:
            a16.close()

    if rank == 0:
        ms_step = t_dev / args.steps * 1e3
        value = world * audio * args.steps / t_dev
        e2e_value = world * audio * e2e_steps / t_e2e if t_e2e else None
        # roofline of the dominant kernel
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak, peak_src = (peaks.get("hbm_gbs"), "measured") if peaks.get("hbm_gbs") else (6650.0, "fallback")
        stage_ms = {k: v / args.steps for k, v in kt.items()}
        kern = {k: v for k, v in stage_ms.items() if k in ALGO_BYTES}
        top = max(kern, key=kern.get)
        algo = (nlong * ALGO_BYTES[top][0] + nshort * ALGO_BYTES[top][1]) * ch
        achieved = algo / (kern[top] * 1e-3) / 1e9
        whole = sum((nlong * v[0] + nshort * v[1]) * ch for v in ALGO_BYTES.values())
        whole_gbs = whole / (sum(kern.values()) * 1e-3) / 1e9
        # DRAM traffic of the dominant kernel: from the committed ncu capture of this very workload, else null
        traffic = None
        for name in ("r02_traffic.json", "r01_traffic.json"):
            try:
                tr = json.load(open(os.path.join(ROOT, "profiles", name)))
                w = tr["workload"]
                if w["streams_per_gpu"] == S and w["seconds_per_stream"] == T and w.get("config", CONFIG) == args.config and top in tr["bytes_per_launch"]:
                    traffic = tr["bytes_per_launch"][top]
                    break
            except Exception:
                pass
        cores = os.cpu_count() or 1
        cpu = None
        if not args.no_cpu_baseline and world == 1:      # reported at N=1 only
            try:
                from oracle import reflib
                if reflib.available("sse"):
                    secs = min(float(T), 60.0)
                    v, wall, _ = reference_throughput(cores, secs, 2, config=args.config)
                    cpu = {"value": v, "unit": "audio-s/s", "cores": cores, "kind": "reference",
                           "sample": "%d processes (one per host core) x two passes over one %g s stream of the workload each, reference build "
                                     "oracle/_ref (Lancer SSE, -O3 -ffast-math), %.1f s wall" % (cores, secs, wall)}
            except Exception as e:  # the baseline is reported, never required
                cpu = {"value": None, "unit": "audio-s/s", "cores": cores, "kind": "reference", "sample": "failed: %r" % (e,)}
        out = {"metric": "encoded audio-sec/sec @q4 44.1k stereo", "value": value, "unit": "audio-s/s", "n_gpus": world,
               "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
               "vs_baseline": None, "dtype": "f32", "data": "synthetic",
               "config": workload_config(args.config, S, T),
               "timing": {"ms_per_step": "CUDA events on the batch streams around the K steps (inputs resident in HBM, packets back on the host), max over ranks",
                          "host_clock_ms_per_step": t_wall / args.steps * 1e3},
               "detail": {"blocks_per_step": nblocks, "long_blocks": nlong, "short_blocks": nshort, "resident_depth": len(fulls),
                          "front_end": "K1 (own radix-4 formulation, MDCT within 1e-5 of the reference)" if args.fast_transform
                                       else "K1x (the reference's own float operations: packets byte-identical to the reference's)",
                          "resident_bytes": {"pcm_timeline": int(timeline.nbytes), "per_stage_vector": int((nlong * 1024 + nshort * 128) * ch * 4)}},
               "e2e": {"value": e2e_value, "unit": "audio-s/s", "h2d_bytes_per_step": int(timeline.nbytes), "d2h_bytes_per_step": int(packet_bytes),
                       "steps": e2e_steps, "includes": "Encoder.encode from pinned host PCM: host segmentation (LPC head/tail, window sequencing; detector kernels on the GPU), "
                                                       "H2D of the PCM, all kernels, bit packing, D2H of the packets; %d pipelined stream groups; host clock, max over ranks" % args.chunks},
               "e2e_int16": None if t_e2e16 is None else {
                   "value": world * audio * e2e_steps / t_e2e16, "unit": "audio-s/s", "h2d_bytes_per_step": int(timeline.nbytes // 2),
                   "note": "the same Encoder.encode call fed with 16-bit PCM (vb200_segmenter_begin_s16: samples go up as int16, the device converts "
                           "sample / 32768.f as examples/encoder_example.c:183-190 does on the host); packets identical to the float entry point on the "
                           "converted samples (tests/test_gpu_encoder.py); an additional number -- the headline e2e is the float API's"},
               "gpu_launches": int(launches),
               "roofline": {"bound": "hbm", "kernel": top, "achieved": achieved, "peak": peak, "peak_source": peak_src, "unit": "GB/s",
                            "frac": achieved / peak, "traffic": traffic, "algorithmic_bytes_per_launch": int(algo), "kernel_ms": kern[top],
                            "all_kernels": {"algorithmic_GBps": whole_gbs, "frac": whole_gbs / peak, "ms": sum(kern.values()),
                                            "per_kernel_frac": {k: ((nlong * ALGO_BYTES[k][0] + nshort * ALGO_BYTES[k][1]) * ch / (kern[k] * 1e-3) / 1e9 / peak) for k in kern}}},
               "stage_ms": stage_ms, "clocks": summarize_clocks(samples), "cpu_baseline": cpu,
               "packet_bytes_per_step": int(packet_bytes), "kbps": packet_bytes * 8 / audio / 1000.0}
        if args.config != CONFIG:
            out["metric"] = "encoded audio-sec/sec @%s" % QUALITY_NAME.get(args.config, args.config)
        emit(out)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--streams", type=int, default=64)
    ap.add_argument("--seconds", type=float, default=60.0)
    ap.add_argument("--chunks", type=int, default=0, help="stream groups the e2e call pipelines (host stage of group g+1 overlaps kernels of g); 0: one per three streams, at most 20")
    ap.add_argument("--depth", type=int, default=2, help="resident arm: batches in flight (2: host half of step i+1 overlaps device half of step i)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-int16", action="store_true", help="skip the additional e2e measurement with 16-bit PCM input")
    ap.add_argument("--no-e2e", action="store_true", help="profiling runs only: skip the e2e arm (the line then carries e2e.value null)")
    ap.add_argument("--fast-transform", action="store_true", help="front end K1 (own formulation, within 1e-5) instead of the default K1x (DESIGN.md section 4)")
    ap.add_argument("--exact-transform", action="store_true", help="(the default since round 2; accepted for old command lines)")
    ap.add_argument("--config", default=CONFIG, help="setup blob: the metric's q4_44k_2ch, or another BASELINE configuration (q5_44k_2ch, q8_48k_6ch, q10_48k_2ch, abr128_44k_2ch ...)")
    args = ap.parse_args()
    # stdout carries the one JSON line and nothing else: libraries that write to fd 1 (NCCL's version banner)
    # are sent to stderr while the run lasts
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
