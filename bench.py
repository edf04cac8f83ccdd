#!/usr/bin/env python
"""bench.py - IQ MSamples/s (+ telegrams/s) of the IQ->telegram decode path on N B200s.

  python bench.py --gpus N --steps K --warmup W            (N>1: launched by torchrun, one rank per GPU)
  python bench.py --impl reference ...                     the unmodified reference CPU path (oracle/_ref/tfrec)

Workload (BASELINE.json configs[1]): `-T 7` (TFA_1 + TFA_2 + TFA_3), default auto threshold, continuous
1.536 MS/s synthetic IQ per stream ("stick") with one telegram every 10 s (15.36 M samples), types rotated.
One step = one pass of the hot path over every stream's buffer.  Streams are independent, so they shard
across ranks with no data-path collective (weak scaling: per-GPU work is fixed).

`value`   whole-job throughput with the IQ bytes already resident in HBM (device time, lib CUDA events)
`e2e`     the same pass through the public C ABI with HOST (pinned) buffers: H2D of every step's input and
          D2H of its results inside the timed region
`roofline` the front-end kernel: 2 algorithmic bytes per raw IQ sample / its CUDA-event time, vs the measured
          HBM peak in MEASURED_PEAKS.json
`cpu_baseline` the unmodified reference binary on the box's host cores, one process per core, bounded sample
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
for p in (ROOT, os.path.join(ROOT, "tools")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

FS = 1536000
PERIOD = 10 * FS               # one telegram every 10 s
TYPES_MASK = 0x07
SENSORS = (0, 1, 2)            # TFA_1, TFA_2, TFA_3
REF_BIN = os.path.join(ROOT, "oracle", "_ref", "tfrec")


def bench_config(args):
    """the `config` object of the JSON line: the same for both arms (the driver compares them)"""
    nbytes = args.mib << 20
    return {"workload": workload_name(args), "streams_per_gpu": args.streams, "bytes_per_stream": nbytes,
            "l2": "inputs (%.1f GiB per GPU) are far larger than the 126 MB L2" % (args.streams * nbytes / 2**30),
            "types": "0x%02x" % int(args.types, 16), "thresh": args.thresh}


def workload_name(args):
    t = int(args.types, 16)
    mix = "T7 default mix" if t == 7 else "T%x (all five decoders)" % t if t == 0x2f else "T%x" % t
    return ("%s, auto threshold, %d streams x %d MiB u8 IQ per GPU, 1.536 MS/s, telegram every 10 s, "
            "noise sigma %.1f LSB" % (mix, args.streams, args.mib, args.sigma))


# ------------------------------------------------------------------------------------------------- data
def stream_bursts(stream_id, n_samples):
    """telegram schedule of one stream: (raw sample offset, sensor, frame bytes)"""
    import iqsynth as g
    rng = np.random.default_rng(1000 + stream_id)
    out, at, k = [], (stream_id * 60000) % PERIOD + 100000, stream_id
    while at + 400000 < n_samples:
        s = SENSORS[k % len(SENSORS)]
        out.append(g.Burst(at, s, g.random_frame(s, rng), {"amp": 100}))
        at += PERIOD
        k += 1
    return out


def make_stream_gpu(stream_id, nbytes, sigma, device):
    """u8 IQ on the GPU: rounded Gaussian noise (torch generator, seeded per stream) + integer-exact bursts"""
    import torch
    import iqsynth as g
    gen = torch.Generator(device=device)
    gen.manual_seed(77000 + stream_id)
    n_samples = nbytes // 2
    buf = torch.empty(nbytes, dtype=torch.uint8, device=device)
    chunk = 1 << 26
    bursts = stream_bursts(stream_id, n_samples)
    rendered = [(b.at, *g.render_burst(b)) for b in bursts]
    for off in range(0, nbytes, chunk):
        n = min(chunk, nbytes - off)
        x = torch.randn(n, device=device, generator=gen).mul_(sigma).round_().to(torch.int16)
        for at, bi, bq in rendered:
            lo, hi = 2 * at, 2 * (at + len(bi))
            a, b = max(lo, off), min(hi, off + n)
            if a >= b:
                continue
            iq = np.empty(2 * len(bi), dtype=np.int16)
            iq[0::2], iq[1::2] = bi, bq
            x[a - off:b - off] += torch.from_numpy(iq[a - lo:b - lo]).to(device)
        buf[off:off + n] = (x + 128).clamp_(0, 255).to(torch.uint8)
        del x
    return buf, len(bursts)


# ------------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "10"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.rows.append([c.strip() for c in ln.split(",")])

    def mark(self):
        return len(self.rows)

    def stop(self, lo=0, hi=None):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        rows = self.rows[lo:hi] if hi is not None and hi > lo else self.rows[lo:]
        if not rows:
            rows = self.rows
        sm = sorted(int(r[0]) for r in rows if r and r[0].isdigit())
        mx = max([int(r[1]) for r in rows if len(r) > 1 and r[1].isdigit()] or [0])
        reasons = []
        for name, col in (("hw_slowdown", 3), ("hw_thermal_slowdown", 4), ("sw_thermal_slowdown", 5), ("sw_power_cap", 6)):
            if any(len(r) > col and r[col].lower().startswith("active") for r in rows):
                reasons.append(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None, "reasons": reasons,
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------- reference arm
def n_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def write_sample_files(td, n_files, nbytes, sigma):
    """bounded CPU sample: stream k's first nbytes, regenerated on the host with the same schedule"""
    import iqsynth as g
    paths = []
    for k in range(n_files):
        iq = g.make_stream(nbytes // 2, [b for b in stream_bursts(k, nbytes // 2)], seed=77000 + k, sigma=sigma)
        p = os.path.join(td, "s%d.iq" % k)
        iq.tofile(p)
        paths.append(p)
    return paths


def run_reference_once(paths, exec_echo=False, types=TYPES_MASK):
    """one reference process per file/core, all in parallel; returns (wall seconds, decoded telegram lines), or with
    exec_echo the per-file lists of `-e /bin/echo` argument lines (decoder.cpp:72-94) instead of the line count"""
    t0 = time.perf_counter()
    procs = []
    for k, p in enumerate(paths):
        cmd = [REF_BIN, "-T", "%x" % types] + (["-q", "-e", "/bin/echo"] if exec_echo else []) + ["-L", p]
        if os.path.exists("/usr/bin/taskset"):
            cmd = ["taskset", "-c", str(sorted(os.sched_getaffinity(0))[k % n_cores()])] + cmd
        procs.append(subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.DEVNULL))
    outs = [pr.communicate()[0] for pr in procs]
    dt = time.perf_counter() - t0
    if exec_echo:
        import re
        pat = re.compile(r"([0-9a-f]+ [+-][0-9.]+ \S+ -?\d+ -?\d+ -?\d+ -?\d+) \d+")
        return dt, [[m.group(1) for m in (pat.fullmatch(ln.strip()) for ln in o.decode("latin1").splitlines()) if m]
                    for o in outs]
    lines = sum(sum(1 for ln in o.decode("latin1").splitlines() if ln.startswith(("TFA1 ID", "TFA2 ID", "TFA3 ID")))
                for o in outs)
    return dt, lines


def executed(lines):
    """the -e lines decoder::store_data lets through in mode 0 (decoder.cpp:46-65): a WeatherHub (13-digit id) repeat
    only when its sequence number differs from the last one stored for that id; everything else always"""
    seen, out = {}, []
    for ln in lines:
        f = ln.split()
        if len(f[0]) == 13:
            if f[0] in seen and seen[f[0]] == f[3]:
                continue
            seen[f[0]] = f[3]
        out.append(ln)
    return out


def parity_check(tb, bufs, types, thresh, device, n_check):
    """Parity gate: the first n_check streams of THIS rank's workload, whole buffers, decoded by a
    fresh handle through the C ABI, against what the unmodified reference binary hands to `-e /bin/echo` for the very
    same bytes (the oracle port stands in only where oracle/_ref was not built).  Returns (streams checked, checker,
    first mismatch or None)."""
    n_check = min(n_check, len(bufs))
    if n_check <= 0:
        return 0, "none", None
    nbytes = int(bufs[0].numel())
    rx = tb.Receiver(types=types, thresh=thresh, n_streams=n_check, device=device, max_blocks_per_submit=nbytes // 65536)
    for s in range(n_check):
        rx.submit(s, bufs[s].data_ptr(), nbytes=nbytes)
    rx.process()
    recs = rx.records()
    rx.close()
    got = [executed([r["exec"] for r in recs if r["stream"] == s]) for s in range(n_check)]
    if os.path.exists(REF_BIN):
        checker = "oracle/_ref/tfrec -q -e /bin/echo"
        with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as td:
            paths = []
            for s in range(n_check):
                p = os.path.join(td, "p%d.iq" % s)
                bufs[s].cpu().numpy().tofile(p)
                paths.append(p)
            _, want = run_reference_once(paths, exec_echo=True, types=types)
    else:
        checker = "oracle port (oracle/_ref not built)"
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle_lib as ol
        want = []
        for s in range(n_check):
            o = ol.Oracle(types=types, thresh=thresh)
            o.process(bufs[s].cpu().numpy())
            want.append(executed([r["exec"] for r in o.records()]))
            o.close()
    for s in range(n_check):
        if got[s] != want[s]:
            return n_check, checker, {"stream": s, "gpu": got[s][:8], "reference": want[s][:8],
                                      "n_gpu": len(got[s]), "n_reference": len(want[s])}
    return n_check, checker, None


def cpu_baseline(sigma, sample_mib=256, gpu_bufs=None, types=TYPES_MASK):
    """the unmodified reference on all host cores, one process per core, each on the first sample_mib MiB of one
    of THIS run's stream buffers (copied back from the GPU, so the bytes are the ones the GPU decoded)"""
    cores = n_cores()
    sample_mib = max(16, min(sample_mib, 8192 // cores))
    if not os.path.exists(REF_BIN):
        return {"value": None, "unit": "MSamples/s", "cores": cores, "kind": "reference",
                "sample": "oracle/_ref/tfrec missing (build it with `make -C oracle ref` where /root/reference exists)"}
    nbytes = sample_mib << 20
    with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as td:
        if gpu_bufs:
            paths = []
            nbytes = min(nbytes, int(gpu_bufs[0].numel()))
            for k in range(cores):
                p = os.path.join(td, "s%d.iq" % k)
                gpu_bufs[k % len(gpu_bufs)][:nbytes].cpu().numpy().tofile(p)
                paths.append(p)
        else:
            paths = write_sample_files(td, cores, nbytes, sigma)
        run_reference_once(paths[:1], types=types)       # page cache / binary warm-up
        dt, lines = run_reference_once(paths, types=types)
    sample_mib = nbytes >> 20
    samples = cores * (nbytes // 2)
    return {"value": round(samples / dt / 1e6, 2), "unit": "MSamples/s", "cores": cores, "kind": "reference",
            "telegrams_per_s": round(lines / dt, 3),
            "sample": "unmodified reference `tfrec -T %x -L` (auto threshold), %d processes x %d MiB (first %d MiB of "
                      "streams 0..%d of this workload), wall %.2f s" % (types, cores, sample_mib, sample_mib, cores - 1, dt)}


def main_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = n_cores()
    if not os.path.exists(REF_BIN):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/tfrec not built"}))
        return
    sample_mib = max(8, min(args.ref_mib, 8192 // cores))
    nbytes = sample_mib << 20
    with tempfile.TemporaryDirectory(dir="/dev/shm" if os.path.isdir("/dev/shm") else None) as td:
        paths = write_sample_files(td, cores, nbytes, args.sigma)
        rtypes = int(args.types, 16)
        for _ in range(args.warmup):
            run_reference_once(paths, types=rtypes)
        tot, lines = 0.0, 0
        for _ in range(args.steps):
            dt, ln = run_reference_once(paths, types=rtypes)
            tot += dt
            lines += ln
    samples = cores * (nbytes // 2) * args.steps
    v = samples / tot / 1e6
    out = {"impl": "reference", "metric": "iq_msamples_per_s", "value": round(v, 2), "unit": "MSamples/s",
           "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(1e3 * tot / args.steps, 3),
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "i32/f64", "data": "synthetic",
           "telegrams_per_s": round(lines / tot, 3),
           "config": bench_config(args), "reference_step": "bounded sample: %d processes x %d MiB" % (cores, sample_mib),
           "cpu_baseline": {"value": round(v, 2), "unit": "MSamples/s", "cores": cores, "kind": "reference",
                            "sample": "unmodified reference `tfrec -T %x -L`, %d processes x %d MiB per step" % (rtypes, cores, sample_mib)},
           "e2e": {"value": round(v, 2), "unit": "MSamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


# ------------------------------------------------------------------------------------------------- our arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--streams", type=int, default=32, help="streams per GPU")
    ap.add_argument("--mib", type=int, default=128, help="MiB of u8 IQ per stream per step")
    ap.add_argument("--sigma", type=float, default=4.0)
    ap.add_argument("--thresh", type=int, default=0, help="0 = auto (reference default)")
    ap.add_argument("--ref-mib", type=int, default=64, help="reference arm: MiB per process per step")
    ap.add_argument("--types", default="7", help="hex mask of registered decoders for the headline pass (7 = TFA_1/2/3, "
                    "2f = all five incl. TX22 and WeatherHub, BASELINE.json configs[2])")
    ap.add_argument("--parity-streams", type=int, default=-1, help="streams per rank checked against the reference "
                    "before timing (-1: all host cores' worth at N=1, 2 per rank under torchrun)")
    ap.add_argument("--no-extra", action="store_true", help="skip the all-five-decoders (-T 2f) extra leg")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return main_reference(args)

    import torch
    import tfrec_b200 as tb

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        sys.exit("bench.py needs a CUDA device: the decode path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(v):
        if dist is None:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(v):
        if dist is None:
            return v
        t = torch.tensor([v], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    nbytes = args.mib << 20
    S = args.streams
    bufs, sent = [], 0
    for s in range(S):
        b, n = make_stream_gpu(rank * S + s, nbytes, args.sigma, dev)
        bufs.append(b)
        sent += n
    samples_per_step = S * (nbytes // 2)

    types_mask = int(args.types, 16)

    # ---- parity gate: no number is printed for a workload the GPU decodes differently from the reference ----------
    n_par = args.parity_streams if args.parity_streams >= 0 else (min(n_cores(), S) if world == 1 else 2)
    par_n, par_checker, par_bad = parity_check(tb, bufs, types_mask, args.thresh, local, n_par)
    n_bad = allsum(1.0 if par_bad is not None else 0.0)
    par_total = int(allsum(float(par_n)))
    if n_bad:
        if par_bad is not None:
            sys.stderr.write("rank %d: PARITY MISMATCH against %s: %s\n" % (rank, par_checker, json.dumps(par_bad)))
        if rank == 0:
            print(json.dumps({"metric": "iq_msamples_per_s", "value": None, "unit": "MSamples/s", "n_gpus": world,
                              "parity_checked_streams": par_total, "parity": "MISMATCH on %d rank(s): no value reported" % int(n_bad),
                              "parity_checker": par_checker, "first_mismatch": par_bad}))
        if dist is not None:
            dist.destroy_process_group()
        sys.exit(1)

    rx = tb.Receiver(types=types_mask, thresh=args.thresh, n_streams=S, device=local, max_blocks_per_submit=nbytes // 65536)

    def step_device():
        for s in range(S):
            rx.submit(s, bufs[s].data_ptr(), nbytes=nbytes)
        rx.process()

    # ---- resident-in-HBM pass ------------------------------------------------------------------
    for _ in range(args.warmup):
        step_device()
        rx.sync()
        rx.clear()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        t_w = time.perf_counter()
        while sampler.mark() == 0 and time.perf_counter() - t_w < 3.0:   # nvidia-smi needs a moment to start
            step_device()
            rx.sync()
            rx.clear()
    barrier()
    m0 = sampler.mark()
    # ---- the timed region: EXACTLY K steps, a barrier + synchronise on both sides.  One step = submit every stream's
    # buffer + tfr_process.  The K calls are issued back to back - what a receiver that is fed continuously does - and
    # the library keeps two of them in flight on the device (two work-buffer slots: the front-end of step i+1 runs beside
    # the latency-bound back-end of step i).  Device time = the library's CUDA events, first front-end launch of the
    # first step to the end of the parsers of the last step.
    rx.stats()
    l0 = rx.stats()["kernel_launches"]
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_device()
    rx.sync()
    st = rx.stats()
    pipe_ms = st["last_total_ms"]
    fe_pipe_ms = st["last_frontend_ms"]
    decoded = rx.n_records()
    rx.clear()
    barrier()
    wall = time.perf_counter() - t0
    launches = rx.stats()["kernel_launches"] - l0
    # ---- the same K steps with a synchronisation after every step: the front-end's own duration is separable here
    # (CUDA-event span of the step's front-end launches, on the streams they run on), which is what the roofline of
    # the front-end kernel is computed from; reported beside the headline as `synchronised`
    dev_ms = fe_ms = be_ms = 0.0
    for _ in range(args.steps):
        step_device()
        rx.sync()
        st = rx.stats()
        dev_ms += st["last_total_ms"]
        fe_ms += st["last_frontend_ms"]
        be_ms += st["last_backend_ms"]
    st_end = rx.stats()
    windows = st_end["windows"]                       # (counted since the clear() above: the K synchronised steps)
    rx.clear()
    barrier()
    m1 = sampler.mark() + 1
    clocks = sampler.stop(m0, m1) if rank == 0 else None
    t_dev = allmax(dev_ms / 1e3)
    t_pipe = allmax(pipe_ms / 1e3)
    t_wall = allmax(wall)
    total_samples = allsum(float(samples_per_step)) * args.steps
    total_decoded = allsum(float(decoded))
    total_sent = allsum(float(sent)) * args.steps
    value = total_samples / t_pipe / 1e6

    # ---- end to end through the C ABI with host buffers ----------------------------------------
    e2e = None
    if not args.no_e2e:
        host = [b.cpu().pin_memory() for b in bufs]
        for _ in range(2):
            for s in range(S):
                rx.submit_host_ptr(s, host[s].data_ptr(), nbytes)
            rx.process()
            rx.records()
            rx.clear()
        barrier()
        t0 = time.perf_counter()
        d2h = 0
        # Every step's input is copied from pinned host memory (tfr_submit, TFR_MEM_HOST) and every step's frames and
        # records are read back, all inside the timed region.  The copies of step i+1 are queued as soon as
        # tfr_process(i) returns (its front-end, the only reader of the input arena, is done by then in auto-threshold
        # mode), so they run while the back-end of step i decodes - what a streaming caller of the ABI does.
        for s in range(S):
            rx.submit_host_ptr(s, host[s].data_ptr(), nbytes)
        for i in range(args.steps):
            rx.process()
            if i + 1 < args.steps:
                for s in range(S):
                    rx.submit_host_ptr(s, host[s].data_ptr(), nbytes)
            fr = rx.frames()
            rc = rx.records()
            d2h += 112 * len(fr) + 64 * len(rc) + 24
            rx.clear()
        barrier()
        t_e2e = allmax(time.perf_counter() - t0)
        e2e = {"value": round(total_samples / t_e2e / 1e6, 2), "unit": "MSamples/s",
               "h2d_bytes_per_step": int(S * nbytes), "d2h_bytes_per_step": int(d2h // max(args.steps, 1)),
               "ms_per_step": round(1e3 * t_e2e / args.steps, 3)}
        del host

    # ---- extra: all five decoders (BASELINE.json configs[2], `-T 2f`) on the same buffers ----------------------------
    extra = None
    if not args.no_extra and types_mask != 0x2f:
        rx.close()
        rx = None
        rx5 = tb.Receiver(types=0x2f, thresh=args.thresh, n_streams=S, device=local, max_blocks_per_submit=nbytes // 65536)
        ms5, k5 = 0.0, max(3, min(5, args.steps))
        for i in range(2 + k5):
            for s in range(S):
                rx5.submit(s, bufs[s].data_ptr(), nbytes=nbytes)
            rx5.process()
            rx5.sync()
            if i >= 2:
                ms5 += rx5.stats()["last_total_ms"]
            rx5.clear()
        rx5.close()
        t5 = allmax(ms5 / 1e3)
        extra = {"all_five_decoders_T2f": {"value": round(allsum(float(samples_per_step)) * k5 / t5 / 1e6, 2), "unit": "MSamples/s",
                                           "ms_per_step": round(1e3 * t5 / k5, 4), "steps": k5, "warmup": 2,
                                           "note": "same streams, TX22 and WeatherHub registered too (resident in HBM, synchronised per step)"}}

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    # ---- extra: ONE stick (BASELINE.json configs[4], single long stream): the first streams of the workload back to back ----
    if extra is not None:
        try:
            n_cat = max(1, min(S, (1 << 31) // nbytes))
            one = torch.cat([bufs[s] for s in range(n_cat)])
            rx1 = tb.Receiver(types=types_mask, thresh=args.thresh, n_streams=1, device=local, max_blocks_per_submit=int(one.numel()) // 65536)
            ms1, k1, rec1 = 0.0, 3, 0
            for i in range(2 + k1):
                rx1.submit(0, one.data_ptr(), nbytes=int(one.numel()))
                rx1.process()
                rx1.sync()
                if i >= 2:
                    ms1 += rx1.stats()["last_total_ms"]
                    rec1 = len(rx1.records())
                rx1.clear()
            rx1.close()
            extra["one_stream"] = {"value": round(int(one.numel()) // 2 * k1 / (ms1 / 1e3) / 1e6, 2), "unit": "MSamples/s",
                                   "ms_per_step": round(ms1 / k1, 4), "raw_samples": int(one.numel()) // 2, "records_per_step": rec1,
                                   "steps": k1, "warmup": 2,
                                   "note": "the first %d streams of the workload back to back as ONE stream on one handle (resident in HBM, "
                                           "synchronised per step): the per-stream serial chains - threshold walk, verifier - decide; not parity-gated itself "
                                           "(the gate above checks the decoders on the streams, tests/test_gpu_parity.py long single streams)" % n_cat}
            del one
        except Exception as e:   # an extra must not cost the headline line
            extra["one_stream"] = {"error": str(e)[:200]}

    # ---- roofline of the front-end kernel -------------------------------------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = 2.0 * samples_per_step * args.steps / (fe_ms / 1e3) / 1e9   # rank 0's kernel
    traffic = None
    screened = st_end.get("screen_blocks", 0) > 0
    tfile = "frontend_screen_traffic.json" if screened else "frontend_traffic.json"
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", tfile)))
        traffic = tj.get("dram_bytes_per_algorithmic_byte")
    except Exception:
        pass
    roofline = {"kernel": "frontend_screen_kernel<narrow>" if screened else "frontend_kernel<narrow>", "bound": "hbm",
                "achieved": round(achieved, 1), "peak": peak,
                "unit": "GB/s", "frac": round(achieved / peak, 4),
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s",
                "traffic": (None if traffic is None else round(traffic * 2.0 * samples_per_step, 0)),
                "traffic_source": "profiles/%s (ncu --set full dram bytes per algorithmic byte of this kernel) x "
                                  "this run's algorithmic bytes; not re-measured in this run" % tfile,
                "algorithmic_bytes_per_step": int(2 * samples_per_step),
                "frontend_ms_per_step": round(fe_ms / args.steps, 4),
                "other_ms_per_step": round(be_ms / args.steps, 4),
                "frontend_ms_per_step_in_flight": round(fe_pipe_ms / args.steps, 4),
                "screen": {k: int(st_end.get(k, 0)) for k in ("screen_blocks", "dense_blocks", "screen_candidates", "screen_triggers")},
                "note": "achieved = algorithmic bytes (2 B per raw IQ sample) / summed CUDA-event time of the front-end launches "
                        "of the K synchronised steps, when the kernel has the GPU to itself (with steps in flight it shares the "
                        "SMs with the previous step's back-end: frontend_ms_per_step_in_flight); see DESIGN.md 4.1"}

    out = {"metric": "iq_msamples_per_s", "value": round(value, 2), "unit": "MSamples/s", "n_gpus": world,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(1e3 * t_pipe / args.steps, 4),
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8->i16 (exact fp32 FMA.RM), f64 demod",
           "data": "synthetic",
           "config": bench_config(args),
           "parity_checked_streams": par_total, "parity_checker": par_checker,
           "telegrams_per_s": round(total_decoded / t_pipe, 2), "telegrams_decoded": int(total_decoded),
           "telegrams_sent": int(total_sent), "wall_ms_per_step": round(1e3 * t_wall / args.steps, 4),
           "gpu_launches": int(launches), "demod_windows_per_step": int(windows // max(args.steps, 1)),
           "timed_region": "K steps issued back to back, barrier + synchronise before and after (two steps in flight on the device)",
           "synchronised": {"value": round(total_samples / t_dev / 1e6, 2), "unit": "MSamples/s",
                            "ms_per_step": round(1e3 * t_dev / args.steps, 4),
                            "note": "the same K steps with a tfr_sync after every step (no overlap between steps)"},
           "clocks": clocks, "roofline": roofline}
    if e2e is not None:
        out["e2e"] = e2e
    if extra is not None:
        out["extra"] = extra
    if world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(args.sigma, gpu_bufs=bufs, types=types_mask)
    print(json.dumps(out))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
