"""BASELINE.json configs[4]: IQ buffer sweep 2^20 .. 2^30 raw samples through the front-end kernel (decimate /4 +
trigger), HBM GB/s against the measured peak.  One stream, input resident in HBM, fixed threshold above the noise
(no windows: the front-end alone), and the same buffer through the whole `-T 7` auto-threshold path.

The reference's decimator supports /2^p for any p >= 1 (dsp_stuff.cpp:232-264); only /4 feeds the decoders (spb
assumes 384 kS/s, main.cpp:186) and runs through the fused front-end.  `sweep.py decim` sweeps the other
decimations, /2 .. /32, through the stand-alone cascade tfr_downconvert (csrc/decim.cu, one launch per stage)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import tfrec_b200 as tb
if os.environ.get("TFR_LIB"):   # experiments: a library built with other compile-time knobs
    tb.LIB_PATH = os.environ["TFR_LIB"]


def decim_sweep(peak):
    """decimation sweep: passes 1..5 over 2^24 .. 2^30 raw samples resident in HBM; GB/s = 2 B per raw sample / time"""
    g = torch.Generator(device="cuda"); g.manual_seed(6)
    print("# tfr_downconvert (stand-alone cascade, one launch per stage), narrow filter; ms per cascade and algorithmic GB/s (2 B per raw sample)")
    print("%12s %8s | %s" % ("raw samples", "MiB", " | ".join("/%-2d %8s %7s %6s" % (1 << p, "ms", "GB/s", "%peak") for p in range(1, 6))))
    for lg in (24, 26, 28, 30):
        n = 1 << lg
        nbytes = 2 * n
        buf = torch.randint(0, 256, (nbytes,), device="cuda", dtype=torch.uint8, generator=g)
        out = torch.empty(n, device="cuda", dtype=torch.int16)      # large enough for /2
        cells = []
        for p in range(1, 6):
            best = 1e9
            for it in range(3):
                cnt, ms = tb.downconvert_device(buf.data_ptr(), nbytes, out.data_ptr(), passes=p, filter=0, reps=5)
                best = min(best, ms)
            cells.append("/%-2d %8.4f %7.1f %5.1f%%" % (1 << p, best, nbytes / best / 1e6, 100 * nbytes / best / 1e6 / peak))
        print("%12d %8.1f | %s" % (n, nbytes / 2**20, " | ".join(cells)))
        del buf, out


def synth_bytes(lo, hi, dev):
    """deterministic byte stream by index (every rank can make any slice of the same stream): a hashed counter"""
    idx = torch.arange(lo, hi, device=dev, dtype=torch.int64)
    return (((idx * 2654435761) >> 13) ^ (idx >> 7)).bitwise_and_(0xff).to(torch.uint8)


def mgpu_sweep(peak):
    """BASELINE configs[4] across GPUs: ONE IQ buffer of 2^20 .. 2^30 raw samples, decimation /2 .. /32, time-sharded over
    the ranks of a torchrun launch.  A FIR has finite support (410 raw samples for /32), so rank r decimates its slice
    plus a lead-in of 64 outputs' worth of raw samples and drops those 64 outputs: what remains is bit-identical to the
    single-GPU result (checked here against rank 0 decimating the whole buffer, up to 2^26 samples).  No collective on
    the data path; time = max over ranks of the CUDA-event time of the cascade, GB/s = 2 B per raw sample of the WHOLE
    buffer / that time.   torchrun --nproc-per-node N tools/sweep.py mgpu"""
    import hashlib
    import torch.distributed as dist
    rank, world, local = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if rank == 0:
        print("# time-sharded tfr_downconvert over %d GPU(s), narrow filter; aggregate algorithmic GB/s (2 B per raw sample) and %% of N x the measured HBM peak" % world)
        print("%12s %8s | %s" % ("raw samples", "MiB", " | ".join("/%-2d %8s %8s %6s" % (1 << p, "ms", "GB/s", "%peak") for p in range(1, 6))))
    for lg in (20, 22, 24, 26, 28, 30):
        n = 1 << lg                                       # raw IQ samples in the whole buffer
        cells = []
        for p in range(1, 6):
            lead = 64 << p                                # raw samples of lead-in (64 outputs)
            lo, hi = rank * n // world, (rank + 1) * n // world
            lo_in = lo - lead if rank else lo
            buf = synth_bytes(2 * lo_in, 2 * hi, dev)
            out = torch.empty(2 * ((hi - lo_in) >> p) + 16, device=dev, dtype=torch.int16)
            best = 1e9
            for it in range(3):
                if world > 1:
                    dist.barrier()
                cnt, ms = tb.downconvert_device(buf.data_ptr(), buf.numel(), out.data_ptr(), passes=p, filter=0, device=local, reps=5)
                best = min(best, ms)
            mine = out[: cnt][(2 * 64 if rank else 0):]    # drop the lead-in's outputs
            t = torch.tensor([best], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ok = ""
            if n <= (1 << 26):                            # parity: the slices, concatenated, are the single-GPU result
                parts = [mine.cpu()]
                if world > 1:
                    gathered = [None] * world
                    dist.all_gather_object(gathered, mine.cpu().numpy().tobytes())
                    parts = gathered
                if rank == 0:
                    whole_in = synth_bytes(0, 2 * n, dev)
                    whole = torch.empty(2 * (n >> p) + 16, device=dev, dtype=torch.int16)
                    c2, _ = tb.downconvert_device(whole_in.data_ptr(), whole_in.numel(), whole.data_ptr(), passes=p, filter=0, device=local)
                    want = hashlib.sha256(whole[:c2].cpu().numpy().tobytes()).hexdigest()
                    got = hashlib.sha256(b"".join(parts) if world > 1 else parts[0].numpy().tobytes()).hexdigest()
                    ok = "=" if got == want else "MISMATCH"
                    assert got == want, "time-sharded result differs from the single-GPU result (n=2^%d, /%d)" % (lg, 1 << p)
                    del whole_in, whole
            tt = float(t.item())
            cells.append("/%-2d %8.4f %8.1f %5.1f%%%s" % (1 << p, tt, 2.0 * n / tt / 1e6, 100 * 2.0 * n / tt / 1e6 / (peak * world), ok))
            del buf, out
        if rank == 0:
            print("%12d %8.1f | %s" % (n, 2.0 * n / 2**20, " | ".join(cells)))
    if world > 1:
        dist.destroy_process_group()


def main():
    peak = 6544.7
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    if len(sys.argv) > 1 and sys.argv[1] == "decim":
        return decim_sweep(peak)
    if len(sys.argv) > 1 and sys.argv[1] == "mgpu":
        return mgpu_sweep(peak)
    g = torch.Generator(device="cuda"); g.manual_seed(5)
    print("%12s %10s %12s %10s %8s | %12s %10s" % ("raw samples", "MiB", "front ms", "GB/s", "%peak", "T7 auto ms", "GS/s"))
    for lg in range(20, 31):
        n = 1 << lg
        nbytes = 2 * n
        buf = (torch.randn(nbytes, device="cuda", generator=g) * 4.0 + 128.0).round_().clamp_(0, 255).to(torch.uint8)
        res = []
        for types, thresh in ((0x07, 20000), (0x07, 0)):
            rx = tb.Receiver(types=types, thresh=thresh, n_streams=1, max_blocks_per_submit=nbytes // 65536)
            best_fe, best_tot = 1e9, 1e9
            for it in range(6):
                rx.submit(0, buf.data_ptr(), nbytes=nbytes)
                rx.process(); rx.sync()
                st = rx.stats()
                if it >= 2:
                    best_fe = min(best_fe, st["last_frontend_ms"]); best_tot = min(best_tot, st["last_total_ms"])
                rx.clear()
            rx.close()
            res.append((best_fe, best_tot))
        fe = res[0][0]
        print("%12d %10.1f %12.4f %10.1f %7.1f%% | %12.4f %10.1f" % (n, nbytes / 2**20, fe, nbytes / fe / 1e6, 100 * nbytes / fe / 1e6 / peak,
                                                                     res[1][1], n / res[1][1] / 1e6))
        del buf


if __name__ == "__main__":
    main()
