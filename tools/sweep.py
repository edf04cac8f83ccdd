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


def main():
    peak = 6544.7
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
    if len(sys.argv) > 1 and sys.argv[1] == "decim":
        return decim_sweep(peak)
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
