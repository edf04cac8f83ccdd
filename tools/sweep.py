"""BASELINE.json configs[4]: IQ buffer sweep 2^20 .. 2^30 raw samples through the front-end kernel (decimate /4 +
trigger), HBM GB/s against the measured peak.  One stream, input resident in HBM, fixed threshold above the noise
(no windows: the front-end alone), and the same buffer through the whole `-T 7` auto-threshold path.

The reference's decimator supports /2^p for any p >= 1 (dsp_stuff.cpp:232-264), but only /4 feeds the decoders
(spb assumes 384 kS/s, main.cpp:186), so /4 is the one decimation this library implements."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import tfrec_b200 as tb


def main():
    peak = 6544.7
    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        pass
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
