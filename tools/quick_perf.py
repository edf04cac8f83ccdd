"""Quick front-end / back-end timing probe on synthetic noise (not the bench; used while iterating)."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tools"))
import numpy as np, torch
import tfrec_b200 as tb
if os.environ.get("TFR_LIB"):   # experiments: a library built with other compile-time knobs
    tb.LIB_PATH = os.environ["TFR_LIB"]

def main():
    n_streams = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    mib = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    thresh = int(sys.argv[3]) if len(sys.argv) > 3 else 500
    types = int(sys.argv[4], 16) if len(sys.argv) > 4 else 7
    sigma = float(sys.argv[5]) if len(sys.argv) > 5 else 4.0
    nbytes = mib << 20
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    bufs = []
    if os.environ.get("TFR_QP_BENCH"):   # the bench workload (telegram every 10 s) instead of noise only
        import bench
        bufs = [bench.make_stream_gpu(s, nbytes, sigma, torch.device("cuda", 0))[0] for s in range(n_streams)]
    else:
        for s in range(n_streams):
            x = torch.randn(nbytes, device="cuda", generator=g) * sigma + 128.0
            bufs.append(x.round_().clamp_(0, 255).to(torch.uint8))
            del x
    rx = tb.Receiver(types=types, thresh=thresh, n_streams=n_streams)
    for it in range(4):
        for s in range(n_streams):
            rx.submit(s, bufs[s].data_ptr(), nbytes=nbytes)
        t0 = time.time(); rx.process(); rx.sync(); dt = time.time() - t0
        st = rx.stats()
        samples = n_streams * nbytes / 2
        print("iter %d wall %.3f ms  frontend %.3f ms (%.1f GS/s, %.1f GB/s)  backend %.3f ms  active %.2f%% frames %d thresh %d" % (
            it, dt * 1e3, st["last_frontend_ms"], samples / st["last_frontend_ms"] / 1e6, 2 * samples / st["last_frontend_ms"] / 1e6,
            st["last_backend_ms"], 100.0 * st["active_samples"] / (st["raw_samples"] / 4), rx.n_records(), rx.thresh(0)), "windows", rx.stats()["windows"], "reruns", rx.stats()["reruns"], "sr/bq/edge", rx.stats()["reruns_sr"], rx.stats()["reruns_biquad"], rx.stats()["reruns_edge"],
              "screen/dense/cand/true", st["screen_blocks"], st["dense_blocks"], st["screen_candidates"], st["screen_triggers"])
        rx.records(); rx.clear()
    # pipelined: K calls in flight, one sync (front-end of call i+1 overlaps the back-end of call i)
    K = 8
    for rep in range(2):
        t0 = time.time()
        for it in range(K):
            for s in range(n_streams):
                rx.submit(s, bufs[s].data_ptr(), nbytes=nbytes)
            rx.process()
        rx.sync(); dt = time.time() - t0
        st = rx.stats()
        samples = K * n_streams * nbytes / 2
        print("pipelined x%d: wall %.3f ms/call, device span %.3f ms/call (%.1f GS/s), front-end %.3f ms/call" % (
            K, dt * 1e3 / K, st["last_total_ms"] / K, samples / st["last_total_ms"] / 1e6, st["last_frontend_ms"] / K), "records", rx.n_records())
        rx.clear()

if __name__ == "__main__":
    main()
