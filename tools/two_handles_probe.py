"""Experiment: does more back-end concurrency help?  The bench workload (32 streams x 128 MiB per step) through ONE handle
(two calls in flight) against TWO handles of 16 streams each whose calls interleave (four half-size calls in flight)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
import tfrec_b200 as tb

def main():
    S, nbytes, K = 32, 128 << 20, 8
    dev = torch.device("cuda", 0)
    bufs = [bench.make_stream_gpu(s, nbytes, 4.0, dev)[0] for s in range(S)]
    def run(handles):
        per = S // len(handles)
        rxs = [tb.Receiver(types=7, thresh=0, n_streams=per, max_blocks_per_submit=nbytes // 65536) for _ in handles]
        def step():
            for h, rx in enumerate(rxs):
                for s in range(per):
                    rx.submit(s, bufs[h * per + s].data_ptr(), nbytes=nbytes)
                rx.process()
        for _ in range(3):
            step()
            for rx in rxs: rx.sync(); rx.clear()
        best = 1e9
        for rep in range(3):
            torch.cuda.synchronize(); t0 = time.perf_counter()
            for _ in range(K): step()
            for rx in rxs: rx.sync()
            torch.cuda.synchronize(); dt = time.perf_counter() - t0
            best = min(best, dt)
            n = sum(rx.n_records() for rx in rxs)
            for rx in rxs: rx.clear()
        for rx in rxs: rx.close()
        return best / K * 1e3, n
    for nh in (1, 2, 4):
        ms, n = run(range(nh))
        print("%d handle(s) x %d streams: %.3f ms per 4 GiB step (%.1f GS/s), records %d" % (nh, S // nh, ms, S * nbytes / 2 / ms / 1e6, n))

if __name__ == "__main__":
    main()
