import sys, os
sys.path.insert(0, "/root/repo")
import torch, tfrec_b200 as tb
if os.environ.get("TFR_LIB"): tb.LIB_PATH = os.environ["TFR_LIB"]
n = 1 << 30
g = torch.Generator(device="cuda"); g.manual_seed(5)
buf = (torch.randn(2 * n, device="cuda", generator=g) * 4.0 + 128.0).round_().clamp_(0, 255).to(torch.uint8)
rx = tb.Receiver(types=7, thresh=0, n_streams=1, max_blocks_per_submit=2 * n // 65536)
for it in range(4):
    rx.submit(0, buf.data_ptr(), nbytes=2 * n)
    rx.process(); rx.sync()
    st = rx.stats()
    print("iter", it, "total %.3f ms fe %.3f be %.3f windows %d" % (st["last_total_ms"], st["last_frontend_ms"], st["last_backend_ms"], st["windows"]))
    rx.clear()
