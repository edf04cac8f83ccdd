"""Debug probe: runs a few bench-shaped steps (32 sticks x 128 MiB, -T 7, auto threshold) with TFR_DEBUG=1 so that the
library prints the timeline of each call (front-end done / threshold walk done / back-end start and end) and, with a
library built with -DTFR_WIN_PROFILE (TFR_LIB=...), the in-kernel phase clocks of the window kernel."""
import sys,os
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tools')
os.environ["TFR_DEBUG"]="1"
import bench, torch, tfrec_b200 as tb
if os.environ.get("TFR_LIB"): tb.LIB_PATH = os.environ["TFR_LIB"]
S=32; nbytes=128<<20
bufs=[bench.make_stream_gpu(s, nbytes, 4.0, torch.device("cuda",0))[0] for s in range(S)]
rx=tb.Receiver(types=7,thresh=0,n_streams=S,max_blocks_per_submit=nbytes//65536)
for it in range(6):
    for s in range(S): rx.submit(s,bufs[s].data_ptr(),nbytes=nbytes)
    rx.process(); rx.sync(); rx.clear()
