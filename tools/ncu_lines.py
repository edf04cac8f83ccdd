"""Per-source-line instruction shares of one kernel: joins the SASS page of an ncu report (--set full --import-source on)
with the line info nvdisasm prints for the cubin of the very library that ran.

usage: python tools/ncu_lines.py report.ncu-rep library.so KERNEL_REGEX MANGLED_NAME [top N]
e.g.   python tools/ncu_lines.py gpurun_out/x.ncu-rep tfrec_b200/libtfrb200.so '^win_kernel' _ZN3tfr10win_kernelENS_10BackParamsE 45"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile


def main():
    rep, lib, kregex, mangled = sys.argv[1:5]
    top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
    with tempfile.TemporaryDirectory() as td:
        subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=td, check=True, capture_output=True)
        dis = None
        for f in sorted(os.listdir(td)):
            out = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(td, f)], capture_output=True, text=True).stdout
            if (".text." + mangled + ":") in out:
                dis = out
                break
    assert dis, "kernel not found in the library's cubins"
    lines = dis.split("\n")
    start = [i for i, l in enumerate(lines) if l.startswith(".text." + mangled + ":")][0]
    cur, off2line = None, {}
    for l in lines[start + 1:]:
        if (l.startswith("//---") or (l.startswith(".section") and "text" in l)) and off2line:
            break
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (m.group(1), int(m.group(2)))
            continue
        m = re.search(r"/\*([0-9a-f]{4,})\*/\s+(\S.*?);", l)
        if m and cur:
            off2line[int(m.group(1), 16)] = cur
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kregex, "--print-source", "sass"],
                         capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[1]
    ia, it, isamp = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
    data = [r for r in rows[2:] if len(r) > it and r[ia].isdigit()]
    base = int(data[0][0], 16)
    tot = sum(int(r[ia]) for r in data)
    ttot = sum(int(r[it]) for r in data)
    stot = sum(int(r[isamp]) for r in data if r[isamp].isdigit())
    print("# %d warp instructions, %.1f active lanes on average, %d stall samples" % (tot, ttot / max(tot, 1), stot))
    by, byt, bys = collections.Counter(), collections.Counter(), collections.Counter()
    for r in data:
        k = off2line.get(int(r[0], 16) - base, ("?", 0))
        by[k] += int(r[ia])
        byt[k] += int(r[it])
        bys[k] += int(r[isamp]) if r[isamp].isdigit() else 0
    cache = {}
    for (f, ln), c in by.most_common(top):
        if f not in cache:
            try:
                cache[f] = open(f).read().split("\n")
            except OSError:
                cache[f] = []
        text = cache[f][ln - 1].strip()[:100] if 0 < ln <= len(cache[f]) else ""
        print("%5.1f%% instr  %5.1f%% time  lanes %4.1f  %s:%d  %s" % (100 * c / tot, 100 * bys[(f, ln)] / max(stot, 1), byt[(f, ln)] / c,
                                                                      os.path.basename(f), ln, text))


if __name__ == "__main__":
    main()
