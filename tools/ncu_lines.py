"""Correlates an ncu SASS source page (ncu -i X.ncu-rep --page source --csv) with nvdisasm -g line info and prints the
executed-instruction and stall-sample share per source line / per file.  Usage:
    python tools/ncu_lines.py <report.ncu-rep> <kernel mangled-name substring> [top_n]
"""
import csv, io, os, re, subprocess, sys, collections

rep, kname = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
so = os.environ.get("GSR_PROF_LIB") or os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "genesis_b200", "csrc", "libgsrigid.so")  # GSR_PROF_LIB: the copy that ran (a model-specialised one)
tmp = "/tmp/ncu_lines"
os.makedirs(tmp, exist_ok=True)
subprocess.run(f"cd {tmp} && rm -f *.cubin && cuobjdump -xelf all {so} > /dev/null && nvdisasm -g *.cubin > all.sass 2>/dev/null", shell=True, check=True)
off2line, cur, inside = {}, None, False
for line in open(f"{tmp}/all.sass"):
    if line.startswith("\t.section\t.text."):
        inside = kname in line
        continue
    if not inside:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', line)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(\S.*);", line)
    if m:
        off2line[int(m.group(1), 16)] = (cur, m.group(2).strip())
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)
hdr = rows[hi]
ca, ci, cs = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
ct = hdr.index("Thread Instructions Executed") if "Thread Instructions Executed" in hdr else None
t_line = collections.Counter()
base = None
by_line, by_file = collections.Counter(), collections.Counter()
s_line, s_file = collections.Counter(), collections.Counter()
tot = stot = 0
for r in rows[hi + 1:]:
    if len(r) < len(hdr):
        continue
    if r[ca] == "Address":  # a second captured launch follows: one launch is enough
        break
    a = int(r[ca], 16)
    base = a if base is None else base
    n, s = int(r[ci] or 0), int(r[cs] or 0)
    loc = off2line.get(a - base, (None, ""))[0] or ("?", 0)
    by_line[loc] += n; by_file[loc[0]] += n; s_line[loc] += s; s_file[loc[0]] += s
    if ct is not None:
        t_line[loc] += int(r[ct] or 0)
    tot += n; stot += s
print(f"total warp instructions {tot}, stall samples {stot}")
for f, n in by_file.most_common():
    print(f"  {f:24s} inst {100 * n / tot:5.1f}%  samples {100 * s_file[f] / max(stot, 1):5.1f}%")
srcs = {}
print("top lines by samples (where warps wait):")
for loc, s in s_line.most_common(top):
    f, l = loc
    if f not in srcs:
        p = os.path.join(os.path.dirname(so), f)
        srcs[f] = open(p).read().split("\n") if os.path.exists(p) else []
    text = srcs[f][l - 1].strip()[:110] if 0 < l <= len(srcs[f]) else ""
    print(f"  {f}:{l:<4d} samples {100 * s / max(stot, 1):5.1f}%  inst {100 * by_line[loc] / tot:5.1f}%  | {text}")

flt = os.environ.get("GSR_LINES_FILTER")  # e.g. gsr_collide.cuh: the hottest lines of one file, with the lanes active on them
if flt:
    print(f"top lines of {flt} by samples (active lanes = thread instructions / warp instructions):")
    rows_f = [(loc, s) for loc, s in s_line.most_common() if loc[0] == flt][:top]
    for loc, s in rows_f:
        f, l = loc
        text = srcs.get(f, [])
        if f not in srcs:
            pth = os.path.join(os.path.dirname(so), f)
            srcs[f] = open(pth).read().split("\n") if os.path.exists(pth) else []
        text = srcs[f][l - 1].strip()[:100] if 0 < l <= len(srcs[f]) else ""
        lanes = t_line[loc] / by_line[loc] if by_line[loc] else 0.0
        print(f"  {f}:{l:<4d} samples {100 * s / max(stot, 1):5.2f}%  inst {100 * by_line[loc] / tot:5.2f}%  lanes {lanes:4.1f} | {text}")
