"""Per-function share of executed warp instructions and stall samples of one kernel: attributes each SASS instruction of
the ncu source page to the enclosing (non-inlined) device function label of the cubin (nvdisasm).
    python tools/ncu_funcs.py <report.ncu-rep> <kernel mangled-name substring>"""
import csv, io, os, re, subprocess, sys, collections

rep, kname = sys.argv[1], sys.argv[2]
so = os.environ.get("GSR_PROF_LIB") or os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "genesis_b200", "csrc", "libgsrigid.so")  # GSR_PROF_LIB: the copy that ran (a model-specialised one)
tmp = "/tmp/ncu_lines"
os.makedirs(tmp, exist_ok=True)
subprocess.run(f"cd {tmp} && rm -f *.cubin && cuobjdump -xelf all {so} > /dev/null && nvdisasm -g *.cubin > all.sass 2>/dev/null", shell=True, check=True)
off2fn, cur, inside = {}, "kernel body", False
for line in open(f"{tmp}/all.sass"):
    if line.startswith("\t.section\t.text."):
        inside = kname in line
        cur = "kernel body"
        continue
    if not inside:
        continue
    m = re.match(r"^\$[^$]+\$(\S+):", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(\S.*);", line)
    if m:
        off2fn[int(m.group(1), 16)] = cur
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if "Instructions Executed" in r)
hdr = rows[hi]
ca, ci, cs = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
inst, samp = collections.Counter(), collections.Counter()
base = None
for r in rows[hi + 1:]:
    if len(r) < len(hdr):
        continue
    if r[ca] == "Address":
        break
    a = int(r[ca], 16)
    base = a if base is None else base
    fn = off2fn.get(a - base, "?")
    inst[fn] += int(r[ci] or 0); samp[fn] += int(r[cs] or 0)
tot, stot = sum(inst.values()), sum(samp.values())
print(f"total warp instructions {tot}, stall samples {stot}")
for fn, n in inst.most_common():
    print(f"  {fn:32s} inst {100 * n / tot:5.1f}%  samples {100 * samp[fn] / max(stot, 1):5.1f}%")
