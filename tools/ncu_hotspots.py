"""Join an ncu SASS-level source page with nvdisasm line info and aggregate executed instructions / stall samples
per CUDA source line and per function.

    ncu -i prof.ncu-rep --page source --csv --print-source sass > sass.csv
    python tools/ncu_hotspots.py sass.csv track-mjx_b200/csrc/libtmjx.so tmjx_env_kernelILb1 [top_n]
"""
import csv
import re
import subprocess
import sys
import tempfile
from collections import defaultdict


def main():
    sass_csv, so, kern = sys.argv[1:4]
    import os
    so = os.path.abspath(so)
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, check=True, capture_output=True)
    import glob

    dis = []   # one cubin per translation unit (the residency variants are separate TUs)
    for cubin in sorted(glob.glob(tmp + "/*.cubin")):
        dis += subprocess.run(["nvdisasm", "-gi", "-c", cubin], capture_output=True, text=True, check=True).stdout.splitlines()
    addr2line, inside, src_file = {}, False, None
    frames, last = [], None   # marker lines since the previous instruction (innermost first)
    for ln in dis:
        if ln.startswith("//---") and ".text." in ln:
            inside = kern in ln
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            frames.append((m.group(1), int(m.group(2))))
            continue
        m = re.match(r"\s*/\*([0-9a-f]+)\*/", ln)
        if m:
            if frames:
                mine = [f for f in frames if f[0].endswith(".cu")]
                if mine:
                    last = mine[0][1]
                    src_file = mine[0][0]
                frames = []
            if last is not None:
                addr2line[int(m.group(1), 16)] = last
    rows = list(csv.reader(open(sass_csv)))
    # find the block of the requested kernel
    want = "(bool)1" if "Lb1" in kern else "(bool)0"
    start = next(i for i, r in enumerate(rows) if r and r[0] == "Kernel Name" and (want in r[1] or ("1>" in r[1] if "Lb1" in kern else "0>" in r[1])))
    hdr = rows[start + 1]
    ci = {h: i for i, h in enumerate(hdr)}
    base = None
    per_line = defaultdict(lambda: [0, 0, 0])
    total = [0, 0, 0]
    reasons = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    per_line_reason = defaultdict(lambda: defaultdict(int))
    for r in rows[start + 2:]:
        if not r or r[0] == "Kernel Name":
            break
        a = int(r[ci["Address"]], 16) if r[ci["Address"]].startswith("0x") else int(r[ci["Address"]])
        if base is None:
            base = a
        line = addr2line.get(a - base, -1)
        ex = int(r[ci["Instructions Executed"]] or 0)
        th = int(r[ci["Thread Instructions Executed"]] or 0)
        sm = int(r[ci["# Samples"]] or 0)
        for acc in (per_line[line], total):
            acc[0] += ex; acc[1] += th; acc[2] += sm
        for h in reasons:
            v = r[ci[h]]
            if v and v != "0":
                per_line_reason[line][h] += int(v)
    src = open(src_file).read().splitlines() if src_file else []
    # function ranges: crude scan for "__device__ ... name(" / "__global__"
    funcs = []
    for i, l in enumerate(src, 1):
        m = re.match(r"\s*(?:template <[^>]*>\s*)?(?:__device__|__global__)[^;]*?\b(\w+)\s*\(", l)
        if m and not l.strip().endswith(";"):
            funcs.append((i, m.group(1)))
    def func_of(line):
        name = "?"
        for i, n in funcs:
            if i <= line:
                name = n
        return name
    per_func = defaultdict(lambda: [0, 0, 0])
    for line, v in per_line.items():
        f = func_of(line)
        for k in range(3):
            per_func[f][k] += v[k]
    per_func_reason = defaultdict(lambda: defaultdict(int))
    tot_reason = defaultdict(int)
    for line, d in per_line_reason.items():
        for h, v in d.items():
            per_func_reason[func_of(line)][h] += v
            tot_reason[h] += v
    print(f"total warp-instr {total[0]:,}  avg active threads {total[1] / max(total[0], 1):.1f}  samples {total[2]:,}")
    allr = sum(tot_reason.values()) or 1
    print("stall reasons (all samples): " + "  ".join(f"{h[6:]} {100 * v / allr:.1f}%" for h, v in sorted(tot_reason.items(), key=lambda kv: -kv[1])[:8]))
    print("--- per function (warp-instr %, samples %, avg threads, top stall reasons)")
    for f, v in sorted(per_func.items(), key=lambda kv: -kv[1][2]):
        rs = per_func_reason[f]
        tr = sum(rs.values()) or 1
        top3 = " ".join(f"{h[6:]}:{100 * x / tr:.0f}" for h, x in sorted(rs.items(), key=lambda kv: -kv[1])[:4])
        print(f"{f:22s} instr {100 * v[0] / total[0]:5.1f}%  samples {100 * v[2] / max(total[2], 1):5.1f}%  thr {v[1] / max(v[0], 1):4.1f}  {top3}")
    print(f"--- top {top} lines by stall samples")
    for line, v in sorted(per_line.items(), key=lambda kv: -kv[1][2])[:top]:
        text = src[line - 1].strip()[:110] if 0 < line <= len(src) else ""
        print(f"{line:5d} instr {100 * v[0] / total[0]:5.2f}% samples {100 * v[2] / max(total[2], 1):5.2f}% thr {v[1] / max(v[0], 1):4.1f} | {text}")


if __name__ == "__main__":
    main()
