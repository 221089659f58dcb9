"""Shared-memory wavefronts and bank-conflict replays per source line (innermost .cu frame) from an ncu SASS source page.
    python tools/ncu_smem_conflicts.py sass.csv track-mjx_b200/csrc/libtmjx.so tmjx_env_kernelILb1ELi14 [top]"""
import csv, glob, os, re, subprocess, sys, tempfile
from collections import defaultdict
sass_csv, so, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 20
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, check=True, capture_output=True)
addr2, inside, frames, last = {}, False, [], None
for cubin in sorted(glob.glob(tmp + "/*.cubin")):
    for ln in subprocess.run(["nvdisasm", "-gi", "-c", cubin], capture_output=True, text=True, check=True).stdout.splitlines():
        if ln.startswith("//---") and ".text." in ln:
            inside = kern in ln
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            frames.append((os.path.basename(m.group(1)), int(m.group(2))))
            continue
        m = re.match(r"\s*/\*([0-9a-f]+)\*/", ln)
        if m:
            if frames:
                cu = [f for f in frames if f[0].endswith(".cu")]
                last = cu[0] if cu else frames[0]
                frames = []
            if last:
                addr2[int(m.group(1), 16)] = last
rows = list(csv.reader(open(sass_csv)))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}
base, exc, tot = None, defaultdict(int), defaultdict(int)
for r in rows[2:]:
    if len(r) < 10 or r[0] == "Kernel Name":
        break
    a = int(r[ix["Address"]], 16)
    base = a if base is None else base
    fl = addr2.get(a - base)
    if fl:
        exc[fl] += int(r[ix["L1 Wavefronts Shared Excessive"]] or 0)
        tot[fl] += int(r[ix["L1 Wavefronts Shared"]] or 0)
E, T = sum(exc.values()), sum(tot.values())
print(f"shared-memory wavefronts {T:,}, of which bank-conflict replays {E:,} ({100 * E / max(T, 1):.1f} %)")
src = {}
for (f, l), e in sorted(tot.items(), key=lambda kv: -kv[1])[:top]:
    if f not in src:
        try:
            src[f] = open(os.path.join(os.path.dirname(os.path.abspath(so)), f)).read().splitlines()
        except OSError:
            src[f] = []
    text = src[f][l - 1].strip()[:90] if 0 < l <= len(src[f]) else ""
    print(f"{100 * tot[(f, l)] / T:5.1f}% wavefronts  {100 * exc[(f, l)] / max(E, 1):5.1f}% replays  {f}:{l}  {text}")
