"""Per-function dynamic opcode mix of the step kernel from an ncu SASS source page (innermost inlined frame, .cu or .cuh).

    python tools/ncu_opmix.py sass.csv track-mjx_b200/csrc/libtmjx.so tmjx_env_kernelILb1ELi14 [nfunc] [nops]
"""
import csv, glob, os, re, subprocess, sys, tempfile
from collections import defaultdict, Counter

def main():
    sass_csv, so, kern = sys.argv[1:4]
    nfunc = int(sys.argv[4]) if len(sys.argv) > 4 else 14
    nops = int(sys.argv[5]) if len(sys.argv) > 5 else 10
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, check=True, capture_output=True)
    dis = []   # one cubin per translation unit (the residency variants are separate TUs)
    for cubin in sorted(glob.glob(tmp + "/*.cubin")):
        dis += subprocess.run(["nvdisasm", "-gi", "-c", cubin], capture_output=True, text=True, check=True).stdout.splitlines()
    addr2, inside, frames, last = {}, False, [], None
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
                mine = [f for f in frames if f[0].endswith((".cu", ".cuh"))]
                if mine:
                    last = mine[0]
                frames = []
            if last is not None:
                addr2[int(m.group(1), 16)] = last
    fcache = {}
    def func_of(fl):
        f, line = fl
        if f not in fcache:
            fs = []
            try:
                for i, l in enumerate(open(f).read().splitlines(), 1):
                    m = re.match(r"\s*(?:template <[^>]*>\s*)?(?:__device__|__global__)[^;]*?\b(\w+)\s*\(", l)
                    if m and not l.strip().endswith(";"):
                        fs.append((i, m.group(1)))
            except OSError:
                pass
            fcache[f] = fs
        name = "?"
        for i, n in fcache[f]:
            if i <= line:
                name = n
        return ("gen::" if f.endswith(".cuh") else "") + name
    rows = list(csv.reader(open(sass_csv)))
    hdr = rows[1]; ci = {h: i for i, h in enumerate(hdr)}
    base = None
    per = defaultdict(Counter); pers = defaultdict(Counter); tot = 0; tots = 0
    for r in rows[2:]:
        if not r or r[0] == "Kernel Name":
            break
        a = int(r[ci["Address"]], 16)
        if base is None:
            base = a
        fl = addr2.get(a - base)
        fn = func_of(fl) if fl else "?"
        m = re.match(r"\s*(@!?U?P\w+\s+)?([A-Z0-9_]+)", r[ci["Source"]])
        op = m.group(2) if m else "?"
        n = int(r[ci["Instructions Executed"]] or 0); s = int(r[ci["# Samples"]] or 0)
        per[fn][op] += n; pers[fn][op] += s; tot += n; tots += s
    for fn, c in sorted(per.items(), key=lambda kv: -sum(pers[kv[0]].values()))[:nfunc]:
        n = sum(c.values()); s = sum(pers[fn].values())
        print(f"{fn:24s} instr {100*n/tot:5.1f}%  samples {100*s/tots:5.1f}% | " + " ".join(f"{op}:{100*v/n:.0f}" for op, v in c.most_common(nops)))

if __name__ == "__main__":
    main()
