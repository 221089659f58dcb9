"""Static count of a SASS opcode per source line (innermost .cu/.cuh frame) for one kernel: where do spills (STL/LDL) sit?
    python tools/sass_lines.py track-mjx_b200/csrc/libtmjx.so tmjx_env_kernelILb1ELi14 STL,LDL [top]"""
import glob, os, re, subprocess, sys, tempfile
from collections import Counter
so, kern, ops = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 25
ops = set(ops.split(","))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, check=True, capture_output=True)
cnt, inside, frames, last = Counter(), False, [], None
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
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_]+)", ln)
        if m:
            if frames:
                last = frames[0]
                frames = []
            if m.group(1) in ops:
                cnt[last] += 1
for (f, l), n in cnt.most_common(top):
    print(f"{n:5d}  {f}:{l}")
