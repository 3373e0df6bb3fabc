"""Static SASS instruction count per source line of one kernel (needs -lineinfo, which the build
uses): where the instructions of the lean Radau kernel come from, without GPU time.

    python scripts/sass_lines.py scipy_dae_b200/variants/libb200dae_X.so [kernel-substring] [top]
"""
import os
import re
import subprocess
import sys
import tempfile
from collections import Counter, defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KERNEL = "radau_kernelINS_9RobertsonELi3ELi64ELi6ELb1ELb1ELb0EEE"


def main():
    lib = os.path.abspath(sys.argv[1])
    kern = (sys.argv[2] if len(sys.argv) > 2 else "") or KERNEL
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    with tempfile.TemporaryDirectory() as d:
        subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=d, capture_output=True)
        txt = ""
        for f in os.listdir(d):
            if f.endswith(".cubin") and "sm_100a" in f:
                r = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(d, f)], capture_output=True, text=True)
                if kern in r.stdout:
                    txt = r.stdout
                    break
    lines = txt.split("\n")
    start = next(i for i, l in enumerate(lines) if l.startswith(".text.") and kern in l)
    end = next(i for i in range(start + 1, len(lines)) if lines[i].startswith(".text.") or lines[i].startswith(".section"))
    cur, cnt, ops = None, Counter(), defaultdict(Counter)
    for l in lines[start:end]:
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(.*?);", l)
        if m and cur:
            t = re.sub(r"^@!?U?P\d+\s+", "", m.group(1).strip())
            cnt[cur] += 1
            ops[cur][t.split()[0].split(".")[0]] += 1
    total = sum(cnt.values())
    allops = Counter()
    for c in ops.values():
        allops.update(c)
    print(f"{total} instructions; {dict(allops.most_common(12))}")
    src = {}
    for (f, n), c in cnt.most_common(top):
        if f not in src:
            path = os.path.join(ROOT, "scipy_dae_b200", "csrc", f)
            src[f] = open(path).read().split("\n") if os.path.exists(path) else None
        text = src[f][n - 1].strip()[:80] if src[f] and n <= len(src[f]) else ""
        print(f"{c:4d} {f}:{n} {dict(ops[(f, n)].most_common(4))} | {text}")


if __name__ == "__main__":
    main()
