"""Static look at a kernel's SASS (no GPU): instruction mix of the whole function and of its hottest-looking loop bodies.
    python tools/sass_loop.py <lib.so> <mangled-function-substring>
Loops are found as backward branches; for each, the opcode histogram of the body is printed (largest bodies first)."""
import re, subprocess, sys
from collections import Counter

def main(lib, pat):
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    cur, funcs = None, {}
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1); funcs[cur] = []; continue
        m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
        if m and cur:
            funcs[cur].append((int(m.group(1), 16), m.group(2).strip()))
    for name, ins in funcs.items():
        if pat not in name:
            continue
        print(name, len(ins), "instructions")
        addr_idx = {a: i for i, (a, _) in enumerate(ins)}
        loops = []
        for i, (a, s) in enumerate(ins):
            m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d,\s*)?0x([0-9a-f]+)", s)
            if m:
                t = int(m.group(1), 16)
                if t <= a and t in addr_idx:
                    loops.append((addr_idx[t], i))
        for lo, hi in sorted(loops, key=lambda p: -(p[1] - p[0]))[:14]:
            body = ins[lo:hi + 1]
            def op(s):
                t = s.split()
                o = t[1] if t[0].startswith("@") else t[0]
                return o.split(".")[0]
            c = Counter(op(s) for _, s in body)
            print(f"  loop {ins[lo][0]:#x}..{ins[hi][0]:#x}: {len(body)} instrs", dict(c.most_common(12)))

if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
