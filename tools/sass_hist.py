"""SASS opcode histogram of the innermost loops of a kernel in the built library (evidence for profiles/).
    python tools/sass_hist.py <mangled-name-substring> [min_len max_len]
Lists every backward branch (loop) of the kernel with its length, and prints the opcode histogram of the loops whose
length is within [min_len, max_len] instructions."""
import collections
import re
import subprocess
import sys

LIB = "llampc_b200/libllampc_b200.so"


def kernels():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    cur, body = None, {}
    for line in out.split("\n"):
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            body[cur] = []
        elif cur:
            body[cur].append(line)
    return body


def main():
    pat = sys.argv[1]
    lo, hi = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (100, 100000)
    for name, lines in kernels().items():
        if pat not in name:
            continue
        ins = []
        for l in lines:
            m = re.search(r"/\*([0-9a-f]{4,})\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", l)
            if m:
                ins.append((int(m.group(1), 16), m.group(3), l))
        loops = []
        for a, op, l in ins:
            if op.startswith("BRA"):
                t = re.search(r"0x([0-9a-f]+)", l.split("BRA")[1])
                if t and int(t.group(1), 16) < a:
                    loops.append((int(t.group(1), 16), a))
        print("== %s: %d instructions, loops (length in instructions): %s" % (name, len(ins), [(b - a) // 16 + 1 for a, b in loops]))
        for a, b in loops:
            n = (b - a) // 16 + 1
            if lo <= n <= hi:
                c = collections.Counter(op.split(".")[0] for x, op, l in ins if a <= x <= b)
                packed = c["FFMA2"] + c["FMUL2"] + c["FADD2"]
                scal = c["FFMA"] + c["FMUL"] + c["FADD"]
                print("   loop of %d: packed f32x2 %d (FFMA2 %d FMUL2 %d FADD2 %d) | scalar FMA-pipe %d | MUFU %d | FMA-pipe cycles %d"
                      % (n, packed, c["FFMA2"], c["FMUL2"], c["FADD2"], scal, c["MUFU"], 2 * packed + scal))
                print("   " + ", ".join("%s %d" % kv for kv in c.most_common(24)))


if __name__ == "__main__":
    main()
