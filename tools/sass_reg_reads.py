"""Register-operand reads of the innermost loops of a kernel (evidence for profiles/): B200 delivers ~1.86 32-bit register
operands per clock per scheduler (tools/ubench/ffma2_operands.cu), so a loop body cannot run faster than reads / 1.86.
    python tools/sass_reg_reads.py <file.so|.o> <mangled-name-substring> [min_len max_len]
Counts, per loop, the register source operands of every instruction (a packed F32x2 operand = 2 reads, R.F32 = 1, RZ /
immediates / constant bank / uniform registers = 0); an operand that the PREVIOUS instruction flagged `.reuse` in the same
slot with the same register is counted as served by the reuse cache."""
import collections
import re
import subprocess
import sys


def main():
    lib, pat = sys.argv[1], sys.argv[2]
    lo, hi = (int(sys.argv[3]), int(sys.argv[4])) if len(sys.argv) > 4 else (300, 1000)
    out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    cur, body = None, {}
    for line in out.split("\n"):
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            body[cur] = []
        elif cur:
            body[cur].append(line)
    for name, lines in body.items():
        if pat not in name:
            continue
        ins = []
        for l in lines:
            m = re.search(r"/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
            if m:
                ins.append((int(m.group(1), 16), m.group(2).strip()))
        loops = []
        for a, t in ins:
            if "BRA" in t:
                m = re.search(r"0x([0-9a-f]+)", t.split("BRA")[1])
                if m and int(m.group(1), 16) < a:
                    loops.append((int(m.group(1), 16), a))
        for a, b in loops:
            n = (b - a) // 16 + 1
            if not (lo <= n <= hi):
                continue
            reads = collections.Counter()
            served = 0
            prev = {}
            fma_cycles = 0
            for addr, t in ins:
                if not (a <= addr <= b):
                    continue
                t = re.sub(r"^@!?U?P\d+\s+", "", t)
                op, _, rest = t.partition(" ")
                opn = op.split(".")[0]
                srcs = [o.strip() for o in rest.split(",")][1:]
                cur_flags = {}
                for slot, s in enumerate(srcs):
                    m = re.match(r"[-|~]*R(\d+)", s)
                    if not m or "RZ" in s:
                        continue
                    w = 2 if "F32x2" in s or ".64" in s else 1
                    reg = m.group(1)
                    if prev.get(slot) == reg:
                        served += w
                    else:
                        reads[opn] += w
                    if ".reuse" in s:
                        cur_flags[slot] = reg
                prev = cur_flags
                if opn in ("FFMA2", "FMUL2", "FADD2"):
                    fma_cycles += 2
                elif opn in ("FFMA", "FMUL", "FADD", "IMAD"):
                    fma_cycles += 1
            tot = sum(reads.values())
            print("%s\n   loop of %d instructions: %d register reads (+ %d served by the reuse cache) -> %.0f clocks at 1.86 reads/clk; "
                  "FMA-pipe clocks %d" % (name[:90], n, tot, served, tot / 1.86, fma_cycles))
            print("   " + ", ".join("%s %d" % kv for kv in reads.most_common(10)))


if __name__ == "__main__":
    main()
