"""Static SASS statistics of one kernel: opcode histogram of the whole function and of every loop (span of a backward
branch).  usage: python tools/sass_mix.py OBJ_OR_SO KERNEL_SUBSTRING [--loops] [--span LO HI]"""
import collections
import re
import subprocess
import sys


def load(obj, sub):
    txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    cur, out = None, []
    for ln in txt.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            continue
        if cur is None or sub not in cur:
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);", ln)
        if m:
            out.append((int(m.group(1), 16), m.group(2).strip()))
    return out


def opcode(ins):
    ins = re.sub(r"^@!?U?P\d+\s+", "", ins)
    return ins.split()[0].split(".")[0]


def hist(ins):
    c = collections.Counter(opcode(i) for _, i in ins)
    return ", ".join(f"{k} {v}" for k, v in c.most_common(24))


def main():
    obj, sub = sys.argv[1], sys.argv[2]
    ins = load(obj, sub)
    print(f"{sub}: {len(ins)} instructions\n  {hist(ins)}")
    if "--span" in sys.argv:
        i = sys.argv.index("--span")
        lo, hi = int(sys.argv[i + 1], 16), int(sys.argv[i + 2], 16)
        part = [x for x in ins if lo <= x[0] <= hi]
        print(f"span {lo:#x}..{hi:#x}: {len(part)}\n  {hist(part)}")
    if "--loops" in sys.argv:
        for a, i in ins:
            m = re.search(r"\bBRA\b.*?0x([0-9a-f]+)", i)
            if m and int(m.group(1), 16) < a:
                lo = int(m.group(1), 16)
                part = [x for x in ins if lo <= x[0] <= a]
                if len(part) >= 24:
                    print(f"loop {lo:#x}..{a:#x}: {len(part)} instrs\n  {hist(part)}")


if __name__ == "__main__":
    main()
