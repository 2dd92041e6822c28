"""Static SASS profile of one kernel: instructions per source line and per opcode class.

  python tools/sass_profile.py lisflood_code_b200/csrc/lf_model.o k_soil_fusedILb0ELi64ELi5 [--top 40] [--sub]

Uses `cuobjdump -xelf` + `nvdisasm -g -c` (the objects are built with -lineinfo).  The kernel body proper (up to the
first `$kernel$callee` label of a non-inlined device function) is reported separately from its out-of-line callees,
so the common path of k_soil_fused (single Darcy sub-step) is not mixed with soil_column_overflow.  This is a static
count: loops count once, both sides of a branch count; it is used to budget instruction issue before going to the
GPU, not as a measurement.
"""
import argparse
import collections
import os
import re
import subprocess
import sys
import tempfile

FP64 = {"DFMA", "DADD", "DMUL", "DSETP"}
MEM = {"LDG", "STG", "LD", "ST", "LDS", "STS", "LDL", "STL", "LDC", "LDCU", "ATOMG", "RED", "ATOMS"}


def disassemble(obj):
    tmp = tempfile.mkdtemp(prefix="sassprof_")
    subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, stdout=subprocess.DEVNULL)
    cubins = [os.path.join(tmp, f) for f in os.listdir(tmp) if f.endswith(".cubin")]
    return subprocess.check_output(["nvdisasm", "-g", "-c"] + cubins, text=True, stderr=subprocess.DEVNULL)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("obj")
    ap.add_argument("kernel", help="substring of the mangled kernel name")
    ap.add_argument("--top", type=int, default=40)
    ap.add_argument("--sub", action="store_true", help="also profile the out-of-line callees")
    args = ap.parse_args()
    text = disassemble(args.obj)
    sec_re = re.compile(r"^\.text\.(\S+):\s*$")
    lines = text.splitlines()
    start = None
    for n, ln in enumerate(lines):
        m = sec_re.match(ln)
        if m and args.kernel in m.group(1) and n > 0 and not lines[n - 1].startswith("_Z"):
            start = n  # the second occurrence follows the header block; instructions start after it
    if start is None:
        for n, ln in enumerate(lines):
            m = sec_re.match(ln)
            if m and args.kernel in m.group(1):
                start = n
    if start is None:
        sys.exit("kernel not found")
    ins_re = re.compile(r"^\s+/\*([0-9a-f]{4,})\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_]+)")
    file_re = re.compile(r'//## File "([^"]+)", line (\d+)')
    parts = collections.OrderedDict()
    cur_part = "body"
    cur_line = ("?", 0)
    for ln in lines[start + 1:]:
        if ln.startswith("//---------------------") or ln.startswith("\t.section"):
            break
        if ln.startswith("$"):
            cur_part = ln.split("$")[-1].rstrip(":")
            continue
        m = file_re.search(ln)
        if m:
            cur_line = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = ins_re.match(ln)
        if m:
            parts.setdefault(cur_part, []).append((cur_line, m.group(2)))
    for part, ins in parts.items():
        if part != "body" and not args.sub:
            print("[%s] %d instructions (out of line; --sub to expand)" % (part[:60], len(ins)))
            continue
        ops = collections.Counter(op for _, op in ins)
        fp64 = sum(c for o, c in ops.items() if o in FP64)
        mem = sum(c for o, c in ops.items() if o in MEM)
        print("[%s] %d instructions: fp64 %d (DFMA %d DADD %d DMUL %d DSETP %d), FSEL %d, memory %d (LDG %d STG %d LDL %d STL %d "
              "LDS %d STS %d LDC %d), IMAD %d, MUFU %d, branches %d" % (
                  part[:60], len(ins), fp64, ops["DFMA"], ops["DADD"], ops["DMUL"], ops["DSETP"], ops["FSEL"], mem, ops["LDG"],
                  ops["STG"], ops["LDL"], ops["STL"], ops["LDS"], ops["STS"], ops["LDC"] + ops["LDCU"], ops["IMAD"],
                  ops["MUFU"], ops["BRA"]))
        by_line = collections.Counter(l for l, _ in ins)
        for (f, l), c in by_line.most_common(args.top):
            o = collections.Counter(op for ll, op in ins if ll == (f, l))
            print("  %5d  %s:%d   %s" % (c, f, l, " ".join("%s=%d" % kv for kv in o.most_common(5))))


if __name__ == "__main__":
    main()
