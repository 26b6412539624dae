#!/usr/bin/env python
"""Per-source-line instruction and stall totals of one kernel: joins `ncu --page source --csv`
(per-SASS-instruction counters of an .ncu-rep) with `nvdisasm -gi` line info of the object that was
profiled, attributing inlined code to the line of the kernel's own file that pulled it in.

    python tools/ncu_hotspots.py REPORT.ncu-rep OBJECT.o KERNEL_REGEX KERNEL_FILE.cu [--top 40] [--launch-index 0]
"""
import argparse
import collections
import csv
import glob
import io
import os
import re
import subprocess
import sys
import tempfile


def disasm_lines(obj, kernel_regex, kernel_file):
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, check=True, capture_output=True)
    cubin = glob.glob(os.path.join(tmp, "*.cubin"))[0]
    text = subprocess.run(["nvdisasm", "-gi", "-c", cubin], check=True, capture_output=True, text=True).stdout
    out, active, line, opclass = [], False, None, None
    base = os.path.basename(kernel_file)
    for row in text.splitlines():
        m = re.match(r"\s*\.section\s+\.text\.(\S+?),", row)
        if m:
            active = re.search(kernel_regex, m.group(1)) is not None
            continue
        if not active:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', row)
        if m:
            if os.path.basename(m.group(1)) == base:
                line = int(m.group(2))
            elif m.group(3) and os.path.basename(m.group(3)) == base:
                line = int(m.group(4))
            # deeper inlining chains keep the previous kernel-file line
            continue
        m = re.match(r"\s*/\*([0-9a-f]+)\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", row)
        if m:
            out.append((line, m.group(2)))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("obj")
    ap.add_argument("kernel")
    ap.add_argument("file")
    ap.add_argument("--top", type=int, default=40)
    ap.add_argument("--section", default=None, help="regex on the MANGLED name for the disassembly (default: the kernel regex)")
    args = ap.parse_args()
    raw = subprocess.run(["ncu", "-i", args.report, "--page", "source", "--csv", "--kernel-name", f"regex:{args.kernel}"],
                         check=True, capture_output=True, text=True).stdout
    # several launches may match: keep the first block
    blocks = raw.split('"Kernel Name"')
    body = blocks[1].split("\n", 1)[1]
    rows = list(csv.DictReader(io.StringIO(body)))
    dis = disasm_lines(args.obj, args.section or args.kernel, args.file)
    if len(dis) != len(rows):
        print(f"warning: {len(rows)} profiled instructions vs {len(dis)} disassembled", file=sys.stderr)
    src = open(args.file).read().splitlines()
    per_line = collections.defaultdict(lambda: collections.Counter())
    per_op = collections.Counter()
    total = collections.Counter()
    for (line, op), row in zip(dis, rows):
        sass_op = row["Source"].split()[0] if not row["Source"].strip().startswith("@") else row["Source"].split()[1]
        n = int(row["Instructions Executed"] or 0)
        s = int(row["# Samples"] or 0)
        per_line[line]["inst"] += n
        per_line[line]["samples"] += s
        per_op[sass_op.split(".")[0]] += n
        total["inst"] += n
        total["samples"] += s
        for k in ("stall_wait", "stall_math", "stall_short_sb", "stall_long_sb", "stall_not_selected", "stall_dispatch",
                  "stall_no_inst", "stall_barrier", "stall_mio", "stall_selected"):
            per_line[line][k] += int(row.get(k) or 0)
            total[k] += int(row.get(k) or 0)
    print(f"# total warp instructions {total['inst']}, samples {total['samples']}")
    print("# stalls: " + ", ".join(f"{k[6:]} {100.0 * v / max(1, total['samples']):.1f}%" for k, v in total.items() if k.startswith("stall_")))
    print("# opcode mix: " + ", ".join(f"{k} {100.0 * v / total['inst']:.1f}%" for k, v in per_op.most_common(18)))
    print(f"# {'line':>5} {'inst%':>6} {'smpl%':>6}  source")
    for line, c in sorted(per_line.items(), key=lambda kv: -kv[1]["samples"])[:args.top]:
        text = src[line - 1].strip()[:110] if line and line <= len(src) else "?"
        print(f"  {str(line):>5} {100.0 * c['inst'] / total['inst']:6.2f} {100.0 * c['samples'] / max(1, total['samples']):6.2f}  {text}")


if __name__ == "__main__":
    main()
