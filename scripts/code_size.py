"""Code-size accounting of a kernel: SASS bytes per source function (dev tool; the Newton loop is instruction-fetch
sensitive). Usage: python scripts/code_size.py [kernel-substring] [lib.so]"""
import collections, os, re, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
kern = sys.argv[1] if len(sys.argv) > 1 else "k_forward_fast"
lib = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "qpth_b200", "libqpth_b200.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, capture_output=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
# function table of the sources: line -> enclosing function name
fn_of = {}
for f in ("qp_fast.cuh", "qp_device.cuh", "qp_kernels.cu"):
    cur = "?"
    for i, line in enumerate(open(os.path.join(ROOT, "qpth_b200", "csrc", f)), 1):
        if re.match(r"^(template|__device__|__global__|k_\w+\()", line) or re.match(r"^\s*__device__", line):
            m = re.search(r"\b(\w+)\s*\(", line)
            if m and m.group(1) not in ("__launch_bounds__", "template", "__align__"):
                cur = m.group(1)
        fn_of[(f, i)] = cur
sec = None; cur = ("?", 0); by_fn = collections.Counter(); by_line = collections.Counter(); total = 0
for line in dis.splitlines():
    if line.startswith(".text."):
        sec = line
        continue
    if sec is None or kern not in sec:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"^\s+/\*[0-9a-f]{4,}\*/", line):
        total += 16
        by_fn[fn_of.get(cur, cur[0])] += 16
        by_line[cur] += 16
print("%s: %.1f KB of SASS" % (kern, total / 1024))
for fn, b in by_fn.most_common(40):
    print("  %6.1f KB  %s" % (b / 1024, fn))
if len(sys.argv) > 3:
    for (f, l), b in by_line.most_common(40):
        print("  %6.2f KB  %s:%d [%s]" % (b / 1024, f, l, fn_of.get((f, l), "?")))
