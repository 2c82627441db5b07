"""Developer aid: attribute the SASS instructions of a kernel to source functions
(needs -lineinfo).  usage: python tools_sass_profile.py <object.o> <kernel-substring>"""
import collections, os, re, subprocess, sys, tempfile
obj, pat = sys.argv[1], sys.argv[2]
d = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=d, stdout=subprocess.DEVNULL)
cub = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
sass = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(d, cub)], stdout=subprocess.PIPE, text=True).stdout
cur = None; kern = None; counts = collections.Counter()
for line in sass.split("\n"):
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', line)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    m = re.match(r'\.text\.(\S+):', line)
    if m:
        kern = m.group(1); continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/', line) and kern and pat in kern:
        counts[cur] += 1
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = open(os.path.join(root, "parpe_b200/csrc/bdf_core.cuh")).read().split('\n')
funcs = []
for i, l in enumerate(src, 1):
    m = re.match(r'\s*(?:template <[^>]*>\s*)?PB_HD(?:_NOINLINE)?\s+(?:static\s+)?[\w:<>\s\*&]+?\s(\w+)\(', l)
    if m: funcs.append((i, m.group(1)))
def fn_of(ln):
    name = '?'
    for i, n in funcs:
        if i <= ln: name = n
        else: break
    return name
agg = collections.Counter()
for (f, ln), c in counts.items():
    agg[fn_of(ln) if f == 'bdf_core.cuh' else f] += c
tot = sum(agg.values()); print("total", tot)
for k, v in agg.most_common(45): print("%-28s %6d %5.1f%%" % (k, v, 100 * v / tot))
