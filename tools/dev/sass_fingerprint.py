#!/usr/bin/env python
"""Per-function SASS fingerprints of the built objects (cuobjdump -sass, split at 'Function :').

Used to show that a change behind a template switch leaves the default kernels' machine code untouched:
    python tools/dev/sass_fingerprint.py > before.json ; <edit, rebuild> ; python tools/dev/sass_fingerprint.py --diff before.json
"""
import concurrent.futures, glob, hashlib, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
OBJ = os.path.join(ROOT, "parpe_b200", "csrc", "build")


def demangle(names):
    r = subprocess.run(["c++filt"], input="\n".join(names), stdout=subprocess.PIPE, text=True)
    return r.stdout.split("\n")


def one(obj):
    out = subprocess.run(["cuobjdump", "-sass", obj], stdout=subprocess.PIPE, text=True).stdout
    funcs, name, buf = {}, None, []
    for line in out.splitlines():
        s = line.strip()
        if s.startswith("Function :"):
            if name:
                funcs[name] = buf
            name, buf = s.split(":", 1)[1].strip(), []
        elif name and s.startswith("/*"):
            buf.append(s)
    if name:
        funcs[name] = buf
    names = list(funcs)
    return {os.path.basename(obj) + "::" + d: [hashlib.md5("\n".join(funcs[n]).encode()).hexdigest(), len(funcs[n]) // 2]
            for n, d in zip(names, demangle(names))}


def main():
    objs = sorted(glob.glob(os.path.join(OBJ, "*.o")))
    res = {}
    with concurrent.futures.ThreadPoolExecutor(8) as ex:
        for d in ex.map(one, objs):
            res.update(d)
    if len(sys.argv) > 2 and sys.argv[1] == "--diff":
        old = json.load(open(sys.argv[2]))
        same = [k for k in res if k in old and old[k] == res[k]]
        print("unchanged functions: %d" % len(same))
        for k in sorted(set(res) | set(old)):
            if k not in old:
                print("NEW      %s (%d instructions)" % (k[:200], res[k][1]))
            elif k not in res:
                print("REMOVED  %s" % k[:200])
            elif old[k] != res[k]:
                print("CHANGED  %s (%d -> %d instructions)" % (k[:200], old[k][1], res[k][1]))
    else:
        json.dump(res, sys.stdout, indent=0)


if __name__ == "__main__":
    main()
