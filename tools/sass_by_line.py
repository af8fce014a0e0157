#!/usr/bin/env python
"""Static SASS instruction count per source line of one kernel (needs -lineinfo).
Usage: sass_by_line.py lib.so kernel_substring [top]"""
import collections, os, re, subprocess, sys, tempfile
lib, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
d = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=d, stdout=subprocess.DEVNULL)
cub = max((os.path.join(d, f) for f in os.listdir(d)), key=os.path.getsize)
lines = subprocess.run(["nvdisasm", "--print-line-info", cub], capture_output=True, text=True).stdout.splitlines()
start = [i for i, l in enumerate(lines) if l.startswith(".text.") and pat in l][0]
end = [i for i, l in enumerate(lines) if l.startswith(".text.") and i > start]
end = end[0] if end else len(lines)
cur = None; cnt = collections.Counter()
for l in lines[start:end]:
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1), int(m.group(2))); continue
    if re.match(r'\s+/\*[0-9a-f]+\*/\s+\S', l): cnt[cur] += 1
print("total", sum(cnt.values()))
srcs = {}
for (f, ln), c in sorted(cnt.items(), key=lambda kv: -kv[1])[:top]:
    if f not in srcs:
        try: srcs[f] = open(f).read().splitlines()
        except OSError: srcs[f] = []
    text = srcs[f][ln - 1].strip()[:100] if ln - 1 < len(srcs[f]) else ""
    print("%5d %s:%d  %s" % (c, os.path.basename(f), ln, text))
