"""SASS bytes per kernel of an object file (the fused kernels must stay below the 32 KB instruction cache)."""
import re, subprocess, sys
out = subprocess.run(["cuobjdump", "-sass", sys.argv[1]], capture_output=True, text=True).stdout
name, last = None, {}
for l in out.split("\n"):
    m = re.search(r"Function : (\S+)", l)
    if m:
        name = m.group(1)
    m = re.match(r"\s*/\*([0-9a-f]{4,6})\*/", l)
    if m and name:
        last[name] = int(m.group(1), 16) + 16
for n, v in sorted(last.items(), key=lambda kv: -kv[1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 12]:
    print(f"{v:7d} B  {subprocess.run(['c++filt', n], capture_output=True, text=True).stdout.strip()[:120]}")
