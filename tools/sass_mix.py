"""Developer tool: static SASS instruction mix per kernel of a cubin / .so (cuobjdump -sass)."""
import collections, re, subprocess, sys
out = subprocess.run(["cuobjdump", "-sass", sys.argv[1]], capture_output=True, text=True).stdout
pat = sys.argv[2] if len(sys.argv) > 2 else ""
fn, mix = None, collections.defaultdict(collections.Counter)
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        fn = m.group(1); continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_]+)", line)
    if m and fn:
        mix[fn][m.group(1)] += 1
for f, c in mix.items():
    if pat in f:
        print(f, sum(c.values()), ", ".join(f"{k} {v}" for k, v in c.most_common(14)))
