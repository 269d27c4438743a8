"""Static SASS instruction mix per kernel of an object file / .so (cuobjdump -sass): a CPU-side check before spending GPU time.
usage: python tools/sass_count.py <file.o|.so> [name-substring]"""
import collections
import re
import subprocess
import sys


def main():
    path = sys.argv[1]
    pat = sys.argv[2] if len(sys.argv) > 2 else ""
    txt = subprocess.run(["cuobjdump", "-sass", path], capture_output=True, text=True, check=True).stdout
    cur, cnt = None, collections.defaultdict(collections.Counter)
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            cnt[cur][m.group(2).split(".")[0]] += 1
    for k, c in cnt.items():
        if pat not in k:
            continue
        tot = sum(c.values())
        f = c["DFMA"] + c["DMUL"] + c["DADD"]
        print("%s\n   total %d  fp64 %d (DFMA %d DMUL %d DADD %d)  LDL %d STL %d  MUFU %d  bytes %d\n   %s" % (
            k[:90], tot, f, c["DFMA"], c["DMUL"], c["DADD"], c["LDL"], c["STL"], c["MUFU"], tot * 16, dict(c.most_common(10))))


if __name__ == "__main__":
    main()
