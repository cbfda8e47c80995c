"""Aggregate an ncu source-page SASS CSV per CUDA source line, using nvdisasm -g line info.
usage: ncu_lines.py <report.ncu-rep> <cubin> <kernel-substring> [top]"""
import csv, re, subprocess, sys, io
rep, cubin, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
# walk: track current function and current line
addr2line = {}
cur_fn, cur_line = None, None
for ln in dis.splitlines():
    m = re.match(r"\s*\.section\s+\.text\.(\S+?),", ln)
    if m: cur_fn = m.group(1); cur_line = None; continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m: cur_line = (m.group(1).split('/')[-1], int(m.group(2))); continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*)", ln)
    if m and cur_fn and kern in cur_fn:
        addr2line[int(m.group(1), 16)] = cur_line
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
h = rows[1]; ix = {n: i for i, n in enumerate(h)}
base = None
agg = {}
tot = [0, 0]
seen_kernels = 0
for r in rows[2:]:
    if r and r[0] == "Kernel Name":
        break                      # only the first matching launch
    if len(r) < len(h) or r[ix["Address"]] == "Address": continue
    a = int(r[ix["Address"]], 16) if r[ix["Address"]].startswith("0x") else int(r[ix["Address"]])
    if base is None: base = a
    line = addr2line.get(a - base)
    s = int(r[ix["# Samples"]] or 0); i = int(r[ix["Instructions Executed"]] or 0)
    e = agg.setdefault(line, [0, 0]); e[0] += s; e[1] += i
    tot[0] += s; tot[1] += i
src = {}
print("total samples %d, warp instructions %d" % tuple(tot))
for line, (s, i) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    text = ""
    if line:
        f = "/root/repo/videoyolo_b200/csrc/" + line[0]
        try:
            if f not in src: src[f] = open(f).read().splitlines()
            text = src[f][line[1] - 1].strip()[:90]
        except Exception: pass
    print("%5.1f%% smp %5.1f%% inst  %-22s %s" % (100.0 * s / max(tot[0], 1), 100.0 * i / max(tot[1], 1), line, text))
