"""Per-source-line cost of a kernel from an ncu capture (--import-source on) + the cubin's line table.

    python tools/ncu_lines.py gpurun_out/X.ncu-rep k_shade [launch_index] [top_n]

ncu's CSV source page is per SASS instruction; nvdisasm -g gives the source line of every SASS instruction of the same
cubin (built with -lineinfo).  The two are joined by instruction order.  Prints, per source line, warp instructions,
thread instructions (-> active lanes) and stall samples, sorted by samples."""
import collections, csv, os, re, subprocess, sys, tempfile

rep, kernel = sys.argv[1], sys.argv[2]
launch = int(sys.argv[3]) if len(sys.argv) > 3 else 0
top = int(sys.argv[4]) if len(sys.argv) > 4 else 45
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(root, "dogeray_b200", "libdogeray_b200.so")
tu = "render" if kernel in ("k_shade", "k_trace", "k_generate", "k_resolve") else "scene"
with tempfile.TemporaryDirectory() as td:
    subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=td, check=True, stdout=subprocess.DEVNULL)
    cub = [f for f in os.listdir(td) if f.startswith(tu) and f.endswith(".cubin")][0]
    sass = subprocess.run(["nvdisasm", "-g", "-c", cub], cwd=td, check=True, capture_output=True, text=True).stdout
lines, cur, infn = [], None, False
for ln in sass.splitlines():
    if ln.startswith(".text."):
        infn = re.search(r"\d+%s[A-Z]" % kernel, ln) is not None
        continue
    if not infn:
        continue
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
        lines.append(cur)
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"]
starts.append(len(rows))
mine = [k for k in range(len(starts) - 1) if "::%s(" % kernel in rows[starts[k]][1]]       # sections of this kernel, in capture order
if len(starts) - 1 > 0 and len(mine) % 2 == 0 and all(rows[starts[mine[2 * j]] + 2] == rows[starts[mine[2 * j + 1]] + 2] for j in range(len(mine) // 2)):
    mine = mine[::2]                                       # this ncu prints every launch's table twice
k = mine[launch]
s = starts[k]
hdr = rows[s + 1]
body = [r for r in rows[s + 2:starts[k + 1]] if len(r) == len(hdr)]
col = {n: i for i, n in enumerate(hdr)}
assert len(body) == len(lines), (len(body), len(lines))
agg = collections.defaultdict(lambda: [0, 0, 0, 0, 0])
tot = [0, 0, 0, 0, 0]
for r, where in zip(body, lines):
    v = [int(r[col["Instructions Executed"]]), int(r[col["Thread Instructions Executed"]]), int(r[col["# Samples"]]),
         int(r[col["stall_long_sb"]]), 1]
    for k in range(5):
        agg[where][k] += v[k]; tot[k] += v[k]
src = {}
print("kernel %s launch %d: %d SASS instr, %.3g warp instr, %.1f lanes/instr, %d samples" % (kernel, launch, len(body), tot[0], tot[1] / max(tot[0], 1), tot[2]))
print("%-22s %8s %7s %6s %8s %7s %5s  source" % ("where", "winstr%", "lanes", "smpl%", "long_sb%", "winstr", "sass"))
for where, v in sorted(agg.items(), key=lambda kv: -kv[1][2])[:top]:
    f, n = where if where else ("?", 0)
    if f not in src:
        try:
            src[f] = open(os.path.join(root, "dogeray_b200", "csrc", f)).read().splitlines()
        except OSError:
            src[f] = []
    text = src[f][n - 1].strip()[:90] if 0 < n <= len(src[f]) else ""
    print("%-22s %7.2f%% %7.1f %5.1f%% %7.1f%% %7.3g %5d  %s" % ("%s:%d" % (f, n), 100 * v[0] / tot[0], v[1] / max(v[0], 1), 100 * v[2] / max(tot[2], 1),
                                                          100 * v[3] / max(tot[3], 1), v[0], v[4], text))
