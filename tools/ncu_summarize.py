"""Condense the two ncu exports of tools/gpu/ncu.sh into what profiles/ keeps:
  launches_<tag>.csv (as captured)  ->  per-kernel totals of one UNet forward
  prof_<tag>_raw.csv (--page raw)   ->  prof_<tag>_key_metrics.csv + a per-kernel table."""
import csv, re, sys, collections

tag = sys.argv[1]
src = sys.argv[2] if len(sys.argv) > 2 else "gpurun_out"

def short(name):
    name = re.sub(r"\(.*", "", name)
    return name.replace("advs::", "").replace("void ", "").strip()

# ---- launch list
rows = [r for r in csv.reader(l for l in open(f"{src}/launches_{tag}.csv") if not l.startswith("==")) if len(r) > 5]
hdr = rows[0]
kn, mv = hdr.index("Kernel Name"), hdr.index("Metric Value")
# keep what lies between two k_ddim_step launches two apart: one forward of each of the two 32-image sub-batches
body = rows[1:]
steps = [i for i, r in enumerate(body) if short(r[kn]) == "k_ddim_step"]
if len(steps) >= 3:
    body = body[steps[0] + 1:steps[2] + 1]
    print(f"(window: launches {steps[0] + 1}..{steps[2]} of the capture = two sub-batch forwards + their DDIM updates)")
agg = collections.OrderedDict()
for r in body:
    d = agg.setdefault(short(r[kn]), [0, 0.0])
    d[0] += 1
    d[1] += float(r[mv].replace(",", "")) / 1e6      # ns -> ms
total = sum(v[1] for v in agg.values())
print(f"launch list ({sum(v[0] for v in agg.values())} launches, {total:.1f} ms serialised, un-throttled clocks)\n")
print("| kernel | launches | ms | share |\n|---|---|---|---|")
for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"| `{k}` | {n} | {ms:.3f} | {ms / total:.3f} |")

# ---- full capture
rows = list(csv.reader(open(f"{src}/prof_{tag}_raw.csv")))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
keys = ["Kernel Name", "gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "FBSP.TriageCompute.dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
keys = [k for k in keys if k in col]
with open(f"profiles/prof_{tag}_key_metrics.csv", "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(keys)
    w.writerow([units[col[k]] for k in keys])
    for r in data:
        w.writerow([short(r[col[k]]) if k == "Kernel Name" else r[col[k]] for k in keys])
per = collections.OrderedDict()
for r in data:
    per.setdefault(short(r[col["Kernel Name"]]), []).append(r)
def f(r, k):
    v = r[col[k]].replace(",", "")
    try:
        return float(v)
    except ValueError:
        return float("nan")
print("\n| kernel | launches captured | duration us (min-max) | tensor pipe active % | XU (MUFU) % | LSU wavefronts % | DRAM rd+wr MB / launch | DRAM throughput % | L2 hit % |\n|---|---|---|---|---|---|---|---|---|")
for k, rs in per.items():
    rng = lambda key: f"{min(f(r, key) for r in rs):.1f}-{max(f(r, key) for r in rs):.1f}"
    mb = [f(r, "dram__bytes_read.sum") + f(r, "dram__bytes_write.sum") for r in rs]
    print(f"| `{k}` | {len(rs)} | {rng('gpu__time_duration.sum')} | {rng('sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed')} | "
          f"{rng('sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed')} | {rng('l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed')} | "
          f"{min(mb):.0f}-{max(mb):.0f} | {rng('FBSP.TriageCompute.dram__throughput.avg.pct_of_peak_sustained_elapsed')} | {rng('lts__t_sector_hit_rate.pct')} |")
