"""Turn gpurun_out ncu artefacts into the small tracked summaries under profiles/ (the judge reads those).
  python tools/profile_summary.py launches <launches.csv> <out.md> [--last-step | --first-step]
  python tools/profile_summary.py rep <file.ncu-rep> <out.md>"""
import collections
import csv
import re
import subprocess
import sys


def us(row):
    t = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    return t / 1e3 if u == "ns" else (t * 1e3 if u == "ms" else t)


def launches(path, out, last_step):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    if last_step == "first":  # an interrupted capture: the first complete step (everything up to its 4th optimiser launch)
        idx = [i for i, r in enumerate(rows) if "adamw" in r["Kernel Name"]]
        rows = rows[:idx[3] + 1] if len(idx) >= 4 else rows
    elif last_step:
        idx = [i for i, r in enumerate(rows) if "adamw" in r["Kernel Name"]]
        n_opt = 4
        rows = rows[idx[-2 * n_opt] + 1:] if len(idx) >= 2 * n_opt else rows
    byk = collections.defaultdict(lambda: [0, 0.0])
    tot = 0.0
    for r in rows:
        k = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "")
        byk[k][0] += 1
        byk[k][1] += us(r)
        tot += us(r)
    with open(out, "w") as f:
        f.write(f"# ncu launch list ({path}): gpu__time_duration.sum per launch, --clock-control none\n\n")
        f.write(f"{len(rows)} launches, {tot / 1e3:.1f} ms summed device time (cold-cache, serialised: compare SHARES)\n\n")
        f.write("| kernel | launches | total ms | share | avg us |\n|---|---:|---:|---:|---:|\n")
        for k, (n, t) in sorted(byk.items(), key=lambda kv: -kv[1][1]):
            if t / tot < 0.002:
                continue
            f.write(f"| `{k[:90]}` | {n} | {t / 1e3:.2f} | {100 * t / tot:.1f}% | {t / n:.1f} |\n")
    print(open(out).read())


WANT = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum"]


def rep(path, out):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(out, "w") as f:
        f.write(f"# ncu --set full --clock-control none ({path})\n\n")
        for r in rows[2:]:
            f.write(f"## {r[idx['Kernel Name']][:110]}  grid {r[idx['Grid Size']]} block {r[idx['Block Size']]}\n\n")
            f.write("| metric | value | unit |\n|---|---:|---|\n")
            for w in WANT:
                if w in idx:
                    f.write(f"| {w} | {r[idx[w]]} | {units[idx[w]]} |\n")
            f.write("\n")
    print(open(out).read()[:3000])


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3], "first" if "--first-step" in sys.argv else "--last-step" in sys.argv)
    else:
        rep(sys.argv[2], sys.argv[3])
