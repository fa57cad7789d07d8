"""Turn ncu exports into the small text summaries committed under profiles/ (run in the build container, no GPU).

    python profiles/summarize.py launches <launches.csv> <out.md> [title]
    python profiles/summarize.py full <report.ncu-rep> <out.md> [title]

`launches`: the `--metrics gpu__time_duration.sum` launch list -> per-kernel count / total / share table.
`full`    : an `ncu --set full` report -> per-launch table of duration, DRAM bytes, DRAM / L2 / tensor-pipe / XU
            utilisation, registers and grid (read with `ncu -i ... --page raw --csv`).
"""
import collections
import csv
import io
import subprocess
import sys


def launches(path, out, title):
    rows = list(csv.reader(open(path)))
    hdr, agg = None, collections.defaultdict(lambda: [0, 0.0])
    for r in rows:
        if hdr is None:
            if "Kernel Name" in r:
                hdr = r
            continue
        d = dict(zip(hdr, r))
        try:
            v = float(d["Metric Value"].replace(",", ""))
        except (KeyError, ValueError):
            continue
        k = d["Kernel Name"].split("(")[0].replace("void ", "")
        agg[k][0] += 1
        agg[k][1] += v
    tot = sum(v[1] for v in agg.values())
    n = sum(v[0] for v in agg.values())
    with open(out, "w") as f:
        f.write(f"# {title}\n\n{n} launches, {tot / 1e6:.3f} ms of kernel time (ncu per-launch times are cold-cache and "
                f"serialised: compare shares, not absolutes)\n\n| kernel | launches | total us | share |\n|---|---:|---:|---:|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{k}` | {v[0]} | {v[1] / 1e3:.1f} | {100 * v[1] / tot:.1f}% |\n")


FULL_COLS = [
    ("gpu__time_duration.sum", "time"), ("dram__bytes_read.sum", "dram rd"), ("dram__bytes_write.sum", "dram wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 %"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps %"),
    ("launch__registers_per_thread", "regs"), ("launch__grid_size", "grid"), ("launch__block_size", "block"),
]


def full(path, out, title):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    with open(out, "w") as f:
        f.write(f"# {title}\n\nsource: `ncu --set full --clock-control none --import-source on` ({path.split('/')[-1]}); "
                f"one row per captured launch\n\n| kernel | " + " | ".join(n for _, n in FULL_COLS) + " |\n|---|" +
                "---:|" * len(FULL_COLS) + "\n")
        for r in rows[2:]:
            name = r[ix["Kernel Name"]].split("(")[0].replace("void ", "")
            cells = []
            for key, _ in FULL_COLS:
                if key in ix:
                    v = r[ix[key]]
                    try:
                        v = f"{float(v.replace(',', '')):.4g} {units[ix[key]]}".strip()
                    except ValueError:
                        pass
                    cells.append(v)
                else:
                    cells.append("-")
            f.write(f"| `{name}` | " + " | ".join(cells) + " |\n")


if __name__ == "__main__":
    mode, src, dst = sys.argv[1:4]
    title = sys.argv[4] if len(sys.argv) > 4 else src
    (launches if mode == "launches" else full)(src, dst, title)
