"""GPU box, under ncu:  DRAM bytes per launch for every (kernel family, layer shape) of one eager denoising step.

    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \\
        --log-file gpurun_out/traffic.csv python profiles/traffic_capture.py [batch]
    python profiles/traffic_capture.py --merge gpurun_out/traffic.csv gpurun_out/traffic_labels.json   (no GPU)

The first form runs bench.kernel_breakdown (which notes the process-wide ordinal of every libcnb200 launch it times) and
writes the labels; the second joins them with ncu's per-launch rows (same order: ncu lists the process's kernels in
launch order; rows of kernels outside the library's namespaces - torch's own - are dropped first) into profiles/ncu_traffic.json = {"family|shape|batch": bytes per launch}, which bench.py reports as
`roofline.traffic`."""
import collections
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def capture(batch):
    import torch
    import bench
    dev = torch.device("cuda", 0)
    cfg, model, sched, hint_host = bench.build_problem(batch, dev)
    x = torch.randn(batch, 1, 28, 28, device=dev)
    with torch.no_grad():
        bench.kernel_breakdown(model, sched, x, hint_host.to(dev), n_steps=1)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "traffic_labels.json"), "w") as f:
        json.dump({"batch": batch, "launches": bench.LAST_ORDINALS}, f)
    print(len(bench.LAST_ORDINALS), "labelled launches")


def merge(csv_path, labels_path):
    lab = json.load(open(labels_path))
    rows, hdr = [], None
    for r in csv.reader(open(csv_path)):
        if hdr is None:
            if "Kernel Name" in r:
                hdr = r
            continue
        rows.append(dict(zip(hdr, r)))
    per_launch = collections.OrderedDict()          # ncu ID -> bytes (read + write)
    ours = ("cnb::", "tma::", "af16::", "atm::", "gnb::")
    for d in rows:
        if not any(ns in d.get("Kernel Name", "") for ns in ours):
            continue
        try:
            v = float(d["Metric Value"].replace(",", ""))
        except (KeyError, ValueError):
            continue
        unit = d.get("Metric Unit", "byte")
        v *= {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
        per_launch[int(d["ID"])] = per_launch.get(int(d["ID"]), 0.0) + v
    ids = sorted(per_launch)
    agg = collections.defaultdict(list)
    for ordinal, fam, shape in lab["launches"]:
        if ordinal < len(ids):
            agg[f"{fam}|{shape}|{lab['batch']}"].append(per_launch[ids[ordinal]])
    out = {"_source": f"{os.path.basename(csv_path)}: dram__bytes_read.sum + dram__bytes_write.sum per launch, mean over the "
                      f"launches of one eager step ({len(ids)} cnb:: launches profiled)"}
    out.update({k: round(sum(v) / len(v)) for k, v in sorted(agg.items())})
    with open(os.path.join(ROOT, "profiles", "ncu_traffic.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(len(out) - 1, "shapes written")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--merge":
        merge(sys.argv[2], sys.argv[3])
    else:
        capture(int(sys.argv[1]) if len(sys.argv) > 1 else 1024)
