"""Turn the two ncu outputs of a bench step into the markdown kept under profiles/:

    python tools/ncu_summary.py launches gpurun_out/r2_launches.csv         # `--metrics gpu__time_duration.sum` launch list
    python tools/ncu_summary.py full gpurun_out/prof_r2.ncu-rep             # `--set full` capture (needs ncu here)
"""
import collections
import csv
import io
import subprocess
import sys

UNIT = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3, "us": 1.0, "ms": 1e3, "ns": 1e-3, "usecond": 1.0, "msecond": 1e3, "nsecond": 1e-3}


def launches(path):
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 5]
    hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    H = rows[hdr]
    iK, iV, iU = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
    per = collections.OrderedDict()
    for r in rows[hdr + 1:]:
        try:
            us = float(r[iV].replace(",", "")) * UNIT.get(r[iU], 1.0)
        except ValueError:
            continue
        k = r[iK][:100]
        per.setdefault(k, []).append(us)
    return per


def full(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    H, U = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(H)}

    def val(r, c, to):
        if c not in ix or r[ix[c]] in ("", "n/a"):
            return float("nan")
        return float(r[ix[c]].replace(",", "")) * (UNIT.get(U[ix[c]], 1.0) if to else 1.0)
    print("| # | kernel | grid x block | time us | DRAM read MB | DRAM write MB | traffic MB | DRAM % | SM % | issue % | warps active % | regs | tensor pipe % |")
    print("|---|---|---|---|---|---|---|---|---|---|---|---|---|")
    agg = collections.OrderedDict()
    for n, r in enumerate(rows[2:]):
        name = r[ix["Kernel Name"]].replace("<unnamed>::", "").split("(")[0].replace("void ", "")
        t = val(r, "gpu__time_duration.sum", True)
        rd, wr = val(r, "dram__bytes_read.sum", True), val(r, "dram__bytes_write.sum", True)
        print(f"| {n} | `{name}` | {r[ix['launch__grid_size']]} x {r[ix['launch__block_size']]} | {t:.1f} | {rd:.1f} | {wr:.1f} | {rd + wr:.1f} | "
              f"{val(r, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', False):.1f} | {val(r, 'sm__throughput.avg.pct_of_peak_sustained_elapsed', False):.1f} | "
              f"{val(r, 'smsp__issue_active.avg.pct_of_peak_sustained_active', False):.1f} | {val(r, 'sm__warps_active.avg.pct_of_peak_sustained_active', False):.1f} | "
              f"{r[ix['launch__registers_per_thread']]} | {val(r, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed', False):.1f} |")
        fam = name.split("<")[0]
        a = agg.setdefault(fam, [0, 0.0, 0.0])
        a[0] += 1; a[1] += rd + wr; a[2] += t
    print("\n| kernel family | launches | traffic MB (sum) | traffic MB / launch | time us (sum) |\n|---|---|---|---|---|")
    for fam, (n, tr, t) in agg.items():
        print(f"| `{fam}` | {n} | {tr:.1f} | {tr / n:.1f} | {t:.1f} |")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        per = launches(sys.argv[2])
        tot = sum(sum(v) for v in per.values())
        print(f"launches: {sum(len(v) for v in per.values())}; sum of kernel time {tot / 1e3:.3f} ms\n")
        print("| kernel | launches | total us | share |\n|---|---|---|---|")
        for k, v in sorted(per.items(), key=lambda kv: -sum(kv[1])):
            print(f"| `{k}` | {len(v)} | {sum(v):.1f} | {100 * sum(v) / tot:.1f} % |")
    else:
        full(sys.argv[2])
