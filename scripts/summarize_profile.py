#!/usr/bin/env python3
"""Turn ncu output (gpurun_out/) into the small tracked summaries under profiles/.

  python scripts/summarize_profile.py launches gpurun_out/launches_X.csv profiles/rNN_launches.md
  python scripts/summarize_profile.py full gpurun_out/prof_X.ncu-rep profiles/rNN_ncu_full.md
"""
import collections
import csv
import io
import re
import subprocess
import sys

OURS = re.compile(r"pb::|recon_kernel|riemann_kernel|riemann_redo_kernel|c2p_redo_kernel|divb_kernel|history_kernel|c2p_fixup_packed_kernel|floors_packed_kernel|arith_selftest_kernel|flux_kernel|c2p_kernel|c2p_fixup_kernel|floors_kernel|flags_scan_kernel|update_\w+_kernel|source_kernel|"
                  r"divergence_kernel|emf_kernel|ct_flux_kernel|fix_fluxes_kernel|halo_\w+_kernel|bc_kernel|p2c_kernel|dt_kernel|"
                  r"weighted_sum_kernel")


def launches(src, dst):
    rows = [r for r in csv.reader(l for l in open(src, errors="replace") if l.startswith('"'))]
    hdr = rows[0]
    ix = {h: i for i, h in enumerate(hdr)}
    per = collections.OrderedDict()
    total_all, n_all = 0.0, 0
    for r in rows[1:]:
        if len(r) < len(hdr) or r[ix["Metric Name"]] != "gpu__time_duration.sum":
            continue
        name = r[ix["Kernel Name"]]
        v = float(r[ix["Metric Value"]].replace(",", ""))
        unit = r[ix["Metric Unit"]]
        us = v / 1e3 if unit.startswith("n") else (v if unit.startswith("u") else v * 1e3)
        n_all += 1
        total_all += us
        if not OURS.search(name):
            continue
        base = re.sub(r"\(.*", "", name).replace("void ", "").replace("pb::", "")
        d = per.setdefault(base, [0, 0.0])
        d[0] += 1
        d[1] += us
    tot = sum(v[1] for v in per.values())
    with open(dst, "w") as f:
        f.write("# ncu launch list: per-kernel device time (gpu__time_duration.sum, --clock-control none)\n\n")
        f.write("Source: `%s`. %d launches captured in total (torch's own init/fill kernels excluded below); "
                "times under ncu are serialised and cold-cache: compare SHARES.\n\n" % (src, n_all))
        f.write("| kernel | launches | total ms | share of our kernels | avg us |\n|---|---:|---:|---:|---:|\n")
        for k, (n, us) in sorted(per.items(), key=lambda kv: -kv[1][1]):
            f.write("| `%s` | %d | %.2f | %.1f %% | %.1f |\n" % (k, n, us / 1e3, 100 * us / tot, us / n))
        f.write("| **sum** | %d | %.2f | 100 %% | |\n" % (sum(v[0] for v in per.values()), tot / 1e3))
    print(open(dst).read())


METRICS = [
    ("gpu__time_duration.sum", "time"),
    ("launch__registers_per_thread", "regs"),
    ("launch__block_size", "block"),
    ("launch__grid_size", "grid"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue active %"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "FP64 pipe %"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("smsp__inst_executed.sum", "warp instr"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "threads/instr"),
]


def full(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    with open(dst, "w") as f:
        f.write("# ncu --set full --clock-control none: one RK stage of the main kernels\n\nSource: `%s`\n\n" % src)
        for r in data:
            f.write("## `%s`\n\n" % r[ix["Kernel Name"]][:90])
            f.write("| metric | value |\n|---|---|\n")
            for m, label in METRICS:
                if m in ix:
                    f.write("| %s (`%s`) | %s %s |\n" % (label, m, r[ix[m]], units[ix[m]]))
            st = [(float(r[i].replace(",", "")), h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""))
                  for h, i in ix.items() if h.startswith("smsp__average_warps_issue_stalled") and r[i] not in ("", "n/a")]
            f.write("| stall cycles per issued instruction | %s |\n\n" % ", ".join("%s %.2f" % (h, v) for v, h in sorted(st, reverse=True)[:6]))
    print(open(dst).read()[:3000])


def _num(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return None


def tojson(src, dst):
    """profiles/ncu_kernel_metrics.json: per kernel (template arguments stripped) the averages over the captured launches of
    dram bytes per launch, FP64-pipe utilisation, duration, registers, threads per instruction — what bench.py quotes"""
    import json
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    acc = collections.OrderedDict()
    for r in data:
        name = re.sub(r"[<(].*", "", r[ix["Kernel Name"]]).replace("void ", "").replace("pb::", "").strip()
        d = acc.setdefault(name, collections.defaultdict(list))
        rd = _num(r[ix["dram__bytes_read.sum"]]) * scale.get(units[ix["dram__bytes_read.sum"]], 1.0)
        wr = _num(r[ix["dram__bytes_write.sum"]]) * scale.get(units[ix["dram__bytes_write.sum"]], 1.0)
        d["dram_bytes_per_launch"].append(rd + wr)
        t = _num(r[ix["gpu__time_duration.sum"]])
        u = units[ix["gpu__time_duration.sum"]]
        d["duration_ms_under_ncu"].append(t * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1e-6))
        for m, key in (("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64_pipe_pct"),
                       ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct"),
                       ("launch__registers_per_thread", "registers"),
                       ("smsp__thread_inst_executed_per_inst_executed.ratio", "threads_per_instruction"),
                       ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct")):
            if m in ix and _num(r[ix[m]]) is not None:
                d[key].append(_num(r[ix[m]]))
    res = {"source": "profiles/%s (ncu --set full --clock-control none of `python bench.py`)" % dst.split("/")[-1].replace(".json", ".md"),
           "kernels": {k: dict({kk: sum(v) / len(v) for kk, v in d.items()}, launches=len(d["dram_bytes_per_launch"])) for k, d in acc.items()}}
    with open(dst, "w") as f:
        json.dump(res, f, indent=1)
    print(json.dumps(res, indent=1)[:2500])


if __name__ == "__main__":
    {"launches": launches, "full": full, "json": tojson}[sys.argv[1]](sys.argv[2], sys.argv[3])
