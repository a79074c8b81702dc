"""Turns the raw artefacts a gpurun profiling call brought back (gpurun_out/<round>/: bench JSON
lines, the ncu launch list, the ncu --set full report) into the small, committed summaries under
profiles/.  Usage (here, no GPU needed):  python profiles/summarize.py gpurun_out/r1 r1"""
import collections
import csv
import json
import os
import shutil
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_elapsed",
    "smsp__inst_executed.sum", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
]


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        k = r[ki].split("(")[0][-70:]
        a = agg.setdefault(k, [0.0, 0])
        a[0] += v
        a[1] += 1
    tot = sum(v[0] for v in agg.values())
    out = ["| share | total us | launches | kernel |", "|---|---|---|---|"]
    for k, (v, c) in sorted(agg.items(), key=lambda x: -x[1][0]):
        out.append(f"| {100 * v / tot:.1f}% | {v / 1e3:.1f} | {c} | `{k}` |")
    return "\n".join(out)


def full_report(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    out = []
    for r in rows[2:]:
        out.append(f"\n**{r[idx['Kernel Name']].split('(')[0]}**  (grid {r[idx['Grid Size']]}, block {r[idx['Block Size']]})\n")
        out.append("| metric | value | unit |\n|---|---|---|")
        for m in METRICS:
            if m in idx:
                out.append(f"| {m} | {r[idx[m]]} | {units[idx[m]]} |")
    return "\n".join(out)


def ncu_metrics(path, key):
    """{kernel family: counters per launch} from an `ncu --set full` report, for bench.py's roofline block
    (profiles/<tag>_ncu_metrics.json, keyed by "<workload>/D<colour>")."""
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}

    def val(r, name, scale_units=True):
        if name not in idx or r[idx[name]] in ("", "n/a"):
            return None
        v = float(r[idx[name]].replace(",", ""))
        u = units[idx[name]]
        if scale_units:
            v *= {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0, "ms": 1.0, "us": 1e-3, "ns": 1e-6}.get(u, 1.0)
        return v

    out = {}
    for r in rows[2:]:
        name = r[idx["Kernel Name"]]
        fam = next((f for f in ("blend_sh_fwd", "blend_sh_bwd", "blend_fwd", "blend_bwd", "fused_project_bwd", "fused_project",
                                "emit_keys", "tile_ranges", "pack_sorted") if f in name), None)
        if fam is None or fam in out:
            continue
        fam = {"blend_sh_fwd": "blend_fwd", "blend_sh_bwd": "blend_bwd"}.get(fam, fam)
        rd, wr = val(r, "dram__bytes_read.sum"), val(r, "dram__bytes_write.sum")
        out[fam] = {
            "kernel": name.split("(")[0][-80:], "ms": val(r, "gpu__time_duration.sum"),
            "dram_bytes": (rd or 0) + (wr or 0), "warp_inst": val(r, "smsp__inst_executed.sum", False),
            "issue_active_frac": (val(r, "smsp__issue_active.avg.pct_of_peak_sustained_active", False) or 0) / 100.0,
            "xu_pipe_frac": (val(r, "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", False) or 0) / 100.0,
            "fma_pipe_frac": (val(r, "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", False) or 0) / 100.0,
            "dram_throughput_frac": (val(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", False) or 0) / 100.0,
            "registers": val(r, "launch__registers_per_thread", False),
            "source": f"profiles/{os.path.basename(path)} ({key}; ncu --set full --clock-control none)"}
    return out


def main(src, tag):
    here = os.path.dirname(os.path.abspath(__file__))
    md = [f"# profiles — round {tag}", "",
          "Raw artefacts come from `gpurun` calls (B200, `--clock-control none`); numbers printed by a run under",
          "ncu are never bench values.  Bench lines are the unmodified stdout of `bench.py`.", ""]
    for name in sorted(os.listdir(src)):
        p = os.path.join(src, name)
        if name.endswith(".json") and os.path.getsize(p) > 0:
            shutil.copyfile(p, os.path.join(here, f"{tag}_{name}"))
            try:
                d = json.loads(open(p).read().strip().splitlines()[-1])
            except Exception:
                continue
            md.append(f"## {name}")
            md.append(f"- value: **{d.get('value'):.2f} {d.get('unit')}** ({d.get('ms_per_step'):.3f} ms/step, n_gpus={d.get('n_gpus')}); "
                      f"e2e {d.get('e2e', {}).get('value')}")
            if "stage_ms" in d:
                md.append(f"- stage_ms: `{d['stage_ms']}`")
            if "roofline" in d:
                r = d["roofline"]
                md.append(f"- roofline ({r['kernel']}): achieved {r['achieved']:.1f} {r['unit']} of {r['peak']} "
                          f"({100 * r['frac']:.1f}% of {r['peak_source']}); traffic {r.get('traffic')}")
            if "clocks" in d:
                md.append(f"- clocks: `{d['clocks']}`")
            if "cpu_baseline" in d:
                md.append(f"- cpu_baseline: `{d['cpu_baseline']}`")
            md.append(f"- workload: {d.get('config', {}).get('workload')}; M={d.get('config', {}).get('tile_instances_M')}, "
                      f"M_eff={d.get('config', {}).get('tile_instances_consumed_M_eff')}")
            md.append("")
    lc = os.path.join(src, "launches.csv")
    if os.path.exists(lc):
        shutil.copyfile(lc, os.path.join(here, f"{tag}_launches.csv"))
        md += ["## ncu launch list (bench.py --steps 2 --warmup 1; cold-cache, serialised: compare shares)", "",
               launches(lc), ""]
    rep = os.path.join(src, "prof_blend.ncu-rep")
    if os.path.exists(rep):
        md += ["## ncu --set full, blend kernels (report kept out of git: 8 MB; regenerate with the command in "
               "gpurun_out/run_r1_profile.sh)", full_report(rep), ""]
    metrics = {}
    for name in sorted(os.listdir(src)):
        if name.startswith("prof_") and name.endswith(".ncu-rep"):
            # prof_<workload>_D<colour>[_anything].ncu-rep
            parts = name[:-8].split("_")
            key = f"{parts[1]}/{parts[2]}" if len(parts) >= 3 else "C3/D3"
            rep = os.path.join(src, name)
            md += [f"## ncu --set full: {name} ({key})", full_report(rep), ""]
            m = ncu_metrics(rep, key)
            metrics.setdefault(key, {}).update(m)
    if metrics:
        json.dump(metrics, open(os.path.join(here, f"{tag}_ncu_metrics.json"), "w"), indent=1)
    open(os.path.join(here, f"{tag}_summary.md"), "w").write("\n".join(md) + "\n")
    print("\n".join(md))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
