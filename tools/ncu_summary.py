"""Summarise an ncu report (--set full) of the bench command: one row per kernel, and the
traversal kernel's figures as profiles/traverse_profile.json (what bench.py's `roofline` reads).

    python tools/ncu_summary.py gpurun_out/prof_step_r02x.ncu-rep --bench gpurun_out/bench_prof_r02x.json \
        --build-id $(cat gpurun_out/build_id_r02x.txt) --md profiles/r02_x_kernels.md --traverse-json profiles/traverse_profile.json

--bench: the JSON line of the PLAIN run (never under ncu) of the same command with one launch per
step (python bench.py --frames 512 --chunk 512 ...): its work counters give node visits per launch.
Runs here (no GPU needed): `ncu -i report --page raw --csv` does the decoding.
"""
import argparse
import csv
import io
import json
import subprocess
import sys

WANT = {
    "time_us": ("gpu__time_duration.sum", 1e-3),
    "cycles": ("sm__cycles_elapsed.max", 1),
    "dram_read": ("dram__bytes_read.sum", None),
    "dram_write": ("dram__bytes_write.sum", None),
    "lsu_wf": ("l1tex__data_pipe_lsu_wavefronts.sum", 1),
    "shared_wf": ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", 1),
    "shared_ld_wf": ("l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum", 1),
    "shared_ld_conflicts": ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum", 1),
    "shared_atom_wf": ("l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum", 1),
    "shared_atom_conflicts": ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_atom.sum", 1),
    "tex_wf": ("l1tex__data_pipe_tex_wavefronts.sum", 1),
    "t_sectors": ("l1tex__t_sectors.sum", 1),
    "t_hit": ("l1tex__t_sectors_lookup_hit.sum", 1),
    "l2_sectors": ("l1tex__m_xbar2l1tex_read_sectors.sum", 1),
    "lsu_pipe_pct": ("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", 1),
    "tex_pipe_pct": ("l1tex__data_pipe_tex_wavefronts.avg.pct_of_peak_sustained_elapsed", 1),
    "issue_pct": ("sm__inst_issued.avg.pct_of_peak_sustained_active", 1),
    "warps_pct": ("sm__warps_active.avg.pct_of_peak_sustained_active", 1),
    "dram_pct": ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 1),
    "regs": ("launch__registers_per_thread", 1),
}
UNIT_SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "nsecond": 1.0, "usecond": 1e3, "msecond": 1e6, "second": 1e9,
              "ns": 1.0, "us": 1e3, "ms": 1e6, "s": 1e9}


def load(report):
    raw = subprocess.run(["ncu", "-i", report, "--page", "raw", "--csv"], capture_output=True, text=True)
    if raw.returncode != 0:
        sys.exit("ncu failed: " + raw.stderr[-500:])
    rows = list(csv.reader(io.StringIO(raw.stdout)))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        d = {"kernel": r[hdr.index("Kernel Name")], "grid": r[hdr.index("Grid Size")] if "Grid Size" in hdr else ""}
        for key, (metric, _) in WANT.items():
            if metric in hdr:
                i = hdr.index(metric)
                try:
                    v = float(r[i].replace(",", ""))
                except ValueError:
                    continue
                v *= UNIT_SCALE.get(units[i], 1.0)
                d[key] = v
        if "time_us" in d:
            d["time_us"] = d["time_us"] / 1e3  # ns -> us
        out.append(d)
    return out


def merge_extra(ks, path):
    """metrics of a second pass (`ncu --metrics ... --csv --log-file path` of the same command and kernel
    filter: long format, one row per launch and metric), matched to the report's launches in order"""
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[hi]
    ix = {h: i for i, h in enumerate(hdr)}
    per = {}
    for r in rows[hi + 1:]:
        if len(r) < len(hdr):
            continue
        per.setdefault(r[ix["ID"]], {"kernel": r[ix["Kernel Name"]]})[r[ix["Metric Name"]]] = float(r[ix["Metric Value"]].replace(",", ""))
    extra = list(per.values())
    inv = {metric: key for key, (metric, _) in WANT.items()}
    for d in ks:
        base = d["kernel"].split("(")[0].split("<")[0].split()[-1]
        for j, e in enumerate(extra):
            if base in e["kernel"]:
                for m, v in e.items():
                    if m in inv and inv[m] not in d:
                        d[inv[m]] = v
                extra.pop(j)
                break


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--extra-csv", default=None, help="a --metrics pass of the same launches (wavefront .sum counters the full set lacks)")
    ap.add_argument("--bench", default=None)
    ap.add_argument("--build-id", default="unknown")
    ap.add_argument("--md", default=None)
    ap.add_argument("--traverse-json", default=None)
    ap.add_argument("--n-sms", type=int, default=148)
    ap.add_argument("--label", default="")
    a = ap.parse_args()
    ks = load(a.report)
    if a.extra_csv:
        merge_extra(ks, a.extra_csv)
    lines = ["# ncu --set full, one launch per kernel%s (build %s)" % ((" — " + a.label) if a.label else "", a.build_id), "",
             "Rates are per cycle per SM (sm__cycles_elapsed.max x %d SMs).  Source: `%s` (not committed), decoded with `tools/ncu_summary.py`." % (a.n_sms, a.report), "",
             "| kernel | us | DRAM rd MB | DRAM wr MB | LSU wf/cyc | of which shared | TEX wf/cyc | L1 sectors/cyc (hit %) | L2->L1 sectors/cyc | shared ld conflicts | shared atom conflicts | issue % | warps % | regs |",
             "|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|"]
    for d in ks:
        cyc = d.get("cycles", 0) * a.n_sms
        if not cyc:
            continue
        r = lambda k: d.get(k, 0.0) / cyc  # noqa: E731
        hit = 100.0 * d.get("t_hit", 0) / d["t_sectors"] if d.get("t_sectors") else 0.0
        lc = "%.0f %% of %.1f M" % (100.0 * d.get("shared_ld_conflicts", 0) / d["shared_ld_wf"], d["shared_ld_wf"] / 1e6) if d.get("shared_ld_wf") else "-"
        ac = "%.0f %% of %.1f M" % (100.0 * d.get("shared_atom_conflicts", 0) / d["shared_atom_wf"], d["shared_atom_wf"] / 1e6) if d.get("shared_atom_wf") else "-"
        lines.append("| %s | %.1f | %.1f | %.1f | %.3f | %.3f | %.3f | %.3f (%.0f) | %.3f | %s | %s | %.1f | %.1f | %d |" % (
            d["kernel"][:48], d.get("time_us", 0), d.get("dram_read", 0) / 1e6, d.get("dram_write", 0) / 1e6, r("lsu_wf"), r("shared_wf"), r("tex_wf"),
            r("t_sectors"), hit, r("l2_sectors"), lc, ac, d.get("issue_pct", 0), d.get("warps_pct", 0), int(d.get("regs", 0))))
    text = "\n".join(lines) + "\n"
    if a.md:
        with open(a.md, "w") as f:
            f.write(text)
    print(text)
    if a.traverse_json:
        tr = [d for d in ks if "traverse_kernel" in d["kernel"]]
        if not tr:
            sys.exit("no traverse_kernel launch in the report")
        d = tr[0]
        visits = None
        if a.bench:
            b = json.loads([ln for ln in open(a.bench).read().splitlines() if ln.startswith("{")][-1])
            visits = b["work_per_step"]["node_visits"] / max(1, round(b["roofline"]["launches"] / b["steps"]))
        cyc = d["cycles"] * a.n_sms
        prof = {"kernel": d["kernel"], "build_id": a.build_id, "node_visits_per_launch": visits,
                "gpu_time_us": d.get("time_us"), "sm_cycles_elapsed": d["cycles"], "sms": a.n_sms,
                "dram_bytes_read": d.get("dram_read"), "dram_bytes_write": d.get("dram_write"),
                "dram_bytes_per_launch": d.get("dram_read", 0) + d.get("dram_write", 0),
                "lsu_wavefronts": d.get("lsu_wf"), "shared_wavefronts": d.get("shared_wf"), "tex_wavefronts": d.get("tex_wf"),
                "shared_ld_bank_conflicts": d.get("shared_ld_conflicts"),
                "lsu_wavefronts_per_visit": d.get("lsu_wf", 0) / visits if visits else None,
                "tex_wavefronts_per_visit": d.get("tex_wf", 0) / visits if visits else None,
                "lsu_wavefronts_per_cycle_per_sm": d.get("lsu_wf", 0) / cyc, "tex_wavefronts_per_cycle_per_sm": d.get("tex_wf", 0) / cyc,
                "l2_sectors_per_cycle_per_sm": d.get("l2_sectors", 0) / cyc, "l1_sector_hit_rate": d.get("t_hit", 0) / d["t_sectors"] if d.get("t_sectors") else None,
                "issue_active_pct": d.get("issue_pct"), "warps_active_pct": d.get("warps_pct"), "registers": d.get("regs"),
                "source": "ncu --set full --clock-control none, one launch (%s); report %s, decoded by tools/ncu_summary.py" % (a.label or "bench command", a.report)}
        with open(a.traverse_json, "w") as f:
            json.dump(prof, f, indent=1)
        print(json.dumps(prof, indent=1))


if __name__ == "__main__":
    main()
