"""Summaries of `ncu --set full` reports for profiles/ (read here, on the CPU box, with `ncu -i ... --page raw --csv`).

    python scripts/ncu_summary.py conv gpurun_out/r2_conv.ncu-rep profiles/r2_conv_ncu_full_summary.csv [precision]
        one row per conv launch of one forward (18 launches, layer names attached in order); with a precision
        name also (re)writes that precision's entry of profiles/conv_traffic.json (DRAM bytes per forward)
    python scripts/ncu_summary.py geom gpurun_out/r2_geom.ncu-rep profiles/r2_geom_ncu_full_summary.csv
        the geometry kernels (prep_images, psv_gather_pair, render_composite_v2): time, DRAM bytes, instruction
        counts, issue utilisation, L1 / L2 hit rates, top stall reasons
"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LAYERS = ["conv1_1", "conv1_2", "conv2_1", "conv2_2", "conv3_1", "conv3_2", "conv3_3", "conv4_1", "conv4_2", "conv4_3",
          "conv6_1", "conv6_2", "conv6_3", "conv7_1", "conv7_2", "conv8_1", "conv8_2", "color_pred"]
COMMON = ["launch__grid_size", "launch__cluster_dim_x", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
          "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
          "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "launch__registers_per_thread",
          "launch__shared_mem_per_block_dynamic"]
CONV = ["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"]
GEOM = ["smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
STALL_PRE, STALL_SUF = "smsp__average_warps_issue_stalled_", "_per_issue_active.ratio"


def raw_rows(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return rows[0], rows[1], rows[2:]


def num(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return None


def to_bytes(v, unit):
    return v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)


def main():
    mode, rep, dst = sys.argv[1], sys.argv[2], sys.argv[3]
    precision = sys.argv[4] if len(sys.argv) > 4 else None
    hdr, units, rows = raw_rows(rep)
    idx = {h: i for i, h in enumerate(hdr)}
    cols = COMMON + (CONV if mode == "conv" else GEOM)
    cols = [c for c in cols if c in idx]
    stall_idx = [i for i, h in enumerate(hdr) if h.startswith(STALL_PRE) and h.endswith(STALL_SUF)]
    with open(dst, "w", newline="") as fh:
        w = csv.writer(fh)
        w.writerow((["layer"] if mode == "conv" else []) + ["kernel"] + cols + (["top_stalls"] if mode == "geom" else []))
        w.writerow(([""] if mode == "conv" else []) + [""] + [units[idx[c]] for c in cols] + ([""] if mode == "geom" else []))
        per_layer, total = {}, 0.0
        n = len(rows)
        for k, r in enumerate(rows):
            name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "").replace("msi::<unnamed>::", "")
            line = ([LAYERS[k - (n - len(LAYERS))] if n >= len(LAYERS) and k >= n - len(LAYERS) else ""] if mode == "conv" else []) + [name]
            line += [r[idx[c]] for c in cols]
            if mode == "geom":
                st = sorted(((num(r[i]) or 0.0, hdr[i][len(STALL_PRE):-len(STALL_SUF)]) for i in stall_idx), reverse=True)[:4]
                line.append("; ".join(f"{n_} {v:.2f}" for v, n_ in st))
            w.writerow(line)
            if mode == "conv" and line[0]:
                rd = to_bytes(num(r[idx["dram__bytes_read.sum"]]), units[idx["dram__bytes_read.sum"]])
                wr = to_bytes(num(r[idx["dram__bytes_write.sum"]]), units[idx["dram__bytes_write.sum"]])
                per_layer[line[0]] = {"dram_read_bytes": rd, "dram_write_bytes": wr,
                                      "tensor_active_pct": num(r[idx[CONV[0]]]), "duration_us": num(r[idx["gpu__time_duration.sum"]])}
                total += rd + wr
    print("wrote", dst, len(rows), "launches")
    if mode == "conv" and precision:
        p = os.path.join(ROOT, "profiles", "conv_traffic.json")
        try:
            doc = json.load(open(p))
        except Exception:
            doc = {}
        if "by_precision" not in doc:
            doc = {"by_precision": {}}
        doc["by_precision"][precision] = {
            "source": f"ncu --set full --clock-control none, one forward (18 conv launches, the shipped kernels), 640x320, 32 spheres, "
                      f"B=1, {precision}; {os.path.relpath(dst, ROOT)}",
            "workload": {"H": 320, "W": 640, "P": 32, "B": 1, "precision": precision},
            "dram_bytes_per_forward": total, "per_layer": per_layer}
        json.dump(doc, open(p, "w"), indent=1)
        print("updated", p, precision, total / 1e6, "MB")


if __name__ == "__main__":
    main()
