"""Summarise ncu exports into profiles/*.md.

    python tools/ncu_summary.py launches <launches.csv> <out.md> [title]
    python tools/ncu_summary.py full <report.ncu-rep> <out.md> [title]
    python tools/ncu_summary.py traffic <dram_metrics.csv> <out.json> [note]
"""
import csv
import re
import subprocess
import sys
from collections import defaultdict


def short(name):
    name = re.sub(r"<unnamed>::", "", name)
    name = re.sub(r"\(.*", "", name)
    return name[:110]


def launches(path, out, title):
    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 5]
    hdr = rows[0]
    i_name, i_val, i_metric = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
    agg = defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        if r[i_metric] != "gpu__time_duration.sum":
            continue
        k = short(r[i_name])
        agg[k][0] += 1
        agg[k][1] += float(r[i_val].replace(",", ""))
    unit = rows[1][hdr.index("Metric Unit")]
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(unit, 1e-3)
    tot = sum(v[1] for v in agg.values()) * scale
    ours = sum(v[1] for k, v in agg.items() if is_ours(k)) * scale
    with open(out, "w") as fh:
        fh.write(f"# {title}\n\nsource: `{path}` (ncu --metrics gpu__time_duration.sum --clock-control none; per-launch times are "
                 f"cold-cache and serialised: compare shares, not absolutes)\n\n")
        fh.write(f"launches: {sum(v[0] for v in agg.values())}, summed kernel time {tot/1e3:.2f} ms; libsd_b200 kernels "
                 f"{ours/1e3:.2f} ms ({100*ours/tot:.1f} %)\n\n| kernel | launches | total us | share |\n|---|---:|---:|---:|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
            fh.write(f"| {'**' if is_ours(k) else ''}{k}{'**' if is_ours(k) else ''} | {v[0]} | {v[1]*scale:.1f} | {100*v[1]*scale/tot:.2f} % |\n")
    print(open(out).read()[:3000])


def is_ours(k):
    return any(s in k for s in ("gemm_tc_kernel", "gemm_f32_kernel", "attn_", "ln_stats", "ln_bwd", "sampler", "adamw", "colsum",
                                "q_sample", "ddim_step", "mse_", "copy_rows", "step_token", "gather_rows", "scatter_add",
                                "dropout_", "affine_joints", "add_kernel", "kv_relayout", "transpose_kernel", "layer_", "bn_reduce",
                                "bn_apply", "bn_bwd", "bn_finalize", "bn_param", "stem_", "maxpool_", "ca_fwd", "ca_bwd", "kv_proj", "kv_dgrad",
                                "wgrad_tma", "pack_bf16", "cast_bf16", "bcast_row", "enc_layer"))


WANT = ["launch__grid_size", "launch__block_size", "launch__registers_per_thread", "gpu__time_duration.sum",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "lts__t_bytes.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio"]


def full(rep, out, title):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    with open(out, "w") as fh:
        fh.write(f"# {title}\n\nsource: `{rep}` (ncu --set full --clock-control none --import-source on)\n\n")
        for r in rows[2:]:
            fh.write(f"## {short(r[idx['Kernel Name']])}\n\n| metric | value |\n|---|---|\n")
            for w in WANT:
                if w in idx:
                    fh.write(f"| {w} | {r[idx[w]]} {units[idx[w]]} |\n")
            fh.write("\n")
    print(open(out).read()[:2500])




def traffic(path, out_json, note):
    """`ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --csv` of one training step ->
    DRAM bytes per bench kernel class (the classes of soccerdiffusion_b200.ops._Timed), per class launch."""
    import json

    rows = [r for r in csv.reader(open(path, errors="replace")) if len(r) > 5]
    hdr = rows[0]
    i_name, i_val, i_metric, i_unit = (hdr.index(k) for k in ("Kernel Name", "Metric Value", "Metric Name", "Metric Unit"))
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3}
    # bench class <- kernels it launches (one class launch = one sd_* entry point call)
    classes = {"bn_bwd": ("bn_reduce_kernel<1>", "bn_bwd_apply_kernel", "bn_param_grads_kernel"),
               "bn_stats": ("bn_reduce_kernel<0>", "bn_finalize_kernel"),
               "bn_apply": ("bn_apply_kernel",),
               "stem_bn_relu_pool_fwd": ("stem_fwd_band_kernel",),
               "stem_bn_relu_pool_bwd": ("stem_bwd_block_kernel", "stem_bwd_sums_pooled_kernel"),
               "stem_fprop_s2d": ("stem_fprop_tma_kernel",),
               "stem_wgrad_s2d": ("stem_wgrad_tma_kernel",),
               "stem_pack_s2d": ("stem_pack_kernel",)}
    primary = {"bn_bwd": "bn_bwd_apply_kernel", "bn_stats": "bn_reduce_kernel<0>", "stem_bn_relu_pool_bwd": "stem_bwd_block_kernel<1>"}
    agg = {c: {"dram_bytes": 0.0, "us": 0.0, "launches": 0} for c in classes}
    for r in rows[1:]:
        name = r[i_name]
        for c, kernels in classes.items():
            if any(k in name for k in kernels):
                v = float(r[i_val].replace(",", "")) * scale.get(r[i_unit], 1.0)
                if r[i_metric].startswith("dram__bytes"):
                    agg[c]["dram_bytes"] += v
                elif r[i_metric] == "gpu__time_duration.sum":
                    agg[c]["us"] += v
                    if primary.get(c, kernels[0]) in name:
                        agg[c]["launches"] += 1
    res = {c: {"dram_bytes_per_launch": v["dram_bytes"] / v["launches"], "launches": v["launches"],
               "kernel_us_per_launch": v["us"] / v["launches"]} for c, v in agg.items() if v["launches"]}
    res["_source"] = note
    with open(out_json, "w") as fh:
        json.dump(res, fh, indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    kind, src, out = sys.argv[1:4]
    title = sys.argv[4] if len(sys.argv) > 4 else src
    {"launches": launches, "full": full, "traffic": traffic}[kind](src, out, title)
