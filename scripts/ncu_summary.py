"""Summarise an `ncu --set full` report into a small JSON (one record per profiled launch + per-kernel-class totals).

    python scripts/ncu_summary.py gpurun_out/prof_mlp_r2.ncu-rep profiles/r2_ncu_full_mlp_kernels.json \
        [--step-launches fwd:3,bwd:3,wgrad:2 --dram profiles/r2_dram_traffic.json]

Reads the report with `ncu -i <rep> --page raw --csv` (works without a GPU).  With --dram it also writes the DRAM
traffic (dram__bytes_read.sum + dram__bytes_write.sum) per kernel class summed over ONE training step's launches
(the first N launches of each class), which bench.py reports as roofline.traffic."""
import csv
import io
import json
import subprocess
import sys

KEEP = {
    "gpu__time_duration.sum": "duration",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed": "tensor_pipe_active_pct",
    "sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_elapsed": "tensor_pipe_hmma_active_pct",
    "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_elapsed": "tc_pipe_active_pct",
    "dram__bytes_read.sum": "dram_bytes_read",
    "dram__bytes_write.sum": "dram_bytes_write",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed": "dram_throughput_pct",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "gpu_dram_throughput_pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_throughput_pct",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed": "l1tex_throughput_pct",
    "l1tex__m_l1tex2xbar_write_bytes.sum.pct_of_peak_sustained_elapsed": "l1tex_to_xbar_write_pct",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed": "lsu_data_pipe_pct",
    "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed": "tensor_smem_read_pipe_pct",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_throughput_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "launch__registers_per_thread": "registers_per_thread",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "launch__cluster_size": "cluster",
    "smsp__cycles_active.avg": "smsp_cycles_active",
    "sm__cycles_elapsed.avg": "sm_cycles_elapsed",
}


def classify(name):
    for k in ("mlp_fwd", "mlp_bwd", "mlp_wgrad", "octree_render", "octree_train", "grid_weight"):
        if k in name:
            return k
    return "other"


def main():
    rep, out = sys.argv[1], sys.argv[2]
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    idx = {h: i for i, h in enumerate(hdr)}
    recs = []
    for r in rows[2:]:
        rec = {"kernel": r[idx["Kernel Name"]][:90]}
        for full, short in KEEP.items():
            if full in idx and r[idx[full]] != "":
                try:
                    rec[short] = float(r[idx[full]].replace(",", ""))
                except ValueError:
                    rec[short] = r[idx[full]]
                if short in ("duration", "dram_bytes_read", "dram_bytes_write"):
                    rec[short + "_unit"] = units[idx[full]]
        recs.append(rec)
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    for rec in recs:
        for k in ("dram_bytes_read", "dram_bytes_write"):
            if k in rec:
                rec[k] = rec[k] * scale.get(rec.pop(k + "_unit", "byte"), 1.0)
        if "duration" in rec:
            u = rec.pop("duration_unit", "ns")
            rec["duration_ms"] = rec.pop("duration") * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1e-6)
    json.dump({"source": rep, "launches": recs}, open(out, "w"), indent=1)
    print(f"{out}: {len(recs)} launches")
    if "--dram" in sys.argv:
        want = dict(kv.split(":") for kv in sys.argv[sys.argv.index("--step-launches") + 1].split(","))
        tot, cnt, ms = {}, {}, {}
        for rec in recs:
            c = classify(rec["kernel"])
            key = c.replace("mlp_", "")
            if key in want and cnt.get(c, 0) < int(want[key]):
                cnt[c] = cnt.get(c, 0) + 1
                tot[c] = tot.get(c, 0.0) + rec.get("dram_bytes_read", 0.0) + rec.get("dram_bytes_write", 0.0)
                ms[c] = ms.get(c, 0.0) + rec.get("duration_ms", 0.0)
        tot["note"] = ("dram__bytes_read.sum + dram__bytes_write.sum per kernel class over one 4096-ray training step "
                       f"({want}); ncu --set full, clocks not controlled; step total = {sum(v for v in tot.values() if isinstance(v, float)):.4g} B")
        tot["ncu_ms_per_class"] = ms
        json.dump(tot, open(sys.argv[sys.argv.index("--dram") + 1], "w"), indent=1)
        print(tot)


if __name__ == "__main__":
    main()
