"""Summarise an ncu report (.ncu-rep, read here without a GPU) into the JSON kept under profiles/.

    python tools/ncu_summary.py gpurun_out/x.ncu-rep profiles/r02_x.json "command that was profiled" [kernel-regex]

Per kernel launch in the report: the roofline-relevant counters (duration, clocks, DRAM bytes, L2, tensor pipes incl. the
FP64 DMMA and INT8 IMMA sub-pipes, shared-memory wavefronts of the tensor core, TMA bytes, registers) and the warp-stall
sampling shares from the source page.
"""
import csv
import io
import json
import re
import subprocess
import sys

KEEP = [
    r"^gpu__time_duration\.sum$", r"^gpc__cycles_elapsed\.max(\.per_second)?$", r"^sm__cycles_elapsed\.max$", r"^sm__cycles_active\.avg$",
    r"^dram__bytes_(read|write)\.sum(\.per_second)?$", r"^gpu__dram_throughput\.avg\.pct_of_peak_sustained_elapsed$",
    r"^dram__throughput\.avg\.pct_of_peak_sustained_elapsed$", r"^lts__t_sector_hit_rate\.pct$", r"^lts__throughput\.avg\.pct_of_peak_sustained_elapsed$",
    r"^lts__t_sectors_srcunit_tex\.avg\.pct_of_peak_sustained_elapsed$", r"^l1tex__m_xbar2l1tex_read_bytes\.sum(\.per_second|\.pct_of_peak_sustained_elapsed)?$",
    r"^l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld\.sum$", r"^l1tex__data_pipe_tc_wavefronts_mem_shared\.sum(\.pct_of_peak_sustained_elapsed)?$",
    r"^l1tex__data_bank_conflicts_pipe_lsu_mem_shared\.sum$", r"^sm__pipe_tensor_cycles_active.*pct_of_peak_sustained_(elapsed|active)$",
    r"^sm__pipe_tensor_subpipe_(imma|dmma|hmma)_cycles_active\.avg\.pct_of_peak_sustained_active$",
    r"^sm__inst_executed_pipe_tensor_subpipe_dmma\.avg\.pct_of_peak_sustained_active$", r"^sm__ops_path_tensor_src_fp64\.sum(\.pct_of_peak_sustained_elapsed|\.per_second)?$",
    r"^sm__ops_path_tensor_src_int8\.sum(\.pct_of_peak_sustained_elapsed|\.per_second)?$", r"^sm__pipe_fp64_cycles_active\.avg\.pct_of_peak_sustained_elapsed$",
    r"^sm__throughput\.avg\.pct_of_peak_sustained_elapsed$", r"^sm__warps_active\.avg\.pct_of_peak_sustained_active$",
    r"^launch__(grid_size|block_size|registers_per_thread|shared_mem_per_block_dynamic|cluster_.*)$", r"^smsp__inst_executed\.sum$",
]


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return out


def main():
    rep, dst, cmd = sys.argv[1:4]
    pat = re.compile(sys.argv[4]) if len(sys.argv) > 4 else None
    rows = list(csv.reader(io.StringIO(page(rep, "raw"))))
    hdr, units = rows[0], rows[1]
    name_col = hdr.index("Kernel Name")
    kernels = []
    for r in rows[2:]:
        if len(r) != len(hdr) or (pat and not pat.search(r[name_col])):
            continue
        k = {"kernel": r[name_col].split("(")[0]}
        for h, u, v in zip(hdr, units, r):
            if any(re.search(p, h) for p in KEEP) and v != "":
                k[h] = (v + " " + u).strip()
        kernels.append(k)
    # warp-stall sampling per kernel from the source page (one block per kernel)
    src = page(rep, "source")
    blocks = re.split(r'^"Kernel Name",', src, flags=re.M)[1:]
    stalls = []
    for b in blocks:
        lines = list(csv.reader(io.StringIO('"Kernel Name",' + b)))
        kname = lines[0][1] if len(lines[0]) > 1 else ""
        if pat and not pat.search(kname):
            continue
        h = lines[1]
        cols = [i for i, c in enumerate(h) if c.startswith("stall_") and "Not Issued" not in c]
        tot = {h[i]: 0 for i in cols}
        for ln in lines[2:]:
            if len(ln) == len(h):
                for i in cols:
                    try:
                        tot[h[i]] += int(ln[i])
                    except ValueError:
                        pass
        s = sum(tot.values()) or 1
        stalls.append({k: round(100.0 * v / s, 1) for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:8]})
    for k, st in zip(kernels, stalls):
        k["warp_stall_sampling_pct"] = st
    json.dump({"command": cmd, "report": rep, "kernels": kernels}, open(dst, "w"), indent=1)
    print("wrote", dst, "with", len(kernels), "kernel(s)")


if __name__ == "__main__":
    main()
