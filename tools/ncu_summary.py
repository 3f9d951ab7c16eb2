"""Reads a .ncu-rep (ncu -i ... --page raw --csv) and prints the per-kernel numbers the rooflines are argued from:
    python tools/ncu_summary.py gpurun_out/r2j_static.ncu-rep > profiles/r2_ncu_static_summary.md"""
import csv
import subprocess
import sys

WANT = [
    ("time (us)", "gpu__time_duration.sum"),
    ("registers / thread", "launch__registers_per_thread"),
    ("dynamic smem / block (KB)", "launch__shared_mem_per_block_dynamic"),
    ("DRAM read (GB)", "dram__bytes_read.sum"),
    ("DRAM write (GB)", "dram__bytes_write.sum"),
    ("DRAM throughput % of peak", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    ("L2 sector hit rate %", "lts__t_sector_hit_rate.pct"),
    ("tensor pipe active % (under-reports cta_group::2: issuing SM only)", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
    ("TMEM pipe inst % of peak", "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active"),
    ("XU (MUFU) pipe % of peak", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"),
    ("SM throughput % of peak", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("warps active % of peak", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("warp instructions executed", "smsp__inst_executed.sum"),
    ("LSU shared-memory wavefronts % of peak", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"),
    ("tensor-core shared-memory wavefronts % of peak", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"),
    ("shared-memory bank conflicts (LSU)", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    print(f"# ncu --set full --clock-control none summary of `{rep}` (one profiled launch per row group; serialised,\n"
          "# cold-cache launches: compare SHARES and ratios, not absolute times)\n")
    for r in rows[2:]:
        name = r[ki].replace("<unnamed>::", "")
        print(f"## `{name[:100]}`\n")
        print("| metric | value |\n|---|---|")
        for label, key in WANT:
            if key in hdr:
                i = hdr.index(key)
                v = r[i]
                try:
                    v = f"{float(v.replace(',', '')):,.3f}".rstrip("0").rstrip(".")
                except ValueError:
                    pass
                print(f"| {label} | {v} {units[i]} |")
        print()


if __name__ == "__main__":
    main()
