"""One-screen summary of the --set full captures of a round: python scripts/ncu_summary.py r02 > profiles/r02_ncu_full_summary.txt"""
import csv, glob, io, os, subprocess, sys
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size"]
print(f"# ncu --set full --clock-control none, one launch each (cold caches); {tag}")
for f in sorted(glob.glob(f"gpurun_out/{tag}_full_*.ncu-rep")):
    out = subprocess.run(["ncu", "-i", f, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    if len(rows) < 3:
        continue
    hdr, units, val = rows[0], rows[1], rows[2]
    name = os.path.basename(f)[len(tag) + 6:-8]
    k = val[hdr.index("Kernel Name")].split("(")[0]
    print(f"== {name}: {k}")
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            print(f"   {w:70s} {val[i]} {units[i]}")
