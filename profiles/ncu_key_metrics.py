"""Print the key metrics of every kernel in an .ncu-rep (read with `ncu -i ... --page raw --csv`)."""
import csv
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "dur"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
    ("lts__t_bytes.sum", "l2_bytes"),
    ("lts__t_sector_hit_rate.pct", "l2hit%"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("smsp__inst_executed.sum", "winst"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "st_long"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "st_short"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "st_bar"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "st_lg"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "st_mio"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "st_math"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "st_wait"),
    ("smsp__average_warps_issue_stalled_membar_per_issue_active.ratio", "st_membar"),
]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    name_i = hdr.index("Kernel Name")
    for r in rows[2:]:
        print("==", r[name_i].split("(")[0].replace("void ", "").replace("<unnamed>::", "")[:60])
        parts = []
        for key, short in WANT:
            if key in hdr:
                i = hdr.index(key)
                parts.append("%s=%s%s" % (short, r[i], units[i] if units[i] not in ("", "%") else ""))
        print("   " + "  ".join(parts))


if __name__ == "__main__":
    main(sys.argv[1])
